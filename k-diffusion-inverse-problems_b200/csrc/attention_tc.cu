// QKVAttentionLegacy (guided_diffusion/unet.py:339-356) on the 5th-gen tensor cores, head width 64, T <= 256 tokens
// (the 16x16 and 8x8 attention levels of the 256x256 ADM UNet).
//
// One CTA per (image, head).  Q, K, V tiles ([T][64] bf16 = rows of 128 bytes) are TMA-loaded straight out of the interleaved
// qkv activation [N,T,3C] (channel = head*192 + {q,k,v}*64 + c, unet.py:349) into 128B-swizzled shared memory.
//   S = Q K^T          tcgen05.mma, M = 128 queries, N = T keys, K = 64; both operands K-major, accumulator in TMEM
//   P = softmax(S/8)   each of the 128 threads owns one TMEM lane = one query row (fp32, two passes over TMEM), writes P as
//                      bf16 into shared memory in the K-major A-operand layout (64-key panels of [128][128 B], swizzled)
//   O = P V            tcgen05.mma, M = 128, N = 64, K = T; V is used in place as an MN-major B operand (d contiguous)
// and the backward (autograd of the same, condition/condition.py:136,146,155,172,269):
//   D = rowsum(dO o O);  P = exp(S/8 - lse);  dS = P o (dP - D)
//   per (128-key block kb, 128-query block qb):  S = Q K^T, dP = dO V^T  ->  P, dS (bf16, shared memory)
//   dV_kb += P^T dO,  dK_kb += dS^T Q  (MN-major A operands: the same [q][keys] panels read "transposed"),  dQ_qb += dS K
#include <stdlib.h>

#include "attention_tc.cuh"

namespace kdip {

static constexpr int ATC_THREADS = 128;

struct AttnTcParams {
  CUtensorMap map_q;    // qkv as [N*T][3C], box {64, 128}
  CUtensorMap map_kv;   // box {64, T}
  int T, heads;
  bf16* out;
  float* lse;
};

__global__ void __launch_bounds__(ATC_THREADS, 1) attn_fwd_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int T = p.T, h = blockIdx.x, n = blockIdx.y;
  const int nqb = (T + 127) / 128;
  const int q_rows = nqb * 128;
  uint8_t* Qs = smem;                               // [q_rows][128 B]
  uint8_t* Ks = Qs + q_rows * 128;                  // [T][128 B]
  uint8_t* Vs = Ks + T * 128;                       // [T][128 B]
  uint8_t* Ps = Vs + T * 128;                       // [T/64 panels][128 rows][128 B]
  uint64_t* load_bar = reinterpret_cast<uint64_t*>(Ps + (T / 64) * 16384);
  uint64_t* mma_bar = load_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_kv);
    mbar_init(load_bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = T <= 128 ? 256u : 512u;     // S: T columns at 0, O: 64 columns at 128 / 256
  if (warp == 0) { tmem_alloc(tmem_ptr_smem, tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + (T <= 128 ? 128u : 256u);

  const int col0 = h * 192, row0 = n * T;
  if (tid == 0) {
    mbar_arrive_expect_tx(load_bar, (uint32_t)((q_rows + 2 * T) * 128));
    for (int qb = 0; qb < nqb; ++qb) tma_load_2d(Qs + qb * 16384, &p.map_q, load_bar, col0, row0 + qb * 128);
    tma_load_2d(Ks, &p.map_kv, load_bar, col0 + 64, row0);
    tma_load_2d(Vs, &p.map_kv, load_bar, col0 + 128, row0);
  }
  mbar_wait(load_bar, 0);

  const float scale = 0.125f * 1.4426950408889634f;   // (ch^-1/4)^2 = 1/8, folded with log2(e) for exp2f
  const uint32_t idesc_s = umma_idesc_bf16(128, T);
  const uint32_t idesc_o = umma_idesc_bf16_major(128, 64, 0, 1);
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
  const int r = tid;
  const uint32_t swz = (uint32_t)(r & 7);
  uint32_t mma_phase = 0;
  const int C = p.heads * 64;

  for (int qb = 0; qb < nqb; ++qb) {
    if (tid == 0) {
      const uint64_t qd = umma_desc_sw128(smem_u32(Qs + qb * 16384)), kd = umma_desc_sw128(smem_u32(Ks));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_S, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
      umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // pass 1: row maximum
    float mx = -INFINITY;
    for (int c = 0; c < T; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S + lane_base + (uint32_t)c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
    const float m2 = mx * scale;   // in log2 units
    // pass 2: probabilities -> bf16 panels, row sum
    float l = 0.f;
    for (int c = 0; c < T; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S + lane_base + (uint32_t)c, v);
      tmem_ld_wait();
      uint8_t* prow = Ps + (c >> 6) * 16384 + r * 128;
      const int j0 = (c & 32) ? 4 : 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          e[i] = exp2f(fmaf(__uint_as_float(v[j * 8 + i]), scale, -m2));
          l += e[i];
        }
        uint4 u;
        u.x = pack_bf16(e[0], e[1]); u.y = pack_bf16(e[2], e[3]); u.z = pack_bf16(e[4], e[5]); u.w = pack_bf16(e[6], e[7]);
        *reinterpret_cast<uint4*>(prow + (((uint32_t)(j0 + j) ^ swz) << 4)) = u;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      for (int kk = 0; kk < T / 16; ++kk) {
        const uint64_t pd = umma_desc_sw128(smem_u32(Ps + (kk >> 2) * 16384)) + (uint64_t)(2 * (kk & 3));
        const uint64_t vd = umma_desc_sw128_ls(smem_u32(Vs + kk * 2048), 16, 1024);
        umma_bf16_ss(tmem_O, pd, vd, idesc_o, kk ? 1u : 0u);
      }
      umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    const int q = qb * 128 + r;
    uint32_t o0[32], o1[32];
    tmem_ld_32x32(tmem_O + lane_base, o0);
    tmem_ld_32x32(tmem_O + lane_base + 32u, o1);
    tmem_ld_wait();
    if (q < T) {
      const float inv = 1.f / l;
      bf16* op = p.out + ((size_t)(row0 + q)) * C + h * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(j < 4 ? o0[j * 8 + i] : o1[(j - 4) * 8 + i]) * inv;
        uint4 u;
        u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
        *reinterpret_cast<uint4*>(op + j * 8) = u;
      }
      // natural-log log-sum-exp of the scaled scores (saved for the backward)
      p.lse[((size_t)n * p.heads + h) * T + q] = (m2 + log2f(l)) * 0.6931471805599453f;
    }
    tc_fence_before();
    __syncthreads();     // all TMEM reads of this block are done before the next block's MMAs overwrite S / O
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
struct AttnTcBwdParams {
  CUtensorMap map_qkv;   // qkv as [N*T][3C], box {64, 128}
  CUtensorMap map_do;    // d_out as [N*T][C], box {64, 128}
  int T, heads;
  const bf16* out;
  const bf16* dout;
  const float* lse;
  bf16* dqkv;
};

__global__ void __launch_bounds__(ATC_THREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ AttnTcBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int T = p.T, h = blockIdx.x, n = blockIdx.y;
  const int nb = (T + 127) / 128;          // 128-row blocks of queries and of keys (rows >= T are masked)
  const int rows = nb * 128;
  uint8_t* Qs = smem;                      // [rows][128 B]
  uint8_t* Ks = Qs + rows * 128;
  uint8_t* Vs = Ks + rows * 128;
  uint8_t* dOs = Vs + rows * 128;
  uint8_t* Ps = dOs + rows * 128;          // [2 panels of 64 keys][128 q][128 B]
  uint8_t* dSs = Ps + 32768;
  uint64_t* load_bar = reinterpret_cast<uint64_t*>(dSs + 32768);
  uint64_t* mma_bar = load_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&p.map_qkv);
    tma_prefetch_desc(&p.map_do);
    mbar_init(load_bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // columns: S [0,128)  dP [128,256)  dV [256,320)  dK [320,384)  dQ of query block 0 / 1 [384,448) / [448,512)
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320, tdQ = tmem_base + 384;

  const int C = p.heads * 64, C3 = 3 * C;
  const int col0 = h * 192, row0 = n * T;
  if (tid == 0) {
    mbar_arrive_expect_tx(load_bar, (uint32_t)(4 * rows * 128));
    for (int b = 0; b < nb; ++b) {
      tma_load_2d(Qs + b * 16384, &p.map_qkv, load_bar, col0, row0 + b * 128);
      tma_load_2d(Ks + b * 16384, &p.map_qkv, load_bar, col0 + 64, row0 + b * 128);
      tma_load_2d(Vs + b * 16384, &p.map_qkv, load_bar, col0 + 128, row0 + b * 128);
      tma_load_2d(dOs + b * 16384, &p.map_do, load_bar, h * 64, row0 + b * 128);
    }
  }
  // D[q] = sum_d dO[q,d] O[q,d] and the log2-scaled lse of this thread's query row in every query block
  const float LOG2E = 1.4426950408889634f;
  const float scale2 = 0.125f * LOG2E;
  float Dq[2] = {0.f, 0.f}, Lq[2] = {0.f, 0.f};
  for (int qb = 0; qb < nb; ++qb) {
    const int q = qb * 128 + tid;
    if (q < T) {
      const uint4* o4 = reinterpret_cast<const uint4*>(p.out + ((size_t)(row0 + q)) * C + h * 64);
      const uint4* d4 = reinterpret_cast<const uint4*>(p.dout + ((size_t)(row0 + q)) * C + h * 64);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 a = __ldg(o4 + j), b = __ldg(d4 + j);
        const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
        const float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
        acc += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
      }
      Dq[qb] = acc;
      Lq[qb] = p.lse[((size_t)n * p.heads + h) * T + q] * LOG2E;
    }
  }
  mbar_wait(load_bar, 0);

  const uint32_t idesc_kk = umma_idesc_bf16(128, 128);                      // S, dP: both operands K-major
  const uint32_t idesc_tn = umma_idesc_bf16_major(128, 64, 1, 1);           // dV, dK: A = P^T / dS^T (MN-major), B MN-major
  const uint32_t idesc_kn = umma_idesc_bf16_major(128, 64, 0, 1);           // dQ: A = dS K-major, B = K MN-major
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
  const int r = tid;
  const uint32_t swz = (uint32_t)(r & 7);
  uint32_t mma_phase = 0;

  auto issue_scores = [&](int kb, int qb) {
    const uint64_t qd = umma_desc_sw128(smem_u32(Qs + qb * 16384)), kd = umma_desc_sw128(smem_u32(Ks + kb * 16384));
    const uint64_t od = umma_desc_sw128(smem_u32(dOs + qb * 16384)), vd = umma_desc_sw128(smem_u32(Vs + kb * 16384));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tdP, od + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
    umma_commit(mma_bar);
  };

  if (tid == 0) issue_scores(0, 0);
  for (int kb = 0; kb < nb; ++kb) {
    for (int qb = 0; qb < nb; ++qb) {
      // S and dP of (kb, qb) are complete, and so is every earlier MMA (the previous iteration's readers of Ps / dSs)
      mbar_wait(mma_bar, mma_phase);
      mma_phase ^= 1;
      tc_fence_after();
      const bool q_ok = qb * 128 + r < T;
      for (int c = 0; c < 128; c += 32) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tS + lane_base + (uint32_t)c, sv);
        tmem_ld_32x32(tdP + lane_base + (uint32_t)c, dv);
        tmem_ld_wait();
        uint8_t* prow = Ps + (c >> 6) * 16384 + r * 128;
        uint8_t* srow = dSs + (c >> 6) * 16384 + r * 128;
        const int j0 = (c & 32) ? 4 : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float pe[8], de[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = q_ok && (kb * 128 + c + j * 8 + i < T);
            const float pr = exp2f(fmaf(__uint_as_float(sv[j * 8 + i]), scale2, -Lq[qb]));
            pe[i] = ok ? pr : 0.f;
            de[i] = ok ? pr * (__uint_as_float(dv[j * 8 + i]) - Dq[qb]) : 0.f;
          }
          uint4 u, w;
          u.x = pack_bf16(pe[0], pe[1]); u.y = pack_bf16(pe[2], pe[3]); u.z = pack_bf16(pe[4], pe[5]); u.w = pack_bf16(pe[6], pe[7]);
          w.x = pack_bf16(de[0], de[1]); w.y = pack_bf16(de[2], de[3]); w.z = pack_bf16(de[4], de[5]); w.w = pack_bf16(de[6], de[7]);
          const uint32_t off = (((uint32_t)(j0 + j) ^ swz) << 4);
          *reinterpret_cast<uint4*>(prow + off) = u;
          *reinterpret_cast<uint4*>(srow + off) = w;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      const bool last_q = (qb == nb - 1);
      if (tid == 0) {
        // dV_kb += P^T dO_qb, dK_kb += dS^T Q_qb: M = 128 keys (two 64-key panels, 16 KB apart), N = 64, K = 128 queries
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t pa = umma_desc_sw128_ls(smem_u32(Ps + kk * 2048), 16384, 1024);
          const uint64_t sa = umma_desc_sw128_ls(smem_u32(dSs + kk * 2048), 16384, 1024);
          const uint64_t ob = umma_desc_sw128_ls(smem_u32(dOs + qb * 16384 + kk * 2048), 16, 1024);
          const uint64_t qbd = umma_desc_sw128_ls(smem_u32(Qs + qb * 16384 + kk * 2048), 16, 1024);
          umma_bf16_ss(tdV, pa, ob, idesc_tn, (qb | kk) ? 1u : 0u);
          umma_bf16_ss(tdK, sa, qbd, idesc_tn, (qb | kk) ? 1u : 0u);
        }
        // dQ_qb += dS K_kb: M = 128 queries, N = 64, K = 128 keys
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t sa = umma_desc_sw128(smem_u32(dSs + (kk >> 2) * 16384)) + (uint64_t)(2 * (kk & 3));
          const uint64_t kbd = umma_desc_sw128_ls(smem_u32(Ks + kb * 16384 + kk * 2048), 16, 1024);
          umma_bf16_ss(tdQ + (uint32_t)(qb * 64), sa, kbd, idesc_kn, (kb | kk) ? 1u : 0u);
        }
        if (!last_q) issue_scores(kb, qb + 1);
        else umma_commit(mma_bar);              // dV_kb / dK_kb complete
      }
      if (last_q) {
        mbar_wait(mma_bar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        // this thread's key row of dK and dV
        const int key = kb * 128 + r;
        uint32_t a0[32], a1[32], b0[32], b1[32];
        tmem_ld_32x32(tdK + lane_base, a0);
        tmem_ld_32x32(tdK + lane_base + 32u, a1);
        tmem_ld_32x32(tdV + lane_base, b0);
        tmem_ld_32x32(tdV + lane_base + 32u, b1);
        tmem_ld_wait();
        if (key < T) {
          bf16* dk = p.dqkv + ((size_t)(row0 + key)) * C3 + col0 + 64;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 u, w;
            const uint32_t* ka = j < 4 ? a0 + j * 8 : a1 + (j - 4) * 8;
            const uint32_t* va = j < 4 ? b0 + j * 8 : b1 + (j - 4) * 8;
            u.x = pack_bf16(__uint_as_float(ka[0]) * 0.125f, __uint_as_float(ka[1]) * 0.125f);
            u.y = pack_bf16(__uint_as_float(ka[2]) * 0.125f, __uint_as_float(ka[3]) * 0.125f);
            u.z = pack_bf16(__uint_as_float(ka[4]) * 0.125f, __uint_as_float(ka[5]) * 0.125f);
            u.w = pack_bf16(__uint_as_float(ka[6]) * 0.125f, __uint_as_float(ka[7]) * 0.125f);
            w.x = pack_bf16(__uint_as_float(va[0]), __uint_as_float(va[1]));
            w.y = pack_bf16(__uint_as_float(va[2]), __uint_as_float(va[3]));
            w.z = pack_bf16(__uint_as_float(va[4]), __uint_as_float(va[5]));
            w.w = pack_bf16(__uint_as_float(va[6]), __uint_as_float(va[7]));
            *reinterpret_cast<uint4*>(dk + j * 8) = u;
            *reinterpret_cast<uint4*>(dk + 64 + j * 8) = w;
          }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0 && kb + 1 < nb) issue_scores(kb + 1, 0);
      }
    }
  }
  // dQ (every MMA has completed: the last wait above followed the final commit)
  for (int qb = 0; qb < nb; ++qb) {
    const int q = qb * 128 + r;
    uint32_t a0[32], a1[32];
    tmem_ld_32x32(tdQ + (uint32_t)(qb * 64) + lane_base, a0);
    tmem_ld_32x32(tdQ + (uint32_t)(qb * 64) + lane_base + 32u, a1);
    tmem_ld_wait();
    if (q < T) {
      bf16* dq = p.dqkv + ((size_t)(row0 + q)) * C3 + col0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t* qa = j < 4 ? a0 + j * 8 : a1 + (j - 4) * 8;
        uint4 u;
        u.x = pack_bf16(__uint_as_float(qa[0]) * 0.125f, __uint_as_float(qa[1]) * 0.125f);
        u.y = pack_bf16(__uint_as_float(qa[2]) * 0.125f, __uint_as_float(qa[3]) * 0.125f);
        u.z = pack_bf16(__uint_as_float(qa[4]) * 0.125f, __uint_as_float(qa[5]) * 0.125f);
        u.w = pack_bf16(__uint_as_float(qa[6]) * 0.125f, __uint_as_float(qa[7]) * 0.125f);
        *reinterpret_cast<uint4*>(dq + j * 8) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int launch_attention_bwd_tc(const bf16* qkv, const bf16* out, const bf16* d_out, const float* lse, int N, int T, int heads, bf16* dqkv,
                            cudaStream_t s) {
  AttnTcBwdParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t rows = (uint64_t)N * T;
  int rc = encode_tmap_bf16_2d(&p.map_qkv, qkv, (uint64_t)heads * 192, rows, 64, 128);
  if (rc != KDIP_OK) return rc;
  rc = encode_tmap_bf16_2d(&p.map_do, d_out, (uint64_t)heads * 64, rows, 64, 128);
  if (rc != KDIP_OK) return rc;
  p.T = T; p.heads = heads; p.out = out; p.dout = d_out; p.lse = lse; p.dqkv = dqkv;
  const int rows_pad = ((T + 127) / 128) * 128;
  const size_t smem = (size_t)4 * rows_pad * 128 + 65536 + 64 + 1024;
  static bool once = false;
  if (!once) {
    KDIP_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    once = true;
  }
  attn_bwd_tc_kernel<<<dim3(heads, N), ATC_THREADS, smem, s>>>(p);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

static size_t attn_fwd_tc_smem(int T) {
  const int q_rows = ((T + 127) / 128) * 128;
  return (size_t)(q_rows + 2 * T) * 128 + (size_t)(T / 64) * 16384 + 64 + 1024;
}

bool attention_tc_supported(int T, int ch) {
  if (getenv("KDIP_ATTN_TC") && atoi(getenv("KDIP_ATTN_TC")) == 0) return false;
  return ch == 64 && (T == 64 || T == 128 || T == 256);
}

int launch_attention_fwd_tc(const bf16* qkv, int N, int T, int heads, bf16* out, float* lse, cudaStream_t s) {
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t cols = (uint64_t)heads * 192, rows = (uint64_t)N * T;
  int rc = encode_tmap_bf16_2d(&p.map_q, qkv, cols, rows, 64, 128);
  if (rc != KDIP_OK) return rc;
  rc = encode_tmap_bf16_2d(&p.map_kv, qkv, cols, rows, 64, (uint32_t)T);
  if (rc != KDIP_OK) return rc;
  p.T = T; p.heads = heads; p.out = out; p.lse = lse;
  const size_t smem = attn_fwd_tc_smem(T);
  static bool once = false;
  if (!once) {
    KDIP_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    once = true;
  }
  attn_fwd_tc_kernel<<<dim3(heads, N), ATC_THREADS, smem, s>>>(p);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

}  // namespace kdip
