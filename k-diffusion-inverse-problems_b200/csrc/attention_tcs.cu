// QKVAttentionLegacy (guided_diffusion/unet.py:339-356) on the 5th-gen tensor cores for LONG sequences: head width 64, T a
// multiple of 128 with T >= 384 (the 32x32-token attention level of the ImageNet ADM UNet, T = 1024).  attention_tc.cu keeps one
// (image, head) entirely in shared memory / TMEM, which stops at T = 256; here one 128-row block stays resident and the other
// side is STREAMED through a TMA ring, 128 threads = the 128 TMEM lanes = one query (or key) row each.
//
//   forward, CTA = (query block, head, image):  pass 1  S = Q K_kb^T for every key block -> row maximum
//                                               pass 2  S again, P = exp2(S/8 log2e - max) -> bf16 panels, O += P V_kb (TMEM)
//       (two passes instead of an online rescale of the TMEM accumulator: the QK^T MMAs are 1/3 of the exp cost)
//   backward (autograd sites condition/condition.py:136,146,155,172,269), three launches:
//     attn_bwd_d_kernel        D[q] = sum_d dO[q,d] O[q,d], parked as fp32 in the (not yet written) dQ slot of row q
//     attn_bwd_dkv_tcs_kernel  CTA = (key block, head, image), K/V resident, (Q, dO) blocks streamed:
//                              S = Q K^T, dP = dO V^T -> P = exp(S/8 - lse), dS = P o (dP - D) -> dV += P^T dO, dK += dS^T Q
//     attn_bwd_dq_tcs_kernel   CTA = (query block, head, image), Q/dO resident, (K, V) blocks streamed: dQ += dS K
//   No atomics, no scratch buffers, deterministic summation order.  Operand layouts / descriptors as in attention_tc.cu.
#include <stdlib.h>

#include "attention_tc.cuh"

namespace kdip {

static constexpr int TCS_THREADS = 128;
static constexpr uint32_t TILE = 16384;   // [128 rows][64 bf16]: rows of 128 bytes in the canonical 128B-swizzle layout

struct AttnTcsParams {
  CUtensorMap map_qkv;   // qkv as [N*T][3C], box {64, 128}
  CUtensorMap map_do;    // d_out as [N*T][C], box {64, 128} (backward only)
  int T, heads;
  bf16* out;             // forward: attention output [N*T][C]
  float* lse;            // natural-log log-sum-exp of the scaled scores [N][heads][T] (forward writes, backward reads)
  const bf16* o;         // backward: the forward output
  const bf16* dout;
  bf16* dqkv;
};

__device__ __forceinline__ void store_row64(bf16* dst, const uint32_t (&a0)[32], const uint32_t (&a1)[32], float mul) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t* a = j < 4 ? a0 + j * 8 : a1 + (j - 4) * 8;
    uint4 u;
    u.x = pack_bf16(__uint_as_float(a[0]) * mul, __uint_as_float(a[1]) * mul);
    u.y = pack_bf16(__uint_as_float(a[2]) * mul, __uint_as_float(a[3]) * mul);
    u.z = pack_bf16(__uint_as_float(a[4]) * mul, __uint_as_float(a[5]) * mul);
    u.w = pack_bf16(__uint_as_float(a[6]) * mul, __uint_as_float(a[7]) * mul);
    *reinterpret_cast<uint4*>(dst + j * 8) = u;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TCS_THREADS, 2) attn_fwd_tcs_kernel(const __grid_constant__ AttnTcsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int T = p.T, qb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int nkb = T >> 7;
  uint8_t* Qs = smem;                 // resident query block
  uint8_t* Ks = Qs + TILE;            // 2 stages
  uint8_t* Vs = Ks + 2 * TILE;        // 1 stage (free again when the next S completes, see below)
  uint8_t* Ps = Vs + TILE;            // 2 panels of 64 keys: [128 q][128 B]
  uint64_t* q_bar = reinterpret_cast<uint64_t*>(Ps + 2 * TILE);
  uint64_t* k_bar = q_bar + 1;        // [2]
  uint64_t* v_bar = q_bar + 3;
  uint64_t* mma_bar = q_bar + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(q_bar + 5);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&p.map_qkv);
    for (int i = 0; i < 5; ++i) mbar_init(q_bar + i, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_ptr_smem, 256); tmem_relinquish(); }   // S: columns [0,128), O: [128,192); 2 CTAs per SM
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128u;

  const int col0 = h * 192, row0 = n * T;
  if (tid == 0) {
    mbar_arrive_expect_tx(q_bar, TILE);
    tma_load_2d(Qs, &p.map_qkv, q_bar, col0, row0 + qb * 128);
    mbar_arrive_expect_tx(k_bar, TILE);
    tma_load_2d(Ks, &p.map_qkv, k_bar, col0 + 64, row0);
    mbar_wait(q_bar, 0);
  }

  const float scale = 0.125f * 1.4426950408889634f;   // (ch^-1/4)^2 = 1/8, folded with log2(e) for exp2f
  const uint32_t idesc_s = umma_idesc_bf16(128, 128);
  const uint32_t idesc_o = umma_idesc_bf16_major(128, 64, 0, 1);
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
  const int r = tid;
  const uint32_t swz = (uint32_t)(r & 7);
  uint32_t mph = 0, kph = 0, vph = 0;
  float mx = -INFINITY, m2 = 0.f, l = 0.f;
  const int total = 2 * nkb;

  for (int it = 0; it < total; ++it) {
    const bool pass2 = it >= nkb;
    const int kb = pass2 ? it - nkb : it;
    const int st = it & 1;
    if (tid == 0) {
      if (it + 1 < total) {     // stage st^1 was last read by S(it-1), whose completion this thread waited for
        const int nk = (it + 1 < nkb) ? it + 1 : it + 1 - nkb;
        mbar_arrive_expect_tx(k_bar + (st ^ 1), TILE);
        tma_load_2d(Ks + (st ^ 1) * TILE, &p.map_qkv, k_bar + (st ^ 1), col0 + 64, row0 + nk * 128);
      }
      mbar_wait(k_bar + st, (kph >> st) & 1u);
      kph ^= 1u << st;
      tc_fence_after();
      const uint64_t qd = umma_desc_sw128(smem_u32(Qs)), kd = umma_desc_sw128(smem_u32(Ks + st * TILE));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_S, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
      umma_commit(mma_bar);
    }
    // S(it) complete, and with it every earlier MMA: the P V product of the previous iteration (readers of Ps and Vs)
    mbar_wait(mma_bar, mph);
    mph ^= 1;
    tc_fence_after();
    if (!pass2) {
      for (int c = 0; c < 128; c += 64) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(tmem_S + lane_base + (uint32_t)c, v0);
        tmem_ld_32x32(tmem_S + lane_base + (uint32_t)c + 32u, v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
      }
      if (it == nkb - 1) m2 = mx * scale;   // in log2 units
    } else {
      if (tid == 0) {
        mbar_arrive_expect_tx(v_bar, TILE);
        tma_load_2d(Vs, &p.map_qkv, v_bar, col0 + 128, row0 + kb * 128);
      }
      for (int c = 0; c < 128; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_S + lane_base + (uint32_t)c, v);
        tmem_ld_wait();
        uint8_t* prow = Ps + (c >> 6) * TILE + r * 128;
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
    }
    tc_fence_before();
    __syncthreads();          // every lane has read S(it) (and written its P row) before the next MMAs are issued
    tc_fence_after();
    if (pass2 && tid == 0) {
      mbar_wait(v_bar, vph);
      vph ^= 1;
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t pd = umma_desc_sw128(smem_u32(Ps + (kk >> 2) * TILE)) + (uint64_t)(2 * (kk & 3));
        const uint64_t vd = umma_desc_sw128_ls(smem_u32(Vs + kk * 2048), 16, 1024);
        umma_bf16_ss(tmem_O, pd, vd, idesc_o, (kb | kk) ? 1u : 0u);
      }
      if (it == total - 1) umma_commit(mma_bar);
    }
  }
  mbar_wait(mma_bar, mph);
  tc_fence_after();
  {
    const int q = qb * 128 + r;
    const int C = p.heads * 64;
    uint32_t o0[32], o1[32];
    tmem_ld_32x32(tmem_O + lane_base, o0);
    tmem_ld_32x32(tmem_O + lane_base + 32u, o1);
    tmem_ld_wait();
    store_row64(p.out + ((size_t)(row0 + q)) * C + h * 64, o0, o1, 1.f / l);
    p.lse[((size_t)n * p.heads + h) * T + q] = (m2 + log2f(l)) * 0.6931471805599453f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
// D[q] for every (row, head): one thread reads the 128-byte rows of O and dO; fp32 result parked in the first 4 bytes of the
// row's dQ slot (read by both backward kernels before attn_bwd_dq_tcs_kernel overwrites it with dQ at its very end).
__global__ void __launch_bounds__(256) attn_bwd_d_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, long long rows,
                                                         int heads, bf16* __restrict__ dqkv) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * heads) return;
  const long long row = idx / heads;
  const int h = (int)(idx - row * heads);
  const int C = heads * 64;
  const uint4* o4 = reinterpret_cast<const uint4*>(out + row * C + h * 64);
  const uint4* d4 = reinterpret_cast<const uint4*>(dout + row * C + h * 64);
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 a = __ldg(o4 + j), b = __ldg(d4 + j);
    const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
    const float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
    acc += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
  }
  *reinterpret_cast<float*>(dqkv + row * (3 * C) + h * 192) = acc;
}

// dK, dV of one key block: K / V resident, (Q, dO) query blocks streamed through 3 stages.
__global__ void __launch_bounds__(TCS_THREADS, 1) attn_bwd_dkv_tcs_kernel(const __grid_constant__ AttnTcsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int T = p.T, kb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int nqb = T >> 7;
  uint8_t* Ks = smem;
  uint8_t* Vs = Ks + TILE;
  uint8_t* St = Vs + TILE;            // 3 stages of {Q block, dO block}
  uint8_t* Ps = St + 6 * TILE;        // [2 panels of 64 keys][128 q][128 B]
  uint8_t* dSs = Ps + 2 * TILE;
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(dSs + 2 * TILE);
  uint64_t* st_bar = res_bar + 1;     // [3]
  uint64_t* mma_bar = res_bar + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 5);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&p.map_qkv);
    tma_prefetch_desc(&p.map_do);
    for (int i = 0; i < 5; ++i) mbar_init(res_bar + i, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // columns: S [0,128)  dP [128,256)  dV [256,320)  dK [320,384)
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320;

  const int C = p.heads * 64, C3 = 3 * C;
  const int col0 = h * 192, row0 = n * T;
  auto load_stage = [&](int qb) {
    const int s = qb % 3;
    mbar_arrive_expect_tx(st_bar + s, 2 * TILE);
    tma_load_2d(St + s * 2 * TILE, &p.map_qkv, st_bar + s, col0, row0 + qb * 128);
    tma_load_2d(St + s * 2 * TILE + TILE, &p.map_do, st_bar + s, h * 64, row0 + qb * 128);
  };
  const uint32_t idesc_kk = umma_idesc_bf16(128, 128);                      // S, dP: both operands K-major
  const uint32_t idesc_tn = umma_idesc_bf16_major(128, 64, 1, 1);           // dV, dK: A = P^T / dS^T (MN-major), B MN-major
  auto issue_scores = [&](int qb) {       // waits for the stage of query block qb, then S = Q K^T, dP = dO V^T
    const int s = qb % 3;
    mbar_wait(st_bar + s, (uint32_t)((qb / 3) & 1));
    tc_fence_after();
    const uint64_t qd = umma_desc_sw128(smem_u32(St + s * 2 * TILE)), kd = umma_desc_sw128(smem_u32(Ks));
    const uint64_t od = umma_desc_sw128(smem_u32(St + s * 2 * TILE + TILE)), vd = umma_desc_sw128(smem_u32(Vs));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tdP, od + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
    umma_commit(mma_bar);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(res_bar, 2 * TILE);
    tma_load_2d(Ks, &p.map_qkv, res_bar, col0 + 64, row0 + kb * 128);
    tma_load_2d(Vs, &p.map_qkv, res_bar, col0 + 128, row0 + kb * 128);
    load_stage(0);
    if (nqb > 1) load_stage(1);
    mbar_wait(res_bar, 0);
    issue_scores(0);
  }

  const float LOG2E = 1.4426950408889634f;
  const float scale2 = 0.125f * LOG2E;
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
  const int r = tid;
  const uint32_t swz = (uint32_t)(r & 7);
  uint32_t mph = 0;

  for (int qb = 0; qb < nqb; ++qb) {
    const int s = qb % 3;
    const int q = qb * 128 + r;       // this thread's query row of the streamed block
    const float Lq = p.lse[((size_t)n * p.heads + h) * T + q] * LOG2E;
    const float Dq = *reinterpret_cast<const float*>(p.dqkv + ((size_t)(row0 + q)) * C3 + col0);
    // S and dP of qb are complete, and so is every earlier MMA (dV / dK of qb-1: readers of Ps, dSs and of stage (qb-1)%3)
    mbar_wait(mma_bar, mph);
    mph ^= 1;
    tc_fence_after();
    if (tid == 0 && qb + 2 < nqb) load_stage(qb + 2);     // into stage (qb-1)%3
    for (int c = 0; c < 128; c += 32) {
      uint32_t sv[32], dv[32];
      tmem_ld_32x32(tS + lane_base + (uint32_t)c, sv);
      tmem_ld_32x32(tdP + lane_base + (uint32_t)c, dv);
      tmem_ld_wait();
      uint8_t* prow = Ps + (c >> 6) * TILE + r * 128;
      uint8_t* srow = dSs + (c >> 6) * TILE + r * 128;
      const int j0 = (c & 32) ? 4 : 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pe[8], de[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pe[i] = exp2f(fmaf(__uint_as_float(sv[j * 8 + i]), scale2, -Lq));
          de[i] = pe[i] * (__uint_as_float(dv[j * 8 + i]) - Dq);
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
    if (tid == 0) {
      // dV += P^T dO_qb, dK += dS^T Q_qb: M = 128 keys (two 64-key panels, 16 KB apart), N = 64, K = 128 queries
      const uint32_t qs = smem_u32(St + s * 2 * TILE), os = qs + TILE;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t pa = umma_desc_sw128_ls(smem_u32(Ps + kk * 2048), TILE, 1024);
        const uint64_t sa = umma_desc_sw128_ls(smem_u32(dSs + kk * 2048), TILE, 1024);
        const uint64_t ob = umma_desc_sw128_ls(os + kk * 2048, 16, 1024);
        const uint64_t qbd = umma_desc_sw128_ls(qs + kk * 2048, 16, 1024);
        umma_bf16_ss(tdV, pa, ob, idesc_tn, (qb | kk) ? 1u : 0u);
        umma_bf16_ss(tdK, sa, qbd, idesc_tn, (qb | kk) ? 1u : 0u);
      }
      if (qb + 1 < nqb) issue_scores(qb + 1);    // commits: covers the products above as well
      else umma_commit(mma_bar);
    }
  }
  mbar_wait(mma_bar, mph);
  tc_fence_after();
  {
    const int key = kb * 128 + r;
    uint32_t a0[32], a1[32], b0[32], b1[32];
    tmem_ld_32x32(tdK + lane_base, a0);
    tmem_ld_32x32(tdK + lane_base + 32u, a1);
    tmem_ld_32x32(tdV + lane_base, b0);
    tmem_ld_32x32(tdV + lane_base + 32u, b1);
    tmem_ld_wait();
    bf16* dk = p.dqkv + ((size_t)(row0 + key)) * C3 + col0 + 64;
    store_row64(dk, a0, a1, 0.125f);
    store_row64(dk + 64, b0, b1, 1.f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// dQ of one query block: Q / dO resident, (K, V) key blocks streamed through 3 stages.
__global__ void __launch_bounds__(TCS_THREADS, 1) attn_bwd_dq_tcs_kernel(const __grid_constant__ AttnTcsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int T = p.T, qb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int nkb = T >> 7;
  uint8_t* Qs = smem;
  uint8_t* dOs = Qs + TILE;
  uint8_t* St = dOs + TILE;           // 3 stages of {K block, V block}
  uint8_t* dSs = St + 6 * TILE;       // [2 panels of 64 keys][128 q][128 B]
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(dSs + 2 * TILE);
  uint64_t* st_bar = res_bar + 1;     // [3]
  uint64_t* mma_bar = res_bar + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 5);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&p.map_qkv);
    tma_prefetch_desc(&p.map_do);
    for (int i = 0; i < 5; ++i) mbar_init(res_bar + i, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdQ = tmem_base + 256;   // S [0,128)  dP [128,256)  dQ [256,320)

  const int C = p.heads * 64, C3 = 3 * C;
  const int col0 = h * 192, row0 = n * T;
  auto load_stage = [&](int kb) {
    const int s = kb % 3;
    mbar_arrive_expect_tx(st_bar + s, 2 * TILE);
    tma_load_2d(St + s * 2 * TILE, &p.map_qkv, st_bar + s, col0 + 64, row0 + kb * 128);
    tma_load_2d(St + s * 2 * TILE + TILE, &p.map_qkv, st_bar + s, col0 + 128, row0 + kb * 128);
  };
  const uint32_t idesc_kk = umma_idesc_bf16(128, 128);
  const uint32_t idesc_kn = umma_idesc_bf16_major(128, 64, 0, 1);           // dQ: A = dS K-major, B = K MN-major
  auto issue_scores = [&](int kb) {
    const int s = kb % 3;
    mbar_wait(st_bar + s, (uint32_t)((kb / 3) & 1));
    tc_fence_after();
    const uint64_t qd = umma_desc_sw128(smem_u32(Qs)), kd = umma_desc_sw128(smem_u32(St + s * 2 * TILE));
    const uint64_t od = umma_desc_sw128(smem_u32(dOs)), vd = umma_desc_sw128(smem_u32(St + s * 2 * TILE + TILE));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tdP, od + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
    umma_commit(mma_bar);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(res_bar, 2 * TILE);
    tma_load_2d(Qs, &p.map_qkv, res_bar, col0, row0 + qb * 128);
    tma_load_2d(dOs, &p.map_do, res_bar, h * 64, row0 + qb * 128);
    load_stage(0);
    if (nkb > 1) load_stage(1);
    mbar_wait(res_bar, 0);
    issue_scores(0);
  }

  const float LOG2E = 1.4426950408889634f;
  const float scale2 = 0.125f * LOG2E;
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
  const int r = tid;
  const uint32_t swz = (uint32_t)(r & 7);
  const int q = qb * 128 + r;
  bf16* dq = p.dqkv + ((size_t)(row0 + q)) * C3 + col0;
  const float Lq = p.lse[((size_t)n * p.heads + h) * T + q] * LOG2E;
  const float Dq = *reinterpret_cast<const float*>(dq);    // parked by attn_bwd_d_kernel; overwritten by this thread at the end
  uint32_t mph = 0;

  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % 3;
    mbar_wait(mma_bar, mph);
    mph ^= 1;
    tc_fence_after();
    if (tid == 0 && kb + 2 < nkb) load_stage(kb + 2);     // into stage (kb-1)%3: its readers completed with the wait above
    for (int c = 0; c < 128; c += 32) {
      uint32_t sv[32], dv[32];
      tmem_ld_32x32(tS + lane_base + (uint32_t)c, sv);
      tmem_ld_32x32(tdP + lane_base + (uint32_t)c, dv);
      tmem_ld_wait();
      uint8_t* srow = dSs + (c >> 6) * TILE + r * 128;
      const int j0 = (c & 32) ? 4 : 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float de[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          de[i] = exp2f(fmaf(__uint_as_float(sv[j * 8 + i]), scale2, -Lq)) * (__uint_as_float(dv[j * 8 + i]) - Dq);
        uint4 w;
        w.x = pack_bf16(de[0], de[1]); w.y = pack_bf16(de[2], de[3]); w.z = pack_bf16(de[4], de[5]); w.w = pack_bf16(de[6], de[7]);
        *reinterpret_cast<uint4*>(srow + (((uint32_t)(j0 + j) ^ swz) << 4)) = w;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      // dQ += dS K_kb: M = 128 queries, N = 64, K = 128 keys
      const uint32_t ks = smem_u32(St + s * 2 * TILE);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t sa = umma_desc_sw128(smem_u32(dSs + (kk >> 2) * TILE)) + (uint64_t)(2 * (kk & 3));
        const uint64_t kbd = umma_desc_sw128_ls(ks + kk * 2048, 16, 1024);
        umma_bf16_ss(tdQ, sa, kbd, idesc_kn, (kb | kk) ? 1u : 0u);
      }
      if (kb + 1 < nkb) issue_scores(kb + 1);
      else umma_commit(mma_bar);
    }
  }
  mbar_wait(mma_bar, mph);
  tc_fence_after();
  {
    uint32_t a0[32], a1[32];
    tmem_ld_32x32(tdQ + lane_base, a0);
    tmem_ld_32x32(tdQ + lane_base + 32u, a1);
    tmem_ld_wait();
    store_row64(dq, a0, a1, 0.125f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------------
bool attention_tcs_supported(int T, int ch) {
  if (getenv("KDIP_ATTN_TC") && atoi(getenv("KDIP_ATTN_TC")) == 0) return false;
  if (getenv("KDIP_ATTN_TCS") && atoi(getenv("KDIP_ATTN_TCS")) == 0) return false;
  return ch == 64 && T >= 384 && T % 128 == 0;
}

static constexpr size_t TCS_SMEM_FWD = 6 * TILE + 64 + 1024;
static constexpr size_t TCS_SMEM_DKV = 12 * TILE + 64 + 1024;
static constexpr size_t TCS_SMEM_DQ = 10 * TILE + 64 + 1024;

int launch_attention_fwd_tcs(const bf16* qkv, int N, int T, int heads, bf16* out, float* lse, cudaStream_t s) {
  AttnTcsParams p;
  memset(&p, 0, sizeof(p));
  int rc = encode_tmap_bf16_2d(&p.map_qkv, qkv, (uint64_t)heads * 192, (uint64_t)N * T, 64, 128);
  if (rc != KDIP_OK) return rc;
  p.T = T; p.heads = heads; p.out = out; p.lse = lse;
  static bool once = false;
  if (!once) {
    KDIP_CUDA(cudaFuncSetAttribute(attn_fwd_tcs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCS_SMEM_FWD));
    once = true;
  }
  attn_fwd_tcs_kernel<<<dim3(T / 128, heads, N), TCS_THREADS, TCS_SMEM_FWD, s>>>(p);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

int launch_attention_bwd_tcs(const bf16* qkv, const bf16* out, const bf16* d_out, const float* lse, int N, int T, int heads,
                             bf16* dqkv, cudaStream_t s) {
  AttnTcsParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t rows = (uint64_t)N * T;
  int rc = encode_tmap_bf16_2d(&p.map_qkv, qkv, (uint64_t)heads * 192, rows, 64, 128);
  if (rc != KDIP_OK) return rc;
  rc = encode_tmap_bf16_2d(&p.map_do, d_out, (uint64_t)heads * 64, rows, 64, 128);
  if (rc != KDIP_OK) return rc;
  p.T = T; p.heads = heads; p.o = out; p.dout = d_out; p.lse = const_cast<float*>(lse); p.dqkv = dqkv;
  static bool once = false;
  if (!once) {
    KDIP_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tcs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCS_SMEM_DKV));
    KDIP_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tcs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCS_SMEM_DQ));
    once = true;
  }
  const long long nd = (long long)rows * heads;
  attn_bwd_d_kernel<<<(unsigned)((nd + 255) / 256), 256, 0, s>>>(out, d_out, (long long)rows, heads, dqkv);
  KDIP_LAUNCH_CHECK();
  attn_bwd_dkv_tcs_kernel<<<dim3(T / 128, heads, N), TCS_THREADS, TCS_SMEM_DKV, s>>>(p);
  KDIP_LAUNCH_CHECK();
  attn_bwd_dq_tcs_kernel<<<dim3(T / 128, heads, N), TCS_THREADS, TCS_SMEM_DQ, s>>>(p);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

}  // namespace kdip
