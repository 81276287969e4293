// QKVAttentionLegacy forward and backward (guided_diffusion/unet.py:339-356), head width 64.
//   w = softmax_fp32((q*s)^T (k*s)), s = ch^-1/4 ;  a[b,c,t] = sum_s w[t,s] v[c,s]
// qkv is bf16 [N,T,3C] with the legacy per-head interleave: channel = head*3*ch + {0:q,1:k,2:v}*ch + c  (unet.py:349).
// Round-1 implementation: fp32 CUDA-core flash-style kernels (scores never leave shared memory; probabilities are
// recomputed from the saved log-sum-exp in the backward).  Attention is 0.1% of the UNet FLOPs;
// the tensor-core (tcgen05) version is a later-round item.
#include <stdlib.h>

#include "unet_kernels.cuh"

namespace kdip {

static constexpr int CH = 64;       // head channels
static constexpr int QB = 32;       // rows per block (queries in fwd/dQ, keys in dK/dV)
static constexpr int KC = 64;       // columns per chunk
static constexpr int LD = 65;       // padded leading dimension

// load a [rows x 64] bf16 tile (row stride = ld_elems) into fp32 smem with leading dimension ld_s, times scale
__device__ __forceinline__ void load_tile(const bf16* __restrict__ g, size_t ld_elems, int rows, float* __restrict__ sdst, int ld_s,
                                          float scale) {
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, v = i & 7;
    uint4 u = __ldg(reinterpret_cast<const uint4*>(g + (size_t)r * ld_elems + v * 8));
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    float* o = sdst + r * ld_s + v * 8;
    o[0] = a.x * scale; o[1] = a.y * scale; o[2] = b.x * scale; o[3] = b.y * scale;
    o[4] = c.x * scale; o[5] = c.y * scale; o[6] = d.x * scale; o[7] = d.y * scale;
  }
}

// acc[r][j] = sum_k A[4w+r][k] * B[lane + 32j][k]   (A: [QB][lda], B: [KC][LD])
__device__ __forceinline__ void rowdot(const float* __restrict__ A, int lda, const float* __restrict__ B, int w, int lane,
                                       float (&acc)[4][2]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = 0.f;
#pragma unroll 8
  for (int k = 0; k < CH; ++k) {
    const float b0 = B[lane * LD + k], b1 = B[(lane + 32) * LD + k];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float a = A[(4 * w + r) * lda + k];
      acc[r][0] += a * b0;
      acc[r][1] += a * b1;
    }
  }
}
// acc[r][j] += sum_s P[4w+r][s] * B[s][lane + 32j]   (P: [QB][LD], B: [KC][LD])
__device__ __forceinline__ void colacc(const float* __restrict__ P, const float* __restrict__ B, int w, int lane,
                                       float (&acc)[4][2]) {
#pragma unroll 8
  for (int s = 0; s < KC; ++s) {
    const float b0 = B[s * LD + lane], b1 = B[s * LD + lane + 32];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float p = P[(4 * w + r) * LD + s];
      acc[r][0] += p * b0;
      acc[r][1] += p * b1;
    }
  }
}

__global__ void __launch_bounds__(256) attn_fwd_kernel(const bf16* __restrict__ qkv, int T, int heads, bf16* __restrict__ out,
                                                       float* __restrict__ lse) {
  extern __shared__ float sm[];
  float* Qs = sm;                 // [QB][LD]
  float* Ks = Qs + QB * LD;       // [KC][LD]
  float* Vs = Ks + KC * LD;       // [KC][LD]
  float* Ps = Vs + KC * LD;       // [QB][LD]
  const int n = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * QB;
  const int C3 = heads * 3 * CH;
  const bf16* base = qkv + (size_t)n * T * C3 + h * 3 * CH;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale2 = rsqrtf((float)CH);   // (ch^-1/4)^2
  load_tile(base + (size_t)t0 * C3, C3, QB, Qs, LD, scale2);
  float m[4], l[4], o[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) { m[r] = -INFINITY; l[r] = 0.f; o[r][0] = o[r][1] = 0.f; }
  for (int s0 = 0; s0 < T; s0 += KC) {
    __syncthreads();
    load_tile(base + (size_t)s0 * C3 + CH, C3, KC, Ks, LD, 1.f);
    load_tile(base + (size_t)s0 * C3 + 2 * CH, C3, KC, Vs, LD, 1.f);
    __syncthreads();
    float sc[4][2];
    rowdot(Qs, LD, Ks, w, lane, sc);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float mx = fmaxf(sc[r][0], sc[r][1]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mn = fmaxf(m[r], mx);
      const float alpha = __expf(m[r] - mn);
      const float p0 = __expf(sc[r][0] - mn), p1 = __expf(sc[r][1] - mn);
      l[r] = l[r] * alpha + warp_sum(p0 + p1);
      m[r] = mn;
      o[r][0] *= alpha; o[r][1] *= alpha;
      Ps[(4 * w + r) * LD + lane] = p0;
      Ps[(4 * w + r) * LD + lane + 32] = p1;
    }
    __syncwarp();
    colacc(Ps, Vs, w, lane, o);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int t = t0 + 4 * w + r;
    const float inv = 1.f / l[r];
    bf16* op = out + ((size_t)n * T + t) * (heads * CH) + h * CH;
    op[lane] = __float2bfloat16(o[r][0] * inv);
    op[lane + 32] = __float2bfloat16(o[r][1] * inv);
    if (lane == 0) lse[((size_t)n * heads + h) * T + t] = m[r] + __logf(l[r]);
  }
}

// dQ: one block per 32 queries, loops over key chunks
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out,
                                                          const bf16* __restrict__ dout, const float* __restrict__ lse, int T,
                                                          int heads, bf16* __restrict__ dqkv) {
  extern __shared__ float sm[];
  float* Qs = sm;                  // [QB][LD] scaled q
  float* dOs = Qs + QB * LD;       // [QB][LD]
  float* Ks = dOs + QB * LD;       // [KC][LD]
  float* Vs = Ks + KC * LD;        // [KC][LD]
  float* Ps = Vs + KC * LD;        // [QB][LD]  (dS)
  const int n = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * QB;
  const int C3 = heads * 3 * CH, C = heads * CH;
  const bf16* base = qkv + (size_t)n * T * C3 + h * 3 * CH;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale2 = rsqrtf((float)CH);
  load_tile(base + (size_t)t0 * C3, C3, QB, Qs, LD, scale2);
  load_tile(dout + ((size_t)n * T + t0) * C + h * CH, C, QB, dOs, LD, 1.f);
  __syncthreads();
  float D[4], L[4], dq[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int t = t0 + 4 * w + r;
    const bf16* op = out + ((size_t)n * T + t) * C + h * CH;
    float d = dOs[(4 * w + r) * LD + lane] * __bfloat162float(op[lane]) +
              dOs[(4 * w + r) * LD + lane + 32] * __bfloat162float(op[lane + 32]);
    D[r] = warp_sum(d);
    L[r] = lse[((size_t)n * heads + h) * T + t];
    dq[r][0] = dq[r][1] = 0.f;
  }
  for (int s0 = 0; s0 < T; s0 += KC) {
    __syncthreads();
    load_tile(base + (size_t)s0 * C3 + CH, C3, KC, Ks, LD, 1.f);
    load_tile(base + (size_t)s0 * C3 + 2 * CH, C3, KC, Vs, LD, 1.f);
    __syncthreads();
    float sc[4][2], dp[4][2];
    rowdot(Qs, LD, Ks, w, lane, sc);
    rowdot(dOs, LD, Vs, w, lane, dp);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float p = __expf(sc[r][j] - L[r]);
        Ps[(4 * w + r) * LD + lane + 32 * j] = p * (dp[r][j] - D[r]);
      }
    }
    __syncwarp();
    colacc(Ps, Ks, w, lane, dq);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int t = t0 + 4 * w + r;
    bf16* dp_ = dqkv + ((size_t)n * T + t) * C3 + h * 3 * CH;
    dp_[lane] = __float2bfloat16(dq[r][0] * scale2);
    dp_[lane + 32] = __float2bfloat16(dq[r][1] * scale2);
  }
}

// dK, dV: one block per 32 keys, loops over query chunks
__global__ void __launch_bounds__(256) attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out,
                                                           const bf16* __restrict__ dout, const float* __restrict__ lse, int T,
                                                           int heads, bf16* __restrict__ dqkv) {
  extern __shared__ float sm[];
  float* Kb = sm;                  // [QB][LD] scaled k (rows = keys)
  float* Vb = Kb + QB * LD;        // [QB][LD]
  float* Qc = Vb + QB * LD;        // [KC][LD] q chunk (cols = queries)
  float* dOc = Qc + KC * LD;       // [KC][LD]
  float* PT = dOc + KC * LD;       // [QB][LD]
  float* dST = PT + QB * LD;       // [QB][LD]
  float* Lc = dST + QB * LD;       // [KC]
  float* Dc = Lc + KC;             // [KC]
  const int n = blockIdx.z, h = blockIdx.y, s0 = blockIdx.x * QB;
  const int C3 = heads * 3 * CH, C = heads * CH;
  const bf16* base = qkv + (size_t)n * T * C3 + h * 3 * CH;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale2 = rsqrtf((float)CH);
  load_tile(base + (size_t)s0 * C3 + CH, C3, QB, Kb, LD, scale2);
  load_tile(base + (size_t)s0 * C3 + 2 * CH, C3, QB, Vb, LD, 1.f);
  float dk[4][2], dv[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) dk[r][0] = dk[r][1] = dv[r][0] = dv[r][1] = 0.f;
  for (int t0 = 0; t0 < T; t0 += KC) {
    __syncthreads();
    load_tile(base + (size_t)t0 * C3, C3, KC, Qc, LD, 1.f);
    load_tile(dout + ((size_t)n * T + t0) * C + h * CH, C, KC, dOc, LD, 1.f);
    if (threadIdx.x < KC) Lc[threadIdx.x] = lse[((size_t)n * heads + h) * T + t0 + threadIdx.x];
    __syncthreads();
    // D[t] = sum_c dO[t][c] * O[t][c]: 4 threads per row
    {
      const int r = threadIdx.x >> 2, part = threadIdx.x & 3;
      const bf16* op = out + ((size_t)n * T + t0 + r) * C + h * CH + part * 16;
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) d += dOc[r * LD + part * 16 + c] * __bfloat162float(op[c]);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      if (part == 0) Dc[r] = d;
    }
    __syncthreads();
    float st[4][2], dpt[4][2];
    rowdot(Kb, LD, Qc, w, lane, st);     // S^T[s][t] (already scaled)
    rowdot(Vb, LD, dOc, w, lane, dpt);   // dP^T[s][t]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int tc = lane + 32 * j;
        const float p = __expf(st[r][j] - Lc[tc]);
        PT[(4 * w + r) * LD + tc] = p;
        dST[(4 * w + r) * LD + tc] = p * (dpt[r][j] - Dc[tc]);
      }
    }
    __syncwarp();
    colacc(PT, dOc, w, lane, dv);
    colacc(dST, Qc, w, lane, dk);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int s = s0 + 4 * w + r;
    bf16* dp_ = dqkv + ((size_t)n * T + s) * C3 + h * 3 * CH;
    dp_[CH + lane] = __float2bfloat16(dk[r][0] * scale2);
    dp_[CH + lane + 32] = __float2bfloat16(dk[r][1] * scale2);
    dp_[2 * CH + lane] = __float2bfloat16(dv[r][0]);
    dp_[2 * CH + lane + 32] = __float2bfloat16(dv[r][1]);
  }
}

static int set_smem(const void* fn, size_t bytes) {
  KDIP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return KDIP_OK;
}

int launch_attention_fwd(const bf16* qkv, int N, int T, int heads, int ch, bf16* out, float* lse, cudaStream_t s) {
  KDIP_REQUIRE(ch == CH, KDIP_ESHAPE, "attention: head channels must be 64 (got %d)", ch);
  KDIP_REQUIRE(T % KC == 0, KDIP_ESHAPE, "attention: T=%d must be a multiple of 64", T);
  if (attention_tc_supported(T, ch)) return launch_attention_fwd_tc(qkv, N, T, heads, out, lse, s);
  if (attention_tcs_supported(T, ch)) return launch_attention_fwd_tcs(qkv, N, T, heads, out, lse, s);
  const size_t smem = (size_t)(2 * QB + 2 * KC) * LD * sizeof(float);
  static bool once = false;
  if (!once) { int rc = set_smem((const void*)attn_fwd_kernel, smem); if (rc) return rc; once = true; }
  attn_fwd_kernel<<<dim3(T / QB, heads, N), 256, smem, s>>>(qkv, T, heads, out, lse);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

int launch_attention_bwd(const bf16* qkv, const bf16* out, const bf16* d_out, const float* lse, int N, int T, int heads, int ch,
                         bf16* dqkv, cudaStream_t s) {
  KDIP_REQUIRE(ch == CH, KDIP_ESHAPE, "attention: head channels must be 64 (got %d)", ch);
  KDIP_REQUIRE(T % KC == 0, KDIP_ESHAPE, "attention: T=%d must be a multiple of 64", T);
  if (attention_tc_supported(T, ch) && !(getenv("KDIP_ATTN_TC_BWD") && atoi(getenv("KDIP_ATTN_TC_BWD")) == 0))
    return launch_attention_bwd_tc(qkv, out, d_out, lse, N, T, heads, dqkv, s);
  if (attention_tcs_supported(T, ch) && !(getenv("KDIP_ATTN_TC_BWD") && atoi(getenv("KDIP_ATTN_TC_BWD")) == 0))
    return launch_attention_bwd_tcs(qkv, out, d_out, lse, N, T, heads, dqkv, s);
  const size_t smem_q = (size_t)(3 * QB + 2 * KC) * LD * sizeof(float);
  const size_t smem_kv = (size_t)((4 * QB + 2 * KC) * LD + 2 * KC) * sizeof(float);
  static bool once = false;
  if (!once) {
    int rc = set_smem((const void*)attn_bwd_dq_kernel, smem_q); if (rc) return rc;
    rc = set_smem((const void*)attn_bwd_dkv_kernel, smem_kv); if (rc) return rc;
    once = true;
  }
  attn_bwd_dq_kernel<<<dim3(T / QB, heads, N), 256, smem_q, s>>>(qkv, out, d_out, lse, T, heads, dqkv);
  KDIP_LAUNCH_CHECK();
  attn_bwd_dkv_kernel<<<dim3(T / QB, heads, N), 256, smem_kv, s>>>(qkv, out, d_out, lse, T, heads, dqkv);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

}  // namespace kdip
