// Orthogonal transforms W^T (forward) / W (inverse) of the lazy posterior covariance Sigma = W diag(theta) W^T
// (condition/utils.py:50-139), replacing the reference's host round-trips through pywt / scipy.fft:
//   DWT: Haar, level 3, axes (-2,-1), packed like pywt.coeffs_to_array (condition/utils.py:116-132).  A level-3 Haar
//        transform only mixes pixels inside 8x8 blocks, so one CTA transforms an 8-row strip in shared memory and
//        writes every sub-band row segment coalesced: one read + one write of the plane (HBM-bound, 2 planes/image).
//   DCT: scipy.fft.dctn(norm='ortho') over ALL axes of [1,3,H,W] (condition/utils.py:91-103): 256-point DCT-II along
//        H and W as fp32 matrix products with the orthonormal DCT matrix, plus the 3-point DCT across channels.
// Both accept an optional per-element multiplier applied to the FORWARD output (theta in the transform domain), which
// fuses the `theta * ot(m)` product of condition/condition.py:182,338,374,427.
#include <stdlib.h>
#include <string.h>

#include "ops.cuh"

namespace kdip {

static constexpr float kInvSqrt2 = 0.70710678118654752440f;

// sub-band placement of pywt.coeffs_to_array: detail on rows -> row offset s, detail on cols -> col offset s
// ('da' = cH -> rows [s,2s), cols [0,s); 'ad' = cV -> rows [0,s), cols [s,2s); 'dd' -> both).  One table so the
// layout can be flipped if a PyWavelets install ever disagrees (parity unpinned, see oracle/transforms_ref.py).
// `swap` = 1 selects the other candidate (KDIP_DWT_LAYOUT=diagram: the PyWavelets docstring diagram, 'da' top-right): the two
// layouts differ only by exchanging the cH / cV blocks of every level.
__device__ __forceinline__ void band_offset(int band /*0 da, 1 ad, 2 dd*/, int s, int swap, int& r0, int& c0) {
  if (swap && band < 2) band ^= 1;
  r0 = (band == 0 || band == 2) ? s : 0;
  c0 = (band == 1 || band == 2) ? s : 0;
}
// KDIP_DWT_LAYOUT: "code" (default; pywt.coeffs_to_array's slicing rule) or "diagram"
static int dwt_layout_swap() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KDIP_DWT_LAYOUT");
    v = (e != nullptr && strcmp(e, "diagram") == 0) ? 1 : 0;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// Haar level-3 forward: grid = planes * S/8 CTAs, one 8-row strip each
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwt_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mul,
                                                      float* __restrict__ out, int S, int mul_planes, int swap) {
  extern __shared__ float sm[];
  float* a0 = sm;                 // [8][S]
  float* a1 = a0 + 8 * S;         // [4][S/2]
  float* a2 = a1 + 2 * S;         // [2][S/4]
  const int strips = S / 8;
  const int p = blockIdx.x / strips, b = blockIdx.x - p * strips;
  const float* src = x + ((size_t)p * S + 8 * b) * S;
  for (int i = threadIdx.x; i < 8 * S; i += blockDim.x) a0[i] = src[i];
  __syncthreads();
  float* dstp = out + (size_t)p * S * S;
  const float* mulp = mul ? mul + (size_t)(p % mul_planes) * S * S : nullptr;
  const float* cur = a0;
  float* nxt = a1;
  int R = 4, C = S / 2;           // output rows / cols of this level
  for (int lvl = 1; lvl <= 3; ++lvl) {
    const int Cin = 2 * C;
    for (int i = threadIdx.x; i < R * C; i += blockDim.x) {
      const int r = i / C, c = i - r * C;
      const float x00 = cur[(2 * r) * Cin + 2 * c], x01 = cur[(2 * r) * Cin + 2 * c + 1];
      const float x10 = cur[(2 * r + 1) * Cin + 2 * c], x11 = cur[(2 * r + 1) * Cin + 2 * c + 1];
      // rows first (axis -2), then columns (axis -1)
      const float lo0 = (x00 + x10) * kInvSqrt2, lo1 = (x01 + x11) * kInvSqrt2;
      const float hi0 = (x00 - x10) * kInvSqrt2, hi1 = (x01 - x11) * kInvSqrt2;
      const float aa = (lo0 + lo1) * kInvSqrt2, ad = (lo0 - lo1) * kInvSqrt2;
      const float da = (hi0 + hi1) * kInvSqrt2, dd = (hi0 - hi1) * kInvSqrt2;
      const float band[3] = {da, ad, dd};
      const int row = R * b + r;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        int r0, c0;
        band_offset(k, C, swap, r0, c0);
        const size_t o = (size_t)(r0 + row) * S + c0 + c;
        dstp[o] = mulp ? band[k] * mulp[o] : band[k];
      }
      if (lvl < 3) nxt[r * C + c] = aa;
      else {
        const size_t o = (size_t)row * S + c;
        dstp[o] = mulp ? aa * mulp[o] : aa;
      }
    }
    __syncthreads();
    cur = nxt;
    nxt = a2;
    R >>= 1;
    C >>= 1;
  }
}

// Haar level-3 inverse: one CTA rebuilds an 8-row strip
__global__ void __launch_bounds__(256) dwt_inv_kernel(const float* __restrict__ cf, float* __restrict__ out, int S, int swap) {
  extern __shared__ float sm[];
  float* a0 = sm;                 // [8][S]   (final)
  float* a1 = a0 + 8 * S;         // [4][S/2]
  float* a2 = a1 + 2 * S;         // [2][S/4]
  float* a3 = a2 + S / 2;         // [1][S/8]
  const int strips = S / 8;
  const int p = blockIdx.x / strips, b = blockIdx.x - p * strips;
  const float* src = cf + (size_t)p * S * S;
  for (int c = threadIdx.x; c < S / 8; c += blockDim.x) a3[c] = src[(size_t)b * S + c];
  __syncthreads();
  const float* cur = a3;
  float* nxt = a2;
  int R = 1, C = S / 8;           // size of the approximation being expanded
  for (int lvl = 3; lvl >= 1; --lvl) {
    for (int i = threadIdx.x; i < R * C; i += blockDim.x) {
      const int r = i / C, c = i - r * C;
      const int row = R * b + r;
      float band[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        int r0, c0;
        band_offset(k, C, swap, r0, c0);
        band[k] = src[(size_t)(r0 + row) * S + c0 + c];
      }
      const float aa = cur[r * C + c], da = band[0], ad = band[1], dd = band[2];
      const float lo0 = (aa + ad) * kInvSqrt2, lo1 = (aa - ad) * kInvSqrt2;
      const float hi0 = (da + dd) * kInvSqrt2, hi1 = (da - dd) * kInvSqrt2;
      const int Co = 2 * C;
      nxt[(2 * r) * Co + 2 * c] = (lo0 + hi0) * kInvSqrt2;
      nxt[(2 * r) * Co + 2 * c + 1] = (lo1 + hi1) * kInvSqrt2;
      nxt[(2 * r + 1) * Co + 2 * c] = (lo0 - hi0) * kInvSqrt2;
      nxt[(2 * r + 1) * Co + 2 * c + 1] = (lo1 - hi1) * kInvSqrt2;
    }
    __syncthreads();
    cur = nxt;
    nxt = (lvl == 3) ? a1 : a0;
    R <<= 1;
    C <<= 1;
  }
  float* dst = out + ((size_t)p * S + 8 * b) * S;
  for (int i = threadIdx.x; i < 8 * S; i += blockDim.x) dst[i] = a0[i];
}

int launch_dwt(const float* x, const float* mul, int mul_planes, float* out, int planes, int S, int inverse, cudaStream_t s) {
  KDIP_REQUIRE(S % 8 == 0 && S >= 8 && S <= 1024, KDIP_ESHAPE, "dwt: S=%d must be a multiple of 8 (level-3 Haar)", S);
  const size_t smem = (size_t)(8 * S + 2 * S + S / 2 + S / 8 + 8) * sizeof(float);
  const int swap = dwt_layout_swap();
  if (!inverse) dwt_fwd_kernel<<<planes * (S / 8), 256, smem, s>>>(x, mul, out, S, mul_planes > 0 ? mul_planes : planes, swap);
  else dwt_inv_kernel<<<planes * (S / 8), 256, smem, s>>>(x, out, S, swap);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// W diag(theta) W^T in ONE pass (S = 256): the posterior covariance applied in image space (condition/utils.py:146-163, used once
// per CG iteration by condition.py:338,374,427).  A level-3 Haar transform only mixes pixels inside 8 x 8 blocks, so a CTA takes
// an 8-row strip down the three levels, scales every coefficient by theta where it is produced, and comes straight back up:
// each thread keeps the scaled details of the 2 x 2 blocks it analysed in registers and synthesises the same blocks on the way
// up, only the approximations go through shared memory.  3 planes of traffic (x, theta, out) instead of the 5 of
// dwt_fwd * theta -> dwt_inv, one launch instead of two; the arithmetic is the two kernels' (the theta product is rounded before
// it is used, as when it went through memory), so the results are bit-identical.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void haar_analyse(float x00, float x01, float x10, float x11, float& aa, float (&band)[3]) {
  const float lo0 = (x00 + x10) * kInvSqrt2, lo1 = (x01 + x11) * kInvSqrt2;
  const float hi0 = (x00 - x10) * kInvSqrt2, hi1 = (x01 - x11) * kInvSqrt2;
  aa = (lo0 + lo1) * kInvSqrt2;
  band[1] = (lo0 - lo1) * kInvSqrt2;   // ad
  band[0] = (hi0 + hi1) * kInvSqrt2;   // da
  band[2] = (hi0 - hi1) * kInvSqrt2;   // dd
}
__device__ __forceinline__ void haar_synthesise(float aa, const float (&band)[3], float& y00, float& y01, float& y10, float& y11) {
  const float da = band[0], ad = band[1], dd = band[2];
  const float lo0 = (aa + ad) * kInvSqrt2, lo1 = (aa - ad) * kInvSqrt2;
  const float hi0 = (da + dd) * kInvSqrt2, hi1 = (da - dd) * kInvSqrt2;
  y00 = (lo0 + hi0) * kInvSqrt2;
  y01 = (lo1 + hi1) * kInvSqrt2;
  y10 = (lo0 - hi0) * kInvSqrt2;
  y11 = (lo1 - hi1) * kInvSqrt2;
}
__global__ void __launch_bounds__(256) dwt_cov_256_kernel(const float* x, const float* __restrict__ theta,
                                                          float* out, int mul_planes, int swap) {   // x may alias out (strip-local)
  constexpr int S = 256;
  __shared__ __align__(16) float a0[8 * S];    // the strip: input, then output
  __shared__ float a1[4 * (S / 2)];
  __shared__ float a2[2 * (S / 4)];
  const int p = blockIdx.x / (S / 8), b = blockIdx.x - p * (S / 8);
  const float4* src = reinterpret_cast<const float4*>(x + ((size_t)p * S + 8 * b) * S);
  for (int i = threadIdx.x; i < 8 * S / 4; i += 256) reinterpret_cast<float4*>(a0)[i] = src[i];
  __syncthreads();
  const float* th = theta + (size_t)(p % mul_planes) * S * S;
  float d1[2][3], d2[3] = {0.f, 0.f, 0.f}, d3[3];
  // level 1: 4 x 128 blocks, two per thread
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = threadIdx.x + 256 * k, r = i >> 7, c = i & 127;
    float aa, band[3];
    haar_analyse(a0[(2 * r) * S + 2 * c], a0[(2 * r) * S + 2 * c + 1], a0[(2 * r + 1) * S + 2 * c], a0[(2 * r + 1) * S + 2 * c + 1], aa, band);
    const int row = 4 * b + r;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      int r0, c0;
      band_offset(q, S / 2, swap, r0, c0);
      d1[k][q] = __fmul_rn(band[q], th[(size_t)(r0 + row) * S + c0 + c]);
    }
    a1[r * (S / 2) + c] = aa;
  }
  __syncthreads();
  // level 2: 2 x 64 blocks
  const int r2 = threadIdx.x >> 6, c2 = threadIdx.x & 63;
  if (threadIdx.x < 128) {
    float aa, band[3];
    haar_analyse(a1[(2 * r2) * (S / 2) + 2 * c2], a1[(2 * r2) * (S / 2) + 2 * c2 + 1], a1[(2 * r2 + 1) * (S / 2) + 2 * c2],
                 a1[(2 * r2 + 1) * (S / 2) + 2 * c2 + 1], aa, band);
    const int row = 2 * b + r2;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      int r0, c0;
      band_offset(q, S / 4, swap, r0, c0);
      d2[q] = __fmul_rn(band[q], th[(size_t)(r0 + row) * S + c0 + c2]);
    }
    a2[r2 * (S / 4) + c2] = aa;
  }
  __syncthreads();
  // level 3: 1 x 32 blocks, down and straight back up (a thread's block of a2 is read and rewritten by that thread only)
  if (threadIdx.x < 32) {
    const int c = threadIdx.x;
    float aa, band[3];
    haar_analyse(a2[2 * c], a2[2 * c + 1], a2[(S / 4) + 2 * c], a2[(S / 4) + 2 * c + 1], aa, band);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      int r0, c0;
      band_offset(q, S / 8, swap, r0, c0);
      d3[q] = __fmul_rn(band[q], th[(size_t)(r0 + b) * S + c0 + c]);
    }
    aa = __fmul_rn(aa, th[(size_t)b * S + c]);
    haar_synthesise(aa, d3, a2[2 * c], a2[2 * c + 1], a2[(S / 4) + 2 * c], a2[(S / 4) + 2 * c + 1]);
  }
  __syncthreads();
  if (threadIdx.x < 128)
    haar_synthesise(a2[r2 * (S / 4) + c2], d2, a1[(2 * r2) * (S / 2) + 2 * c2], a1[(2 * r2) * (S / 2) + 2 * c2 + 1],
                    a1[(2 * r2 + 1) * (S / 2) + 2 * c2], a1[(2 * r2 + 1) * (S / 2) + 2 * c2 + 1]);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = threadIdx.x + 256 * k, r = i >> 7, c = i & 127;
    haar_synthesise(a1[r * (S / 2) + c], d1[k], a0[(2 * r) * S + 2 * c], a0[(2 * r) * S + 2 * c + 1], a0[(2 * r + 1) * S + 2 * c],
                    a0[(2 * r + 1) * S + 2 * c + 1]);
  }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(out + ((size_t)p * S + 8 * b) * S);
  for (int i = threadIdx.x; i < 8 * S / 4; i += 256) dst[i] = reinterpret_cast<const float4*>(a0)[i];
}

// out = W (theta .* W^T x): the fused kernel at S = 256 (KDIP_DWT_FUSED=0: the two transforms through `tmp`)
int launch_dwt_cov(const float* x, const float* theta, int theta_planes, float* tmp, float* out, int planes, int S, cudaStream_t s) {
  const bool fused = S == 256 && !(getenv("KDIP_DWT_FUSED") && atoi(getenv("KDIP_DWT_FUSED")) == 0);
  if (fused) {
    dwt_cov_256_kernel<<<planes * (S / 8), 256, 0, s>>>(x, theta, out, theta_planes > 0 ? theta_planes : planes, dwt_layout_swap());
    KDIP_LAUNCH_CHECK();
    return KDIP_OK;
  }
  int rc = launch_dwt(x, theta, theta_planes, tmp, planes, S, 0, s);
  if (rc) return rc;
  return launch_dwt(tmp, nullptr, 0, out, planes, S, 1, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// DCT-II (orthonormal) as matrix products
// ---------------------------------------------------------------------------------------------------------------------
// Cm[k][n] = alpha_k cos(pi (2n+1) k / (2S))
__global__ void dct_matrix_kernel(float* __restrict__ Cm, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * S) return;
  const int k = i / S, n = i - k * S;
  const double a = (k == 0) ? sqrt(1.0 / S) : sqrt(2.0 / S);
  Cm[i] = (float)(a * cospi((double)((2 * n + 1) * k) / (double)(2 * S)));
}

// out[z][m][n] = sum_k A[z*sAz + m*sAm + k*sAk] * B[z*sBz + k*sBk + n*sBn] ;  M, N, K multiples of 64 / 64 / 16
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ A, long sAz, int sAm, int sAk,
                                                            const float* __restrict__ B, long sBz, int sBk, int sBn,
                                                            float* __restrict__ Cout, long sCz, int K, int N) {
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const int z = blockIdx.z;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const float* Ap = A + (size_t)z * sAz;
  const float* Bp = B + (size_t)z * sBz;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      // choose the faster-varying index by stride so global reads coalesce for either orientation
      int kk, mm;
      if (sAk == 1) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }
      As[kk][mm] = Ap[(size_t)(m0 + mm) * sAm + (size_t)(k0 + kk) * sAk];
      int kb, nn;
      if (sBk == 1) { kb = i & 15; nn = i >> 4; } else { nn = i & 63; kb = i >> 6; }
      Bs[kb][nn] = Bp[(size_t)(k0 + kb) * sBk + (size_t)(n0 + nn) * sBn];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  float* Cp = Cout + (size_t)z * sCz;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Cp[(size_t)(m0 + ty * 4 + i) * N + n0 + tx * 4 + j] = acc[i][j];
}

// 3-point orthonormal DCT-II across the channel axis of [images][3][HW] (+ optional multiplier on the output)
__global__ void dct3_chan_kernel(const float* __restrict__ x, const float* __restrict__ mul, float* __restrict__ out, size_t HW,
                                 int inverse, int mul_images) {
  const float r3 = 0.57735026918962576451f, r2 = 0.70710678118654752440f, r6 = 0.40824829046386301637f;
  const size_t img = blockIdx.y;
  const float* xp = x + img * 3 * HW;
  float* op = out + img * 3 * HW;
  const float* mp = mul ? mul + (img % mul_images) * 3 * HW : nullptr;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x) {
    const float a = xp[i], b = xp[HW + i], c = xp[2 * HW + i];
    float y0, y1, y2;
    if (!inverse) {
      y0 = (a + b + c) * r3;
      y1 = (a - c) * r2;
      y2 = (a - 2.f * b + c) * r6;
    } else {
      y0 = a * r3 + b * r2 + c * r6;
      y1 = a * r3 - 2.f * c * r6;
      y2 = a * r3 - b * r2 + c * r6;
    }
    if (mp) { y0 *= mp[i]; y1 *= mp[HW + i]; y2 *= mp[2 * HW + i]; }
    op[i] = y0; op[HW + i] = y1; op[2 * HW + i] = y2;
  }
}

size_t dct_workspace_bytes(int planes, int S) { return ((size_t)S * S + (size_t)planes * S * S) * sizeof(float) + 512; }

// x, out: [images][3][S][S]; ws: dct_workspace_bytes.  forward: out = mul .* DCT(x); inverse: out = IDCT(x)
int launch_dct(const float* x, const float* mul, int mul_images, float* out, int images, int S, int inverse, void* ws,
               cudaStream_t s) {
  KDIP_REQUIRE(S % 64 == 0, KDIP_ESHAPE, "dct: S=%d must be a multiple of 64", S);
  KDIP_REQUIRE(ws != nullptr, KDIP_ENOMEM, "dct: workspace required");
  const int planes = images * 3;
  float* Cm = reinterpret_cast<float*>(ws);
  float* T = Cm + (((size_t)S * S + 63) & ~(size_t)63);
  dct_matrix_kernel<<<(S * S + 255) / 256, 256, 0, s>>>(Cm, S);
  KDIP_LAUNCH_CHECK();
  const long PS = (long)S * S;
  dim3 grid(S / 64, S / 64, planes);
  const size_t HW = (size_t)S * S;
  dim3 g3((unsigned)((HW + 255) / 256 < 1184 ? (HW + 255) / 256 : 1184), images);
  if (!inverse) {
    // along W: T[y][k] = sum_n X[y][n] C[k][n]   (A = X row-major, B[k'][n'] = C[n'][k'])
    sgemm_strided_kernel<<<grid, 256, 0, s>>>(x, PS, S, 1, Cm, 0, 1, S, T, PS, S, S);
    KDIP_LAUNCH_CHECK();
    // along H: Y[k][x] = sum_n C[k][n] T[n][x]
    sgemm_strided_kernel<<<grid, 256, 0, s>>>(Cm, 0, S, 1, T, PS, S, 1, out, PS, S, S);
    KDIP_LAUNCH_CHECK();
    dct3_chan_kernel<<<g3, 256, 0, s>>>(out, mul, out, HW, 0, mul_images > 0 ? mul_images : images);
    KDIP_LAUNCH_CHECK();
  } else {
    dct3_chan_kernel<<<g3, 256, 0, s>>>(x, nullptr, out, HW, 1, images);
    KDIP_LAUNCH_CHECK();
    // along W: T[y][n] = sum_k Y[y][k] C[k][n]
    sgemm_strided_kernel<<<grid, 256, 0, s>>>(out, PS, S, 1, Cm, 0, S, 1, T, PS, S, S);
    KDIP_LAUNCH_CHECK();
    // along H: X[n][x] = sum_k C[k][n] T[k][x]   (A[m=n][k] = C[k][n])
    sgemm_strided_kernel<<<grid, 256, 0, s>>>(Cm, 0, 1, S, T, PS, S, 1, out, PS, S, S);
    KDIP_LAUNCH_CHECK();
  }
  return KDIP_OK;
}

}  // namespace kdip
