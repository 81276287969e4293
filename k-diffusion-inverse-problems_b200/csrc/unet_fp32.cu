// fp32 reference-precision engine of the ADM UNet: forward and input-VJP in plain fp32 FMA arithmetic on the CUDA cores.
//
// Same reference as unet.cu (guided_diffusion/unet.py:143-668, nn.py:17-121) and the same C ABI (kdip_unet_* with
// kdip_unet_arch.precision = KDIP_PRECISION_FP32), at the reference's own arithmetic: the reference runs fp32
// (use_fp16=False, condition/diffpir_utils/utils_model.py:364).  The bf16 tcgen05 engine is the product's fast path; this
// engine exists so that the guided path can be held to a tight, fixed tolerance against the reference's outputs - with the
// synthetic weights hat_x0 = clip(x0 + sigma^2 J^T v) amplifies a relative perturbation of the network by 20-80x at sigma = 10
// (measured on the oracle: a 6e-8 relative weight perturbation moves hat_x0 by 5e-6..2e-5, a 4e-6 one - the level of a
// 3 x bf16 operand split - by 1e-4..3.5e-4 and a 4-step trajectory by 7e-2), so only fp32 accumulation of fp32 operands
// separates "kernel wrong" from "bf16 + ill-conditioning" everywhere.
//
// Layout: activations fp32 NHWC, one value buffer and one gradient buffer per tensor, all inside the caller's workspace.
// Forward = a tape of unfused ops (conv / GroupNorm+act / resample / attention); the VJP walks the tape backwards and every op
// ACCUMULATES into its inputs' gradient buffers (zeroed at the start), which covers the skip stack and the ResBlock skip path
// with no special cases.  Nothing here is tuned: a 64x64x16 shared-memory tile SGEMM-style implicit conv, strided GroupNorm
// reductions in fp64, one CTA per attention row.
#include <stdlib.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "kdip_common.cuh"
#include "unet_plan.h"

namespace kdip {
namespace f32 {

static constexpr float kEps = 1e-5f;

static inline int ew_grid(size_t total) {
  size_t b = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}
#define F32_LOOP(i, n) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

// ---------------------------------------------------------------------------------------------------------------------
// implicit-GEMM convolution (3x3 pad 1, or 1x1), NHWC fp32, two-source (concatenated) input, two-destination output
// ---------------------------------------------------------------------------------------------------------------------
struct ConvArgs {
  const float* s0; int C0;
  const float* s1; int C1;
  const float* w;        // [taps][C0+C1][Cout], Cout contiguous
  const float* bias;     // [Cout] or NULL
  const float* res;      // [N,H,W,Cout] added to the result, or NULL
  float* d0; int D0;     // output channels [0, D0)
  float* d1; int D1;     // output channels [D0, D0+D1)
  int accumulate;        // 1: destinations += result
  int N, H, W, taps;
};

__global__ void __launch_bounds__(256) conv_f32_kernel(const ConvArgs a) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int Cin = a.C0 + a.C1, Cout = a.D0 + a.D1;
  const size_t M = (size_t)a.N * a.H * a.W;
  const size_t m0 = (size_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // A loads: pixel lm, 4 channels from kq
  const int lm = tid >> 2, kq = (tid & 3) * 4;
  const size_t pm = m0 + lm;
  const bool pvalid = pm < M;
  int pn = 0, py = 0, px = 0;
  if (pvalid) {
    px = (int)(pm % a.W);
    py = (int)((pm / a.W) % a.H);
    pn = (int)(pm / ((size_t)a.W * a.H));
  }
  // B loads: row kb, 4 columns from nq
  const int kb = tid >> 4, nq = (tid & 15) * 4;
  const bool vecA = (a.C0 % 4 == 0) && (a.C1 % 4 == 0);
  const bool vecB = (Cout % 4 == 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int tap = 0; tap < a.taps; ++tap) {
    const int dy = a.taps == 9 ? tap / 3 - 1 : 0, dx = a.taps == 9 ? tap % 3 - 1 : 0;
    const int yy = py + dy, xx = px + dx;
    const bool inb = pvalid && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
    const size_t pix = ((size_t)pn * a.H + yy) * a.W + xx;
    for (int k0 = 0; k0 < Cin; k0 += BK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      const int c = k0 + kq;
      if (inb) {
        if (vecA && c + 3 < Cin) {
          const float4 v = (c < a.C0) ? *reinterpret_cast<const float4*>(a.s0 + pix * a.C0 + c)
                                      : *reinterpret_cast<const float4*>(a.s1 + pix * a.C1 + (c - a.C0));
          av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cj = c + j;
            if (cj < Cin) av[j] = (cj < a.C0) ? a.s0[pix * a.C0 + cj] : a.s1[pix * a.C1 + (cj - a.C0)];
          }
        }
      }
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      const int kk = k0 + kb, nn = n0 + nq;
      if (kk < Cin) {
        const float* wp = a.w + ((size_t)tap * Cin + kk) * Cout + nn;
        if (vecB && nn + 3 < Cout) {
          const float4 v = *reinterpret_cast<const float4*>(wp);
          bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (nn + j < Cout) bv[j] = wp[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[kq + j][lm] = av[j];
        Bs[kb][nq + j] = bv[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Cout) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[n];
      if (a.res) v += a.res[m * Cout + n];
      float* dst = (n < a.D0) ? a.d0 + m * a.D0 + n : a.d1 + m * a.D1 + (n - a.D0);
      if (a.accumulate) *dst += v; else *dst = v;
    }
  }
}

static int launch_conv(const ConvArgs& a, cudaStream_t s) {
  const size_t M = (size_t)a.N * a.H * a.W;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((a.D0 + a.D1 + 63) / 64));
  conv_f32_kernel<<<grid, 256, 0, s>>>(a);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

// OIHW fp32 -> forward [tap][ci][co] (flip = 0) or input-gradient operator [tap][co][ci] with w[co][ci][taps-1-tap] (flip = 1)
__global__ void pack_f32_kernel(const float* __restrict__ w, int O, int I, int taps, int flip, float* __restrict__ dst) {
  const size_t total = (size_t)taps * O * I;
  F32_LOOP(i, total) {
    if (!flip) {
      const int co = (int)(i % O), ci = (int)((i / O) % I), tap = (int)(i / ((size_t)O * I));
      dst[i] = w[((size_t)co * I + ci) * taps + tap];
    } else {
      const int ci = (int)(i % I), co = (int)((i / I) % O), tap = (int)(i / ((size_t)O * I));
      dst[i] = w[((size_t)co * I + ci) * taps + (taps - 1 - tap)];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// layout converters at the boundary (NCHW fp32 <-> NHWC fp32)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_f32_kernel(const float* __restrict__ src, const float* __restrict__ scale, int C, size_t HW,
                                        float* __restrict__ dst, size_t total) {
  F32_LOOP(i, total) {
    const int c = (int)(i % C);
    const size_t np = i / C, n = np / HW, p = np % HW;
    const float v = src[(n * C + c) * HW + p];
    dst[i] = scale ? v * scale[n] : v;
  }
}
__global__ void nhwc_to_nchw_f32_kernel(const float* __restrict__ src, int C, size_t HW, float* __restrict__ dst, size_t total) {
  F32_LOOP(i, total) {
    const size_t p = i % HW, nc = i / HW, n = nc / C, c = nc % C;
    dst[i] = src[(n * HW + p) * C + c];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// GroupNorm32 (+FiLM, +SiLU), nn.py:17-19, unet.py:237-253, and its backward
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld2(const float* s0, int C0, const float* s1, int C1, size_t np, int c) {
  return (c < C0) ? s0[np * C0 + c] : s1[np * C1 + (c - C0)];
}
__device__ __forceinline__ double block_sum_d(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// mr[n][g] = (mean, rstd) over the group's channels and all pixels (biased variance, eps 1e-5)
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ s0, int C0, const float* __restrict__ s1, int C1,
                                                        int P, float* __restrict__ mr) {
  __shared__ double red[8];
  const int g = blockIdx.x, n = blockIdx.y, C = C0 + C1, cpg = C / 32;
  const size_t total = (size_t)P * cpg;
  double s = 0.0, q = 0.0;
  for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
    const size_t p = i / cpg;
    const int c = g * cpg + (int)(i - p * cpg);
    const double v = (double)ld2(s0, C0, s1, C1, (size_t)n * P + p, c);
    s += v;
    q += v * v;
  }
  s = block_sum_d(s, red);
  q = block_sum_d(q, red);
  if (threadIdx.x == 0) {
    const double mean = s / (double)total;
    double var = q / (double)total - mean * mean;
    if (var < 0.0) var = 0.0;
    mr[((size_t)n * 32 + g) * 2] = (float)mean;
    mr[((size_t)n * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)kEps));
  }
}
// per-(image, channel) affine u = A x + B: A = gamma rstd (1 + scale), B = (beta - mean gamma rstd)(1 + scale) + shift
__global__ void gn_coef_kernel(const float* __restrict__ mr, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ film, int film_stride, int film_off, int C, float* __restrict__ ab, int N) {
  const int cpg = C / 32;
  F32_LOOP(i, (size_t)N * C) {
    const int c = (int)(i % C), n = (int)(i / C), g = c / cpg;
    const float mean = mr[((size_t)n * 32 + g) * 2], rstd = mr[((size_t)n * 32 + g) * 2 + 1];
    float A = gamma[c] * rstd, B = beta[c] - mean * A;
    if (film) {
      const float sc = film[(size_t)n * film_stride + film_off + c], sh = film[(size_t)n * film_stride + film_off + C + c];
      A = A * (1.f + sc);
      B = B * (1.f + sc) + sh;
    }
    ab[i * 2] = A;
    ab[i * 2 + 1] = B;
  }
}
__device__ __forceinline__ float silu_p(float u) { return u / (1.f + expf(-u)); }
__device__ __forceinline__ float dsilu_p(float u) {
  const float s = 1.f / (1.f + expf(-u));
  return s * (1.f + u * (1.f - s));
}
__global__ void gn_apply_f32_kernel(const float* __restrict__ s0, int C0, const float* __restrict__ s1, int C1, int P,
                                    const float* __restrict__ ab, int silu, float* __restrict__ out, size_t total) {
  const int C = C0 + C1;
  F32_LOOP(i, total) {
    const int c = (int)(i % C);
    const size_t np = i / C, n = np / P;
    const float u = fmaf(ab[(n * C + c) * 2], ld2(s0, C0, s1, C1, np, c), ab[(n * C + c) * 2 + 1]);
    out[i] = silu ? silu_p(u) : u;
  }
}
// red[n][g] = (sum g_xhat, sum g_xhat xhat), g_xhat = g_out act'(u) gamma(1+scale), xhat = (x - mean) rstd
__global__ void __launch_bounds__(256) gn_bwd_stats_kernel(const float* __restrict__ s0, int C0, const float* __restrict__ s1, int C1,
                                                            int P, const float* __restrict__ ab, const float* __restrict__ mr, int silu,
                                                            const float* __restrict__ gout, double* __restrict__ red_out) {
  __shared__ double red[8];
  const int g = blockIdx.x, n = blockIdx.y, C = C0 + C1, cpg = C / 32;
  const size_t total = (size_t)P * cpg;
  const float mean = mr[((size_t)n * 32 + g) * 2], rstd = mr[((size_t)n * 32 + g) * 2 + 1];
  double s = 0.0, q = 0.0;
  for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
    const size_t p = i / cpg;
    const int c = g * cpg + (int)(i - p * cpg);
    const size_t np = (size_t)n * P + p;
    const float x = ld2(s0, C0, s1, C1, np, c);
    const float A = ab[((size_t)n * C + c) * 2], B = ab[((size_t)n * C + c) * 2 + 1];
    float gu = gout[np * C + c];
    if (silu) gu *= dsilu_p(fmaf(A, x, B));
    const double gx = (double)gu * (double)(A / rstd);
    s += gx;
    q += gx * (double)((x - mean) * rstd);
  }
  s = block_sum_d(s, red);
  q = block_sum_d(q, red);
  if (threadIdx.x == 0) {
    red_out[((size_t)n * 32 + g) * 2] = s / (double)total;
    red_out[((size_t)n * 32 + g) * 2 + 1] = q / (double)total;
  }
}
// g_x = rstd (g_xhat - mean(g_xhat) - xhat mean(g_xhat xhat)), accumulated into the gradient buffers of the two sources
__global__ void gn_bwd_apply_f32_kernel(const float* __restrict__ s0, int C0, const float* __restrict__ s1, int C1, int P,
                                        const float* __restrict__ ab, const float* __restrict__ mr, const double* __restrict__ red,
                                        int silu, const float* __restrict__ gout, float* __restrict__ g0, float* __restrict__ g1,
                                        size_t total) {
  const int C = C0 + C1, cpg = C / 32;
  F32_LOOP(i, total) {
    const int c = (int)(i % C), g = c / cpg;
    const size_t np = i / C, n = np / P;
    const float mean = mr[(n * 32 + g) * 2], rstd = mr[(n * 32 + g) * 2 + 1];
    const float x = ld2(s0, C0, s1, C1, np, c);
    const float A = ab[(n * C + c) * 2], B = ab[(n * C + c) * 2 + 1];
    float gu = gout[i];
    if (silu) gu *= dsilu_p(fmaf(A, x, B));
    const float gxh = gu * (A / rstd);
    const float xh = (x - mean) * rstd;
    const float gx = rstd * (gxh - (float)red[(n * 32 + g) * 2] - xh * (float)red[(n * 32 + g) * 2 + 1]);
    if (c < C0) g0[np * C0 + c] += gx; else g1[np * C1 + (c - C0)] += gx;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// resampling (unet.py:100-110 nearest x2, :136-140 AvgPool2d(2)), elementwise helpers
// ---------------------------------------------------------------------------------------------------------------------
__global__ void avgpool2_kernel(const float* __restrict__ x, int H, int W, int C, float* __restrict__ out, size_t total) {
  const int Ho = H / 2, Wo = W / 2;
  F32_LOOP(i, total) {
    const int c = (int)(i % C), xo = (int)((i / C) % Wo), yo = (int)((i / ((size_t)C * Wo)) % Ho);
    const size_t n = i / ((size_t)C * Wo * Ho);
    const float* p = x + ((n * H + 2 * yo) * W + 2 * xo) * C + c;
    out[i] = 0.25f * ((p[0] + p[C]) + (p[(size_t)W * C] + p[(size_t)W * C + C]));
  }
}
// gx[n, y, x, c] += 0.25 * gout[n, y/2, x/2, c]
__global__ void avgpool2_bwd_kernel(const float* __restrict__ gout, int H, int W, int C, float* __restrict__ gx, size_t total) {
  F32_LOOP(i, total) {
    const int c = (int)(i % C), xx = (int)((i / C) % W), yy = (int)((i / ((size_t)C * W)) % H);
    const size_t n = i / ((size_t)C * W * H);
    gx[i] += 0.25f * gout[((n * (H / 2) + yy / 2) * (W / 2) + xx / 2) * C + c];
  }
}
__global__ void upsample2_kernel(const float* __restrict__ x, int H, int W, int C, float* __restrict__ out, size_t total) {
  const int Ho = 2 * H, Wo = 2 * W;
  F32_LOOP(i, total) {
    const int c = (int)(i % C), xo = (int)((i / C) % Wo), yo = (int)((i / ((size_t)C * Wo)) % Ho);
    const size_t n = i / ((size_t)C * Wo * Ho);
    out[i] = x[((n * H + yo / 2) * W + xo / 2) * C + c];
  }
}
// gx[n, y, x, c] += sum of the 2x2 replicas of gout
__global__ void upsample2_bwd_kernel(const float* __restrict__ gout, int H, int W, int C, float* __restrict__ gx, size_t total) {
  F32_LOOP(i, total) {
    const int c = (int)(i % C), xx = (int)((i / C) % W), yy = (int)((i / ((size_t)C * W)) % H);
    const size_t n = i / ((size_t)C * W * H);
    const float* p = gout + ((n * 2 * H + 2 * yy) * 2 * W + 2 * xx) * C + c;
    gx[i] += (p[0] + p[C]) + (p[(size_t)2 * W * C] + p[(size_t)2 * W * C + C]);
  }
}
__global__ void add_into_kernel(float* __restrict__ y, const float* __restrict__ x, size_t n) { F32_LOOP(i, n) y[i] += x[i]; }

// ---------------------------------------------------------------------------------------------------------------------
// timestep embedding + emb_layers with precise transcendental functions (nn.py:103-121, unet.py:199-205,473-477)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void time_embed_f32_kernel(const float* __restrict__ t, int mc, const float* __restrict__ w1, const float* __restrict__ b1,
                                      const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ semb) {
  extern __shared__ float sm[];  // e0[mc] | h1[4mc]
  float* e0 = sm;
  float* h1 = sm + mc;
  const int n = blockIdx.x, ted = 4 * mc, half = mc / 2;
  const float tv = t[n];
  for (int k = threadIdx.x; k < mc; k += blockDim.x) {
    const int kk = k < half ? k : k - half;
    const float a = tv * expf(-logf(10000.f) * (float)kk / (float)half);
    e0[k] = k < half ? cosf(a) : sinf(a);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < ted; j += nw) {
    float acc = 0.f;
    for (int k = lane; k < mc; k += 32) acc += w1[(size_t)j * mc + k] * e0[k];
    acc = warp_sum(acc);
    if (lane == 0) h1[j] = silu_p(acc + b1[j]);
  }
  __syncthreads();
  for (int j = warp; j < ted; j += nw) {
    float acc = 0.f;
    for (int k = lane; k < ted; k += 32) acc += w2[(size_t)j * ted + k] * h1[k];
    acc = warp_sum(acc);
    if (lane == 0) semb[(size_t)n * ted + j] = silu_p(acc + b2[j]);     // the SiLU that opens every emb_layers (unet.py:199-200)
  }
}
__global__ void emb_proj_f32_kernel(const float* __restrict__ semb, int N, int ted, const float* __restrict__ wall,
                                    const float* __restrict__ ball, int R, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= R) return;
  for (int n = 0; n < N; ++n) {
    float acc = 0.f;
    for (int k = lane; k < ted; k += 32) acc += wall[(size_t)r * ted + k] * semb[(size_t)n * ted + k];
    acc = warp_sum(acc);
    if (lane == 0) out[(size_t)n * R + r] = acc + ball[r];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// QKVAttentionLegacy (unet.py:339-356): qkv [N,T,3C], channel = head*192 + {q,k,v}*64 + c; softmax in fp32
// ---------------------------------------------------------------------------------------------------------------------
static constexpr int AT_THREADS = 128;
__device__ __forceinline__ float block_max_f(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) t = fmaxf(t, red[i]);
  return t;
}
__device__ __forceinline__ float block_sum_f(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}
// grid (T, heads, N): one query row per CTA
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_f32_kernel(const float* __restrict__ qkv, int T, int heads, float* __restrict__ out,
                                                                   float* __restrict__ lse) {
  extern __shared__ float sm[];   // sc[T] | q[64] | part[2][64] | red[8]
  float* sc = sm;
  float* q = sc + T;
  float* part = q + 64;
  float* red = part + 128;
  const int t = blockIdx.x, h = blockIdx.y, n = blockIdx.z, C = heads * 64;
  const float scale = 1.f / sqrtf(sqrtf(64.f));
  const float* base = qkv + (size_t)n * T * 3 * C + (size_t)h * 192;
  if (threadIdx.x < 64) q[threadIdx.x] = base[(size_t)t * 3 * C + threadIdx.x] * scale;
  __syncthreads();
  float mx = -INFINITY;
  for (int s = threadIdx.x; s < T; s += AT_THREADS) {
    const float* kp = base + (size_t)s * 3 * C + 64;
    float d = 0.f;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) d = fmaf(q[c], kp[c] * scale, d);
    sc[s] = d;
    mx = fmaxf(mx, d);
  }
  mx = block_max_f(mx, red);
  float sum = 0.f;
  for (int s = threadIdx.x; s < T; s += AT_THREADS) {
    const float e = expf(sc[s] - mx);
    sc[s] = e;
    sum += e;
  }
  sum = block_sum_f(sum, red);
  __syncthreads();
  const int c = threadIdx.x & 63, hf = threadIdx.x >> 6;
  float acc = 0.f;
  for (int s = hf; s < T; s += 2) acc = fmaf(sc[s], base[(size_t)s * 3 * C + 128 + c], acc);
  part[hf * 64 + c] = acc;
  __syncthreads();
  if (threadIdx.x < 64) out[((size_t)n * T + t) * C + h * 64 + c] = (part[c] + part[64 + c]) / sum;
  if (threadIdx.x == 0) lse[((size_t)n * heads + h) * T + t] = mx + logf(sum);
}
// query side: dQ[t] and D[t] = sum_s P dP
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_q_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                                     const float* __restrict__ lse, int T, int heads,
                                                                     float* __restrict__ dqkv, float* __restrict__ Dbuf) {
  extern __shared__ float sm[];   // ds[T] | q[64] | go[64] | part[128] | red[8]
  float* ds = sm;
  float* q = ds + T;
  float* go = q + 64;
  float* part = go + 64;
  float* red = part + 128;
  const int t = blockIdx.x, h = blockIdx.y, n = blockIdx.z, C = heads * 64;
  const float scale = 1.f / sqrtf(sqrtf(64.f));
  const float* base = qkv + (size_t)n * T * 3 * C + (size_t)h * 192;
  if (threadIdx.x < 64) {
    q[threadIdx.x] = base[(size_t)t * 3 * C + threadIdx.x] * scale;
    go[threadIdx.x] = dout[((size_t)n * T + t) * C + h * 64 + threadIdx.x];
  }
  __syncthreads();
  const float l = lse[((size_t)n * heads + h) * T + t];
  float dsum = 0.f;
  for (int s = threadIdx.x; s < T; s += AT_THREADS) {
    const float* kp = base + (size_t)s * 3 * C + 64;
    const float* vp = kp + 64;
    float d = 0.f, dp = 0.f;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
      d = fmaf(q[c], kp[c] * scale, d);
      dp = fmaf(go[c], vp[c], dp);
    }
    const float p = expf(d - l);
    ds[s] = p;                 // P for now; dP is recomputed in the second pass once D is known
    dsum = fmaf(p, dp, dsum);
  }
  const float D = block_sum_f(dsum, red);
  __syncthreads();
  for (int s = threadIdx.x; s < T; s += AT_THREADS) {
    const float* vp = base + (size_t)s * 3 * C + 128;
    float dp = 0.f;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) dp = fmaf(go[c], vp[c], dp);
    ds[s] = ds[s] * (dp - D);
  }
  __syncthreads();
  const int c = threadIdx.x & 63, hf = threadIdx.x >> 6;
  float acc = 0.f;
  for (int s = hf; s < T; s += 2) acc = fmaf(ds[s], base[(size_t)s * 3 * C + 64 + c] * scale, acc);
  part[hf * 64 + c] = acc;
  __syncthreads();
  if (threadIdx.x < 64) dqkv[((size_t)n * T + t) * 3 * C + h * 192 + c] = (part[c] + part[64 + c]) * scale;
  if (threadIdx.x == 0) Dbuf[((size_t)n * heads + h) * T + t] = D;
}
// key side: dK[s], dV[s]
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_kv_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                                      const float* __restrict__ lse, const float* __restrict__ Dbuf,
                                                                      int T, int heads, float* __restrict__ dqkv) {
  extern __shared__ float sm[];   // p[T] | ds[T] | k[64] | v[64] | part[256]
  float* pp = sm;
  float* ds = pp + T;
  float* k = ds + T;
  float* v = k + 64;
  float* part = v + 64;
  const int s = blockIdx.x, h = blockIdx.y, n = blockIdx.z, C = heads * 64;
  const float scale = 1.f / sqrtf(sqrtf(64.f));
  const float* base = qkv + (size_t)n * T * 3 * C + (size_t)h * 192;
  if (threadIdx.x < 64) {
    k[threadIdx.x] = base[(size_t)s * 3 * C + 64 + threadIdx.x] * scale;
    v[threadIdx.x] = base[(size_t)s * 3 * C + 128 + threadIdx.x];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += AT_THREADS) {
    const float* qp = base + (size_t)t * 3 * C;
    const float* gp = dout + ((size_t)n * T + t) * C + h * 64;
    float d = 0.f, dp = 0.f;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
      d = fmaf(qp[c] * scale, k[c], d);
      dp = fmaf(gp[c], v[c], dp);
    }
    const float p = expf(d - lse[((size_t)n * heads + h) * T + t]);
    pp[t] = p;
    ds[t] = p * (dp - Dbuf[((size_t)n * heads + h) * T + t]);
  }
  __syncthreads();
  const int c = threadIdx.x & 63, hf = threadIdx.x >> 6;
  float av = 0.f, ak = 0.f;
  for (int t = hf; t < T; t += 2) {
    av = fmaf(pp[t], dout[((size_t)n * T + t) * C + h * 64 + c], av);
    ak = fmaf(ds[t], base[(size_t)t * 3 * C + c] * scale, ak);
  }
  part[hf * 64 + c] = av;
  part[128 + hf * 64 + c] = ak;
  __syncthreads();
  if (threadIdx.x < 64) {
    float* o = dqkv + ((size_t)n * T + s) * 3 * C + h * 192;
    o[128 + c] = part[c] + part[64 + c];
    o[64 + c] = (part[128 + c] + part[192 + c]) * scale;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------------------------------
struct Ten {
  float* v = nullptr;
  float* g = nullptr;
  int N = 0, H = 0, W = 0, C = 0;
  size_t numel() const { return (size_t)N * H * W * C; }
};

typedef std::function<int(cudaStream_t)> Op;

struct ResW {
  const float *g1, *b1, *g2, *b2, *w1, *w1d, *w2, *w2d, *ws, *wsd, *bias1, *bias2, *bias_s;
  int film_off;
};
struct AttW {
  const float *g, *b, *wqkv, *wqkvd, *wproj, *wprojd, *bqkv, *bproj;
};

}  // namespace f32

struct Fp32Engine {
  kdip_unet_arch arch;
  std::vector<BlockDesc> plan;
  std::vector<void*> owned;
  std::vector<f32::ResW> resw;
  std::vector<f32::AttW> attw;
  const float *tw1 = nullptr, *tb1 = nullptr, *tw2 = nullptr, *tb2 = nullptr;
  const float *w_in = nullptr, *w_ind = nullptr, *b_in = nullptr;
  const float *g_head = nullptr, *be_head = nullptr, *w_head = nullptr, *w_headd = nullptr, *b_head = nullptr;
  const float *w_cov = nullptr, *b_cov = nullptr;
  float *wall = nullptr, *ball = nullptr;
  int R = 0;
  bool has_cov = false;
  // launch plan
  int planned_N = 0;
  void* planned_ws = nullptr;
  size_t planned_bytes = 0;
  std::vector<f32::Op> fwd, bwd;
  void* grad_base = nullptr;
  size_t grad_bytes = 0;
  f32::Ten hlast;
  // per-call I/O
  const float *io_x = nullptr, *io_xscale = nullptr, *io_t = nullptr, *io_seed = nullptr;
  float *io_out = nullptr, *io_cov = nullptr, *io_grad = nullptr;
};

namespace f32 {

static int dev_alloc(Fp32Engine* e, size_t bytes, void** out) {
  void* p = nullptr;
  KDIP_CUDA(cudaMalloc(&p, bytes));
  e->owned.push_back(p);
  *out = p;
  return KDIP_OK;
}

struct Arena {
  char* base;
  size_t cur = 0;
  void* take(size_t bytes) {
    const size_t o = cur;
    cur += (bytes + 255) & ~(size_t)255;
    return base ? (void*)(base + o) : nullptr;
  }
};

// The plan is built twice: sizing (base == NULL, nothing emitted) and emission.  Value buffers first, then ONE contiguous
// gradient region (zeroed by the first backward op).
struct Builder {
  Fp32Engine* e;
  int N;
  Arena val, grd;
  bool emit;
  std::vector<Op>* F;
  std::vector<Op> B;     // backward ops in forward order; reversed at the end
  float* film = nullptr;

  Ten make(int H, int W, int C, bool need_grad = true) {
    Ten t;
    t.N = N; t.H = H; t.W = W; t.C = C;
    t.v = (float*)val.take(t.numel() * 4);
    t.g = need_grad ? (float*)grd.take(t.numel() * 4) : nullptr;
    return t;
  }
  float* scratch(size_t floats) { return (float*)val.take(floats * 4); }

  // out = conv(concat(a, b)) + bias (+ res)
  Ten conv(const Ten& a, const Ten* b, const float* w, const float* wd, int taps, const float* bias, int Cout, const Ten* res) {
    Ten o = make(a.H, a.W, Cout);
    if (!emit) return o;
    ConvArgs f;
    memset(&f, 0, sizeof(f));
    f.s0 = a.v; f.C0 = a.C; f.s1 = b ? b->v : nullptr; f.C1 = b ? b->C : 0;
    f.w = w; f.bias = bias; f.res = res ? res->v : nullptr;
    f.d0 = o.v; f.D0 = Cout; f.d1 = nullptr; f.D1 = 0; f.accumulate = 0;
    f.N = N; f.H = a.H; f.W = a.W; f.taps = taps;
    F->push_back([f](cudaStream_t s) { return launch_conv(f, s); });
    ConvArgs g;
    memset(&g, 0, sizeof(g));
    g.s0 = o.g; g.C0 = Cout; g.w = wd;
    g.d0 = a.g; g.D0 = a.C; g.d1 = b ? b->g : nullptr; g.D1 = b ? b->C : 0; g.accumulate = 1;
    g.N = N; g.H = a.H; g.W = a.W; g.taps = taps;
    float* rg = res ? res->g : nullptr;
    const float* og = o.g;
    const size_t n = o.numel();
    B.push_back([g, rg, og, n](cudaStream_t s) {
      if (rg) {
        add_into_kernel<<<ew_grid(n), 256, 0, s>>>(rg, og, n);
        KDIP_LAUNCH_CHECK();
      }
      return launch_conv(g, s);
    });
    return o;
  }

  // out = act(GroupNorm32(concat(a, b)) * (1 + scale) + shift)
  Ten gn(const Ten& a, const Ten* b, const float* gamma, const float* beta, int film_off, int silu) {
    const int C = a.C + (b ? b->C : 0), P = a.H * a.W, n_ = N, R = e->R;
    Ten o = make(a.H, a.W, C);
    float* mr = scratch((size_t)N * 64);
    float* ab = scratch((size_t)N * C * 2);
    double* red = (double*)scratch((size_t)N * 64 * 2);
    if (!emit) return o;
    const float *s0 = a.v, *s1 = b ? b->v : nullptr;
    const int C0 = a.C, C1 = b ? b->C : 0;
    const float* fl = film_off >= 0 ? film : nullptr;
    const size_t total = o.numel();
    float* ov = o.v;
    F->push_back([=](cudaStream_t s) {
      gn_stats_kernel<<<dim3(32, n_), 256, 0, s>>>(s0, C0, s1, C1, P, mr);
      KDIP_LAUNCH_CHECK();
      gn_coef_kernel<<<ew_grid((size_t)n_ * C), 256, 0, s>>>(mr, gamma, beta, fl, R, film_off, C, ab, n_);
      KDIP_LAUNCH_CHECK();
      gn_apply_f32_kernel<<<ew_grid(total), 256, 0, s>>>(s0, C0, s1, C1, P, ab, silu, ov, total);
      KDIP_LAUNCH_CHECK();
      return KDIP_OK;
    });
    const float* og = o.g;
    float *g0 = a.g, *g1 = b ? b->g : nullptr;
    B.push_back([=](cudaStream_t s) {
      gn_bwd_stats_kernel<<<dim3(32, n_), 256, 0, s>>>(s0, C0, s1, C1, P, ab, mr, silu, og, red);
      KDIP_LAUNCH_CHECK();
      gn_bwd_apply_f32_kernel<<<ew_grid(total), 256, 0, s>>>(s0, C0, s1, C1, P, ab, mr, red, silu, og, g0, g1, total);
      KDIP_LAUNCH_CHECK();
      return KDIP_OK;
    });
    return o;
  }

  Ten resample(const Ten& a, int updown) {
    if (updown == 0) return a;
    const int H = a.H, W = a.W, C = a.C;
    Ten o = updown == 1 ? make(H / 2, W / 2, C) : make(2 * H, 2 * W, C);
    if (!emit) return o;
    const float* av = a.v;
    float *ov = o.v, *ag = a.g;
    const float* og = o.g;
    const size_t to = o.numel(), ta = a.numel();
    if (updown == 1) {
      F->push_back([=](cudaStream_t s) { avgpool2_kernel<<<ew_grid(to), 256, 0, s>>>(av, H, W, C, ov, to); KDIP_LAUNCH_CHECK(); return KDIP_OK; });
      B.push_back([=](cudaStream_t s) { avgpool2_bwd_kernel<<<ew_grid(ta), 256, 0, s>>>(og, H, W, C, ag, ta); KDIP_LAUNCH_CHECK(); return KDIP_OK; });
    } else {
      F->push_back([=](cudaStream_t s) { upsample2_kernel<<<ew_grid(to), 256, 0, s>>>(av, H, W, C, ov, to); KDIP_LAUNCH_CHECK(); return KDIP_OK; });
      B.push_back([=](cudaStream_t s) { upsample2_bwd_kernel<<<ew_grid(ta), 256, 0, s>>>(og, H, W, C, ag, ta); KDIP_LAUNCH_CHECK(); return KDIP_OK; });
    }
    return o;
  }

  Ten attention(const Ten& qkv, int heads) {
    const int T = qkv.H * qkv.W, n_ = N;
    Ten o = make(qkv.H, qkv.W, heads * 64);
    float* lse = scratch((size_t)N * heads * T);
    float* Db = scratch((size_t)N * heads * T);
    if (!emit) return o;
    const float* qv = qkv.v;
    float *ov = o.v, *qg = qkv.g;
    const float* og = o.g;
    const size_t sm_f = (size_t)(T + 64 + 128 + 8) * 4, sm_q = (size_t)(T + 64 + 64 + 128 + 8) * 4, sm_kv = (size_t)(2 * T + 128 + 256) * 4;
    F->push_back([=](cudaStream_t s) {
      attn_fwd_f32_kernel<<<dim3(T, heads, n_), AT_THREADS, sm_f, s>>>(qv, T, heads, ov, lse);
      KDIP_LAUNCH_CHECK();
      return KDIP_OK;
    });
    B.push_back([=](cudaStream_t s) {
      attn_bwd_q_f32_kernel<<<dim3(T, heads, n_), AT_THREADS, sm_q, s>>>(qv, og, lse, T, heads, qg, Db);
      KDIP_LAUNCH_CHECK();
      attn_bwd_kv_f32_kernel<<<dim3(T, heads, n_), AT_THREADS, sm_kv, s>>>(qv, og, lse, Db, T, heads, qg);
      KDIP_LAUNCH_CHECK();
      return KDIP_OK;
    });
    return o;
  }
};

static int build(Fp32Engine* e, int N, void* ws, size_t ws_bytes, size_t* need) {
  const kdip_unet_arch& a = e->arch;
  const int mc = a.model_channels, ted = 4 * mc, S = a.image_size;
  // sizing pass first (value region size is needed to place the gradient region)
  size_t val_bytes = 0, grd_bytes = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const bool emit = pass == 1;
    if (emit && ws == nullptr) break;
    Builder b;
    b.e = e; b.N = N; b.emit = emit; b.F = &e->fwd;
    b.val.base = emit ? (char*)ws : nullptr;
    b.grd.base = emit ? (char*)ws + val_bytes : nullptr;
    if (emit) { e->fwd.clear(); e->bwd.clear(); }
    float* semb = b.scratch((size_t)N * ted);
    b.film = b.scratch((size_t)N * e->R);
    Ten xin = b.make(S, S, 3);
    if (emit) {
      Fp32Engine* ee = e;
      const int R = e->R;
      float* film = b.film;
      const float *tw1 = e->tw1, *tb1 = e->tb1, *tw2 = e->tw2, *tb2 = e->tb2;
      const size_t tot = xin.numel(), HW = (size_t)S * S;
      float* xv = xin.v;
      e->fwd.push_back([=](cudaStream_t s) {
        time_embed_f32_kernel<<<N, 512, (size_t)5 * mc * sizeof(float), s>>>(ee->io_t, mc, tw1, tb1, tw2, tb2, semb);
        KDIP_LAUNCH_CHECK();
        emb_proj_f32_kernel<<<(R + 7) / 8, 256, 0, s>>>(semb, N, ted, ee->wall, ee->ball, R, film);
        KDIP_LAUNCH_CHECK();
        nchw_to_nhwc_f32_kernel<<<ew_grid(tot), 256, 0, s>>>(ee->io_x, ee->io_xscale, 3, HW, xv, tot);
        KDIP_LAUNCH_CHECK();
        return KDIP_OK;
      });
      const float* xg = xin.g;
      b.B.push_back([=](cudaStream_t s) {
        nhwc_to_nchw_f32_kernel<<<ew_grid(tot), 256, 0, s>>>(xg, 3, HW, ee->io_grad, tot);
        KDIP_LAUNCH_CHECK();
        return KDIP_OK;
      });
    }
    std::vector<Ten> hs;
    Ten h;
    for (size_t i = 0; i < e->plan.size(); ++i) {
      const BlockDesc& bd = e->plan[i];
      if (bd.kind == 0) {
        h = b.conv(xin, nullptr, e->w_in, e->w_ind, 9, e->b_in, bd.cout, nullptr);
      } else if (bd.kind == 1) {
        const ResW& w = e->resw[i];
        const bool two = (bd.stage == 2 && bd.first_of_block);
        Ten skip;
        if (two) { skip = hs.back(); hs.pop_back(); }
        const Ten* s1 = two ? &skip : nullptr;
        if (h.C + (two ? skip.C : 0) != bd.cin) { set_error("fp32 unet plan: channel mismatch at %s", bd.prefix.c_str()); return KDIP_ESHAPE; }
        // in_layers: GN, SiLU, (resample), conv3x3   (unet.py:240-247)
        Ten a1 = b.gn(h, s1, w.g1, w.b1, -1, 1);
        Ten a1r = b.resample(a1, bd.updown);
        Ten h1 = b.conv(a1r, nullptr, w.w1, w.w1d, 9, w.bias1, bd.cout, nullptr);
        // out_layers: GN * (1 + scale) + shift, SiLU, conv3x3 (unet.py:248-253); skip (+resample) added (:257)
        Ten a2 = b.gn(h1, nullptr, w.g2, w.b2, w.film_off, 1);
        Ten out;
        if (bd.cin != bd.cout) {
          // skip_connection = conv1x1 on the (concatenated) block input; no resample together with a channel change in this family
          if (bd.updown != 0) { set_error("fp32 unet plan: resampling ResBlock with a channel change is unsupported"); return KDIP_ESHAPE; }
          Ten sk = b.conv(h, s1, w.ws, w.wsd, 1, w.bias_s, bd.cout, nullptr);
          out = b.conv(a2, nullptr, w.w2, w.w2d, 9, w.bias2, bd.cout, &sk);
        } else {
          if (two) { set_error("fp32 unet plan: identity skip over a concatenated input is unsupported"); return KDIP_ESHAPE; }
          Ten xr = b.resample(h, bd.updown);
          out = b.conv(a2, nullptr, w.w2, w.w2d, 9, w.bias2, bd.cout, &xr);
        }
        h = out;
      } else {
        const AttW& w = e->attw[i];
        const int c = bd.cin, heads = c / 64;
        Ten xn = b.gn(h, nullptr, w.g, w.b, -1, 0);
        Ten qkv = b.conv(xn, nullptr, w.wqkv, w.wqkvd, 1, w.bqkv, 3 * c, nullptr);
        Ten att = b.attention(qkv, heads);
        h = b.conv(att, nullptr, w.wproj, w.wprojd, 1, w.bproj, c, &h);
      }
      if (bd.stage == 0 && bd.last_of_block) hs.push_back(h);
    }
    if (!hs.empty()) { set_error("fp32 unet plan: skip stack not empty at the end"); return KDIP_EINVAL; }
    // head: GN, SiLU, conv3x3 -> 6 channels (unet.py:613-618); optional out_cov 1x1 on the pre-head feature (external.py:141)
    Ten hn = b.gn(h, nullptr, e->g_head, e->be_head, -1, 1);
    Ten o6 = b.conv(hn, nullptr, e->w_head, e->w_headd, 9, e->b_head, 6, nullptr);
    Ten cov6;
    if (e->has_cov) cov6 = b.make(S, S, 6, false);
    if (emit) {
      Fp32Engine* ee = e;
      const size_t tot = o6.numel(), HW = (size_t)S * S;
      const float* ov = o6.v;
      float* og = o6.g;
      ConvArgs f;
      memset(&f, 0, sizeof(f));
      if (e->has_cov) {
        f.s0 = h.v; f.C0 = h.C; f.w = e->w_cov; f.bias = e->b_cov; f.d0 = cov6.v; f.D0 = 6; f.N = N; f.H = S; f.W = S; f.taps = 1;
      }
      const float* cv = cov6.v;
      e->fwd.push_back([=](cudaStream_t s) {
        nhwc_to_nchw_f32_kernel<<<ew_grid(tot), 256, 0, s>>>(ov, 6, HW, ee->io_out, tot);
        KDIP_LAUNCH_CHECK();
        if (ee->io_cov) {
          int rc = launch_conv(f, s);
          if (rc) return rc;
          nhwc_to_nchw_f32_kernel<<<ew_grid(tot), 256, 0, s>>>(cv, 6, HW, ee->io_cov, tot);
          KDIP_LAUNCH_CHECK();
        }
        return KDIP_OK;
      });
      // first backward op: zero every gradient buffer, then land the seed
      b.B.push_back([=](cudaStream_t s) {
        KDIP_CUDA(cudaMemsetAsync(ee->grad_base, 0, ee->grad_bytes, s));
        nchw_to_nhwc_f32_kernel<<<ew_grid(tot), 256, 0, s>>>(ee->io_seed, nullptr, 6, HW, og, tot);
        KDIP_LAUNCH_CHECK();
        return KDIP_OK;
      });
      e->bwd.assign(b.B.rbegin(), b.B.rend());
      e->hlast = h;
      e->grad_base = (char*)ws + val_bytes;
      e->grad_bytes = b.grd.cur;
    }
    if (!emit) { val_bytes = b.val.cur; grd_bytes = b.grd.cur; }
  }
  *need = val_bytes + grd_bytes;
  if (ws != nullptr) {
    if (*need > ws_bytes) { set_error("unet(fp32): workspace too small: need %zu bytes, got %zu", *need, ws_bytes); return KDIP_ENOMEM; }
    e->planned_N = N; e->planned_ws = ws; e->planned_bytes = ws_bytes;
  }
  return KDIP_OK;
}

}  // namespace f32

// ---- interface used by unet.cu's C-ABI entry points ------------------------------------------------------------------
void fp32_destroy(Fp32Engine* e) {
  if (!e) return;
  for (void* p : e->owned) cudaFree(p);
  delete e;
}

int fp32_create(const kdip_unet_arch* arch, const std::map<std::string, std::pair<const float*, int64_t>>& src, Fp32Engine** out) {
  using namespace f32;
  Fp32Engine* e = new Fp32Engine();
  e->arch = *arch;
  build_block_plan(*arch, e->plan);
  cudaStream_t s = 0;
  auto fail = [&](int code) { fp32_destroy(e); return code; };
#define TRY(x) do { int _r = (x); if (_r != KDIP_OK) return fail(_r); } while (0)
  auto get = [&](const std::string& name, int64_t numel, const float** p) -> int {
    auto it = src.find(name);
    KDIP_REQUIRE(it != src.end(), KDIP_EINVAL, "unet_create: missing tensor '%s'", name.c_str());
    KDIP_REQUIRE(it->second.second == numel, KDIP_ESHAPE, "unet_create: tensor '%s' has %lld elements, expected %lld", name.c_str(),
                 (long long)it->second.second, (long long)numel);
    *p = it->second.first;
    return KDIP_OK;
  };
  auto keep = [&](const std::string& name, int64_t numel, const float** dst) -> int {
    const float* p;
    int r = get(name, numel, &p);
    if (r) return r;
    void* d;
    r = dev_alloc(e, (size_t)numel * 4, &d);
    if (r) return r;
    KDIP_CUDA(cudaMemcpyAsync(d, p, (size_t)numel * 4, cudaMemcpyDeviceToDevice, s));
    *dst = (const float*)d;
    return KDIP_OK;
  };
  // conv weight in both operator layouts
  auto pack2 = [&](const std::string& name, int O, int I, int taps, const float** fwd, const float** bwd) -> int {
    const float* p;
    int r = get(name, (int64_t)O * I * taps, &p);
    if (r) return r;
    for (int flip = 0; flip < 2; ++flip) {
      if (flip == 1 && bwd == nullptr) break;
      void* d;
      r = dev_alloc(e, (size_t)O * I * taps * 4, &d);
      if (r) return r;
      pack_f32_kernel<<<ew_grid((size_t)O * I * taps), 256, 0, s>>>(p, O, I, taps, flip, (float*)d);
      KDIP_LAUNCH_CHECK();
      *(flip ? bwd : fwd) = (const float*)d;
    }
    return KDIP_OK;
  };
  const int mc = arch->model_channels, ted = 4 * mc;
  TRY(keep("time_embed.0.weight", (int64_t)ted * mc, &e->tw1));
  TRY(keep("time_embed.0.bias", ted, &e->tb1));
  TRY(keep("time_embed.2.weight", (int64_t)ted * ted, &e->tw2));
  TRY(keep("time_embed.2.bias", ted, &e->tb2));
  int R = 0;
  for (auto& b : e->plan) if (b.kind == 1) R += 2 * b.cout;
  e->R = R;
  TRY(dev_alloc(e, (size_t)R * ted * 4, (void**)&e->wall));
  TRY(dev_alloc(e, (size_t)R * 4, (void**)&e->ball));
  e->resw.resize(e->plan.size());
  e->attw.resize(e->plan.size());
  int film_off = 0;
  for (size_t i = 0; i < e->plan.size(); ++i) {
    const BlockDesc& b = e->plan[i];
    const std::string& p = b.prefix;
    if (b.kind == 0) {
      TRY(pack2(p + ".weight", b.cout, 3, 9, &e->w_in, &e->w_ind));
      TRY(keep(p + ".bias", b.cout, &e->b_in));
    } else if (b.kind == 1) {
      ResW& w = e->resw[i];
      memset(&w, 0, sizeof(w));
      const int ci = b.cin, co = b.cout;
      TRY(keep(p + ".in_layers.0.weight", ci, &w.g1));
      TRY(keep(p + ".in_layers.0.bias", ci, &w.b1));
      TRY(keep(p + ".out_layers.0.weight", co, &w.g2));
      TRY(keep(p + ".out_layers.0.bias", co, &w.b2));
      TRY(pack2(p + ".in_layers.2.weight", co, ci, 9, &w.w1, &w.w1d));
      TRY(pack2(p + ".out_layers.3.weight", co, co, 9, &w.w2, &w.w2d));
      TRY(keep(p + ".in_layers.2.bias", co, &w.bias1));
      TRY(keep(p + ".out_layers.3.bias", co, &w.bias2));
      if (ci != co) {
        TRY(pack2(p + ".skip_connection.weight", co, ci, 1, &w.ws, &w.wsd));
        TRY(keep(p + ".skip_connection.bias", co, &w.bias_s));
      }
      w.film_off = film_off;
      const float *ew, *eb;
      TRY(get(p + ".emb_layers.1.weight", (int64_t)2 * co * ted, &ew));
      TRY(get(p + ".emb_layers.1.bias", 2 * co, &eb));
      KDIP_CUDA(cudaMemcpyAsync(e->wall + (size_t)film_off * ted, ew, (size_t)2 * co * ted * 4, cudaMemcpyDeviceToDevice, s));
      KDIP_CUDA(cudaMemcpyAsync(e->ball + film_off, eb, (size_t)2 * co * 4, cudaMemcpyDeviceToDevice, s));
      film_off += 2 * co;
    } else {
      AttW& w = e->attw[i];
      memset(&w, 0, sizeof(w));
      const int c = b.cin;
      TRY(keep(p + ".norm.weight", c, &w.g));
      TRY(keep(p + ".norm.bias", c, &w.b));
      TRY(pack2(p + ".qkv.weight", 3 * c, c, 1, &w.wqkv, &w.wqkvd));
      TRY(pack2(p + ".proj_out.weight", c, c, 1, &w.wproj, &w.wprojd));
      TRY(keep(p + ".qkv.bias", 3 * c, &w.bqkv));
      TRY(keep(p + ".proj_out.bias", c, &w.bproj));
    }
  }
  const int c0 = (int)(arch->channel_mult[0] * mc);
  TRY(keep("out.0.weight", c0, &e->g_head));
  TRY(keep("out.0.bias", c0, &e->be_head));
  TRY(pack2("out.2.weight", 6, c0, 9, &e->w_head, &e->w_headd));
  TRY(keep("out.2.bias", 6, &e->b_head));
  if (src.count("out_cov.weight")) {
    TRY(pack2("out_cov.weight", 6, c0, 1, &e->w_cov, nullptr));
    TRY(keep("out_cov.bias", 6, &e->b_cov));
    e->has_cov = true;
  }
  KDIP_CUDA(cudaStreamSynchronize(s));
#undef TRY
  *out = e;
  return KDIP_OK;
}

bool fp32_has_cov(const Fp32Engine* e) { return e->has_cov; }

int fp32_workspace_bytes(Fp32Engine* e, int N, size_t* bytes) { return f32::build(e, N, nullptr, 0, bytes); }

int fp32_prepare(Fp32Engine* e, int N, void* ws, size_t ws_bytes) {
  KDIP_REQUIRE(ws != nullptr && ((uintptr_t)ws % 256) == 0, KDIP_EALIGN, "unet: workspace must be 256-byte aligned");
  if (e->planned_N == N && e->planned_ws == ws && e->planned_bytes == ws_bytes) return KDIP_OK;
  e->planned_N = 0;
  size_t need = 0;
  return f32::build(e, N, ws, ws_bytes, &need);
}

int fp32_forward(Fp32Engine* e, const float* x, const float* x_scale, const float* t, int N, float* out, float* cov_out, void* ws,
                 size_t ws_bytes, cudaStream_t s) {
  int rc = fp32_prepare(e, N, ws, ws_bytes);
  if (rc) return rc;
  e->io_x = x; e->io_xscale = x_scale; e->io_t = t; e->io_out = out; e->io_cov = cov_out;
  for (auto& op : e->fwd) {
    rc = op(s);
    if (rc) return rc;
  }
  return KDIP_OK;
}

int fp32_vjp(Fp32Engine* e, const float* seed, int N, float* grad_x, void* ws, size_t ws_bytes, cudaStream_t s) {
  KDIP_REQUIRE(e->planned_N == N && e->planned_ws == ws && e->planned_bytes == ws_bytes, KDIP_EINVAL,
               "unet_vjp: must follow kdip_unet_forward with the same N and workspace (saved activations live there)");
  e->io_seed = seed; e->io_grad = grad_x;
  for (auto& op : e->bwd) {
    int rc = op(s);
    if (rc) return rc;
  }
  return KDIP_OK;
}

int fp32_feature(Fp32Engine* e, int N, float* feat, cudaStream_t s) {
  KDIP_REQUIRE(e->planned_N == N && e->hlast.v, KDIP_EINVAL, "unet_feature: must follow kdip_unet_forward with the same N");
  const size_t tot = e->hlast.numel();
  f32::nhwc_to_nchw_f32_kernel<<<f32::ew_grid(tot), 256, 0, s>>>(e->hlast.v, e->hlast.C, (size_t)e->hlast.H * e->hlast.W, feat, tot);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

}  // namespace kdip
