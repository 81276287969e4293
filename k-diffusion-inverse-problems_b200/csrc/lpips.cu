// LPIPS (Learned Perceptual Image Patch Similarity, Zhang et al. 2018; the `lpips` package v0.1, net='vgg') - the perceptual
// metric the reference's sample scripts report next to PSNR / SSIM (sample_condition_openai.py:46,161, sample_condition_openai_v2.py:39,147,
// analytic_variance.py:30, condition/dps_utils/compute_metric.py:6).  The package is a third-party dependency that is not vendored
// in the reference tree; its published algorithm is restated here and in oracle/lpips_ref.py:
//   features = VGG16 conv stack (torchvision `features`), taps after relu1_2, relu2_2, relu3_3, relu4_3, relu5_3
//   d(x, y)  = sum_l mean_{h,w} sum_c w_lc ( f_l(x)/(|f_l(x)|_c + 1e-10) - f_l(y)/(|f_l(y)|_c + 1e-10) )^2
// The 3x3 convolutions run on the tcgen05 implicit-GEMM kernel (conv_gemm.cu) through the public conv-plan API; this file holds the
// three small kernels around them: ReLU, 2x2 max pooling, and the per-layer normalise / difference / 1x1 "lin" reduction.
#include "kdip_common.cuh"

namespace kdip {

__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t v) {
  // a bf16 is negative iff its sign bit is set (-0 -> +0 is harmless): clear each negative half
  const uint32_t lo = (v & 0x00008000u) ? 0u : (v & 0x0000FFFFu);
  const uint32_t hi = (v & 0x80000000u) ? 0u : (v & 0xFFFF0000u);
  return lo | hi;
}

__global__ void relu_bf16_kernel(uint4* __restrict__ x, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = x[i];
    v.x = relu_bf16x2(v.x); v.y = relu_bf16x2(v.y); v.z = relu_bf16x2(v.z); v.w = relu_bf16x2(v.w);
    x[i] = v;
  }
}

__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  const float2 fa = unpack_bf16(a), fb = unpack_bf16(b);
  return pack_bf16(fmaxf(fa.x, fb.x), fmaxf(fa.y, fb.y));
}
// bf16 NHWC [N,H,W,C] -> [N,H/2,W/2,C], max over 2x2 windows (nn.MaxPool2d(2, 2)); one thread per 8 channels of an output pixel
__global__ void maxpool2_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N, int H, int W, int C8) {
  const int Ho = H >> 1, Wo = W >> 1;
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t r = i / C8;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const size_t base = (((size_t)n * H + 2 * yo) * W + 2 * xo) * C8 + c;
    const uint4 a = in[base], b = in[base + C8], d = in[base + (size_t)W * C8], e = in[base + (size_t)W * C8 + C8];
    uint4 o;
    o.x = max_bf16x2(max_bf16x2(a.x, b.x), max_bf16x2(d.x, e.x));
    o.y = max_bf16x2(max_bf16x2(a.y, b.y), max_bf16x2(d.y, e.y));
    o.z = max_bf16x2(max_bf16x2(a.z, b.z), max_bf16x2(d.z, e.z));
    o.w = max_bf16x2(max_bf16x2(a.w, b.w), max_bf16x2(d.w, e.w));
    out[i] = o;
  }
}

// One warp per pixel: lanes walk the channels in 8-channel chunks (C a multiple of 8, <= 512: at most two chunks per lane), first
// the two channel norms, then the weighted squared difference of the unit-normalised features; a block adds its pixels' sum
// (scaled by 1 / HW) to out[n] in double precision.
__global__ void __launch_bounds__(256) lpips_layer_kernel(const uint4* __restrict__ f0, const uint4* __restrict__ f1, const float* __restrict__ w,
                                                          int HW, int C, double* __restrict__ out) {
  const int n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int C8 = C >> 3;
  float acc = 0.f;
  for (int p = blockIdx.x * nw + warp; p < HW; p += gridDim.x * nw) {
    const size_t base = ((size_t)n * HW + p) * C8;
    float a[2][8], b[2][8];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c8 = lane + 32 * q;
      if (c8 < C8) {
        const uint4 u = f0[base + c8], v = f1[base + c8];
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = unpack_bf16(uu[e]), y = unpack_bf16(vv[e]);
          a[q][2 * e] = x.x; a[q][2 * e + 1] = x.y; b[q][2 * e] = y.x; b[q][2 * e + 1] = y.y;
          s0 = fmaf(x.x, x.x, fmaf(x.y, x.y, s0));
          s1 = fmaf(y.x, y.x, fmaf(y.y, y.y, s1));
        }
      }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    s0 = __shfl_sync(0xffffffffu, s0, 0); s1 = __shfl_sync(0xffffffffu, s1, 0);
    const float i0 = 1.f / (sqrtf(s0) + 1e-10f), i1 = 1.f / (sqrtf(s1) + 1e-10f);
    float d = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c8 = lane + 32 * q;
      if (c8 < C8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float t = __fmul_rn(a[q][e], i0) - __fmul_rn(b[q][e], i1);   // both products rounded: identical inputs give exactly 0
          d = fmaf(__ldg(w + c8 * 8 + e) * t, t, d);
        }
      }
    }
    acc += d;
  }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += red[i];
    atomicAdd(out + n, (double)t / (double)HW);
  }
}

static inline int ew_blocks(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace kdip

using namespace kdip;

extern "C" int kdip_relu_bf16(void* x, size_t n, kdip_stream_t s) {
  KDIP_REQUIRE(x != nullptr && n % 8 == 0 && ((uintptr_t)x % 16) == 0, KDIP_EALIGN, "relu_bf16: need a 16-byte aligned tensor of a multiple of 8 elements");
  relu_bf16_kernel<<<ew_blocks(n / 8, 256), 256, 0, (cudaStream_t)s>>>(reinterpret_cast<uint4*>(x), n / 8);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_maxpool2_bf16(const void* in, void* out, int N, int H, int W, int C, kdip_stream_t s) {
  KDIP_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, KDIP_ESHAPE,
               "maxpool2_bf16: need even H, W and C a multiple of 8 (got %d x %d x %d)", H, W, C);
  KDIP_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0, KDIP_EALIGN, "maxpool2_bf16: tensors must be 16-byte aligned");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  maxpool2_bf16_kernel<<<ew_blocks(total, 256), 256, 0, (cudaStream_t)s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), N, H, W, C / 8);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_lpips_layer(const void* f0, const void* f1, const float* w, int N, int HW, int C, double* out, kdip_stream_t s) {
  KDIP_REQUIRE(f0 && f1 && w && out && N > 0 && HW > 0, KDIP_EINVAL, "lpips_layer: bad argument");
  KDIP_REQUIRE(C % 8 == 0 && C >= 8 && C <= 512, KDIP_ESHAPE, "lpips_layer: C=%d must be a multiple of 8, at most 512", C);
  KDIP_REQUIRE(((uintptr_t)f0 % 16) == 0 && ((uintptr_t)f1 % 16) == 0, KDIP_EALIGN, "lpips_layer: features must be 16-byte aligned");
  int bx = (HW + 7) / 8;
  const int cap = num_sms() * 8;
  if (bx > cap) bx = cap;
  lpips_layer_kernel<<<dim3(bx, N), 256, 0, (cudaStream_t)s>>>(reinterpret_cast<const uint4*>(f0), reinterpret_cast<const uint4*>(f1), w, HW, C, out);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}
