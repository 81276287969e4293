// Data front end and evaluation reductions either side of the sampling path.  All HBM-bound byte /
// fp32 work, one pass over the data, coalesced 16-byte accesses on the fp32 planes and 12-byte (4-pixel) accesses on the
// interleaved 8-bit side:
//   * images_u8_to_f32 / images_f32_to_u8: the dataset transform of sample_condition_openai.py:140-144 (torchvision ToTensor =
//     u8 / 255, then x * 2 - 1) and k_diffusion/utils.py:24-31 to_pil_image ((clamp(x,-1,1) + 1) / 2, then torchvision's
//     mul(255).byte() truncation).  Bit-exact against those expressions (no FMA contraction, IEEE division).
//   * sqerr_sum: per-image sum of squared differences in fp64 - the reduction of analytic_variance.py:129 (raw) and of PSNR
//     (sample_condition_openai.py:41-44: both images mapped through to_eval = (x / 2 + 0.5).clip(0, 1) first).
//   * denoise_sqerr: analytic_variance.py:128-129 fused: hat_x0 = x_noised + eps * c_out (c_out = -sigma, k_diffusion/external.py
//     :97-115) and sum (x0 - hat_x0)^2 without materialising hat_x0.
//   * ssim_sum: structural_similarity(channel_axis=0, data_range=1) of sample_condition_openai.py:45 (scikit-image defaults:
//     7x7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03, mean over the window-valid interior), per image.
// NVCC_FLAGS: -fmad=false
#include "kdip_common.cuh"

namespace kdip {

#define IMG_THREADS 256

static inline dim3 img_grid(size_t per_image_items, int B) {
  size_t gx = (per_image_items + IMG_THREADS - 1) / IMG_THREADS;
  size_t cap = ((size_t)num_sms() * 16 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)B);
}

// one thread = 4 consecutive pixels: 12 interleaved bytes <-> one float4 per channel plane
__global__ void u8_to_f32_kernel(const uint32_t* __restrict__ src, float* __restrict__ dst, size_t HW4) {
  const size_t b = blockIdx.y;
  const uint32_t* sb = src + b * HW4 * 3;
  float* db = dst + b * HW4 * 12;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW4; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t w0 = __ldg(sb + 3 * i), w1 = __ldg(sb + 3 * i + 1), w2 = __ldg(sb + 3 * i + 2);
    uint8_t by[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      by[k] = (uint8_t)(w0 >> (8 * k));
      by[4 + k] = (uint8_t)(w1 >> (8 * k));
      by[8 + k] = (uint8_t)(w2 >> (8 * k));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float4 o;
      o.x = (float)by[c] / 255.f * 2.f - 1.f;
      o.y = (float)by[3 + c] / 255.f * 2.f - 1.f;
      o.z = (float)by[6 + c] / 255.f * 2.f - 1.f;
      o.w = (float)by[9 + c] / 255.f * 2.f - 1.f;
      reinterpret_cast<float4*>(db + (size_t)c * HW4 * 4)[i] = o;
    }
  }
}

__device__ __forceinline__ uint32_t to_byte(float x) {
  const float v = (fminf(fmaxf(x, -1.f), 1.f) + 1.f) / 2.f * 255.f;
  return (uint32_t)(uint8_t)v;   // truncation, as torch's .byte() on a value in [0, 255]
}

__global__ void f32_to_u8_kernel(const float* __restrict__ src, uint32_t* __restrict__ dst, size_t HW4) {
  const size_t b = blockIdx.y;
  const float* sb = src + b * HW4 * 12;
  uint32_t* db = dst + b * HW4 * 3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW4; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t by[12];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(sb + (size_t)c * HW4 * 4) + i);
      by[c] = to_byte(v.x); by[3 + c] = to_byte(v.y); by[6 + c] = to_byte(v.z); by[9 + c] = to_byte(v.w);
    }
    db[3 * i] = by[0] | (by[1] << 8) | (by[2] << 16) | (by[3] << 24);
    db[3 * i + 1] = by[4] | (by[5] << 8) | (by[6] << 16) | (by[7] << 24);
    db[3 * i + 2] = by[8] | (by[9] << 8) | (by[10] << 16) | (by[11] << 24);
  }
}

// block-wide fp64 sum -> one atomicAdd per block
__device__ __forceinline__ void block_add_double(double v, double* dst) {
  __shared__ double part[IMG_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < IMG_THREADS / 32 ? part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) atomicAdd(dst, t);
  }
}

__device__ __forceinline__ float to_eval(float x) { return fminf(fmaxf(x / 2.f + 0.5f, 0.f), 1.f); }

template <bool EVAL>
__global__ void sqerr_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ out, size_t CHW4) {
  const size_t img = blockIdx.y;
  const float4* pa = reinterpret_cast<const float4*>(a) + img * CHW4;
  const float4* pb = reinterpret_cast<const float4*>(b) + img * CHW4;
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < CHW4; i += (size_t)gridDim.x * blockDim.x) {
    float4 u = __ldg(pa + i), v = __ldg(pb + i);
    if (EVAL) {
      u.x = to_eval(u.x); u.y = to_eval(u.y); u.z = to_eval(u.z); u.w = to_eval(u.w);
      v.x = to_eval(v.x); v.y = to_eval(v.y); v.z = to_eval(v.z); v.w = to_eval(v.w);
      // skimage converts both images to float64 before subtracting
      const double d0 = (double)u.x - (double)v.x, d1 = (double)u.y - (double)v.y, d2 = (double)u.z - (double)v.z,
                   d3 = (double)u.w - (double)v.w;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    } else {
      // torch: (x0 - hat_x0).pow(2) in fp32, then the mean
      const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
      acc += (double)(d0 * d0) + (double)(d1 * d1) + (double)(d2 * d2) + (double)(d3 * d3);
    }
  }
  block_add_double(acc, out + img);
}

// hat_x0 = x_noised + eps * (-sigma); sum (x0 - hat_x0)^2.  unet_out is [B,6,H,W]: eps = channels 0..2.
__global__ void denoise_sqerr_kernel(const float* __restrict__ unet_out, const float* __restrict__ xn, const float* __restrict__ x0,
                                     const float* __restrict__ sigma, double* __restrict__ out, float* __restrict__ hat, size_t CHW4) {
  const size_t img = blockIdx.y;
  const float c_out = -sigma[img];
  const float4* pe = reinterpret_cast<const float4*>(unet_out) + img * CHW4 * 2;
  const float4* pn = reinterpret_cast<const float4*>(xn) + img * CHW4;
  const float4* p0 = reinterpret_cast<const float4*>(x0) + img * CHW4;
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < CHW4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 e = __ldg(pe + i), n = __ldg(pn + i), t = __ldg(p0 + i);
    float4 h;
    h.x = n.x + e.x * c_out; h.y = n.y + e.y * c_out; h.z = n.z + e.z * c_out; h.w = n.w + e.w * c_out;
    if (hat) reinterpret_cast<float4*>(hat)[img * CHW4 + i] = h;
    const float d0 = t.x - h.x, d1 = t.y - h.y, d2 = t.z - h.z, d3 = t.w - h.w;
    acc += (double)(d0 * d0) + (double)(d1 * d1) + (double)(d2 * d2) + (double)(d3 * d3);
  }
  block_add_double(acc, out + img);
}

// SSIM: one block = a 16x16 tile of window centres of one (image, channel); 22x22 input patch of both images in shared memory,
// horizontal 7-sums of (x, y, xx, yy, xy) in fp64, then the vertical pass + the SSIM map value per centre.
#define SS_T 16
#define SS_W 7
#define SS_P (SS_T + SS_W - 1)
__global__ void ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ out, int H, int W) {
  __shared__ float sa[SS_P][SS_P + 1], sb[SS_P][SS_P + 1];
  __shared__ double hs[5][SS_P][SS_T];
  const int tiles_x = (W - SS_W + 1 + SS_T - 1) / SS_T;
  const int ty0 = (blockIdx.x / tiles_x) * SS_T, tx0 = (blockIdx.x % tiles_x) * SS_T;   // top-left of the patch = first window origin
  const int img = blockIdx.z, ch = blockIdx.y;
  const float* pa = a + ((size_t)img * 3 + ch) * H * W;
  const float* pb = b + ((size_t)img * 3 + ch) * H * W;
  for (int i = threadIdx.x; i < SS_P * SS_P; i += blockDim.x) {
    const int r = i / SS_P, c = i % SS_P, y = ty0 + r, x = tx0 + c;
    const bool in = y < H && x < W;
    sa[r][c] = in ? to_eval(__ldg(pa + (size_t)y * W + x)) : 0.f;
    sb[r][c] = in ? to_eval(__ldg(pb + (size_t)y * W + x)) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SS_P * SS_T; i += blockDim.x) {
    const int r = i / SS_T, c = i % SS_T;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
    for (int k = 0; k < SS_W; ++k) {
      const double u = sa[r][c + k], v = sb[r][c + k];
      s0 += u; s1 += v; s2 += u * u; s3 += v * v; s4 += u * v;
    }
    hs[0][r][c] = s0; hs[1][r][c] = s1; hs[2][r][c] = s2; hs[3][r][c] = s3; hs[4][r][c] = s4;
  }
  __syncthreads();
  const int ny = H - SS_W + 1, nx = W - SS_W + 1;      // number of valid window origins per axis
  const double NP = SS_W * SS_W, cov_norm = NP / (NP - 1.0), C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;
  double acc = 0.0;
  for (int i = threadIdx.x; i < SS_T * SS_T; i += blockDim.x) {
    const int r = i / SS_T, c = i % SS_T;
    if (ty0 + r >= ny || tx0 + c >= nx) continue;
    double s[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      double t = 0;
#pragma unroll
      for (int k = 0; k < SS_W; ++k) t += hs[q][r + k][c];
      s[q] = t / NP;
    }
    const double ux = s[0], uy = s[1];
    const double vx = cov_norm * (s[2] - ux * ux), vy = cov_norm * (s[3] - uy * uy), vxy = cov_norm * (s[4] - ux * uy);
    const double A1 = 2 * ux * uy + C1, A2 = 2 * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
    acc += (A1 * A2) / (B1 * B2);
  }
  block_add_double(acc, out + img);
}

}  // namespace kdip

using namespace kdip;

#define REQ_A16(p) KDIP_REQUIRE(((uintptr_t)(p) & 15) == 0, KDIP_EALIGN, #p " must be 16-byte aligned")

extern "C" int kdip_images_u8_to_f32(const uint8_t* src_hwc, float* dst_nchw, int B, int H, int W, kdip_stream_t s) {
  KDIP_REQUIRE(src_hwc && dst_nchw && B > 0 && H > 0 && W > 0, KDIP_EINVAL, "images_u8_to_f32: bad argument");
  KDIP_REQUIRE(((size_t)H * W) % 4 == 0, KDIP_ESHAPE, "images_u8_to_f32: H*W must be a multiple of 4");
  KDIP_REQUIRE(((uintptr_t)src_hwc & 3) == 0, KDIP_EALIGN, "images_u8_to_f32: src must be 4-byte aligned");
  REQ_A16(dst_nchw);
  const size_t HW4 = (size_t)H * W / 4;
  u8_to_f32_kernel<<<img_grid(HW4, B), IMG_THREADS, 0, (cudaStream_t)s>>>((const uint32_t*)src_hwc, dst_nchw, HW4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_images_f32_to_u8(const float* src_nchw, uint8_t* dst_hwc, int B, int H, int W, kdip_stream_t s) {
  KDIP_REQUIRE(src_nchw && dst_hwc && B > 0 && H > 0 && W > 0, KDIP_EINVAL, "images_f32_to_u8: bad argument");
  KDIP_REQUIRE(((size_t)H * W) % 4 == 0, KDIP_ESHAPE, "images_f32_to_u8: H*W must be a multiple of 4");
  KDIP_REQUIRE(((uintptr_t)dst_hwc & 3) == 0, KDIP_EALIGN, "images_f32_to_u8: dst must be 4-byte aligned");
  REQ_A16(src_nchw);
  const size_t HW4 = (size_t)H * W / 4;
  f32_to_u8_kernel<<<img_grid(HW4, B), IMG_THREADS, 0, (cudaStream_t)s>>>(src_nchw, (uint32_t*)dst_hwc, HW4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_sqerr_sum(const float* a, const float* b, int to_eval_first, double* out, int B, int CHW, kdip_stream_t s) {
  KDIP_REQUIRE(a && b && out && B > 0 && CHW > 0 && CHW % 4 == 0, KDIP_EINVAL, "sqerr_sum: bad argument (CHW must be a multiple of 4)");
  REQ_A16(a); REQ_A16(b);
  KDIP_CUDA(cudaMemsetAsync(out, 0, (size_t)B * sizeof(double), (cudaStream_t)s));
  const size_t n4 = (size_t)CHW / 4;
  if (to_eval_first) sqerr_kernel<true><<<img_grid(n4, B), IMG_THREADS, 0, (cudaStream_t)s>>>(a, b, out, n4);
  else sqerr_kernel<false><<<img_grid(n4, B), IMG_THREADS, 0, (cudaStream_t)s>>>(a, b, out, n4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_denoise_sqerr(const float* unet_out, const float* x_noised, const float* x0, const float* sigma, double* out,
                                  float* hat_x0, int B, int HW, kdip_stream_t s) {
  KDIP_REQUIRE(unet_out && x_noised && x0 && sigma && out && B > 0 && HW > 0 && HW % 4 == 0, KDIP_EINVAL,
               "denoise_sqerr: bad argument (HW must be a multiple of 4)");
  REQ_A16(unet_out); REQ_A16(x_noised); REQ_A16(x0); REQ_A16(hat_x0);
  KDIP_CUDA(cudaMemsetAsync(out, 0, (size_t)B * sizeof(double), (cudaStream_t)s));
  const size_t n4 = (size_t)3 * HW / 4;
  denoise_sqerr_kernel<<<img_grid(n4, B), IMG_THREADS, 0, (cudaStream_t)s>>>(unet_out, x_noised, x0, sigma, out, hat_x0, n4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_ssim_sum(const float* a, const float* b, double* out, int B, int H, int W, kdip_stream_t s) {
  KDIP_REQUIRE(a && b && out && B > 0 && H >= SS_W && W >= SS_W, KDIP_EINVAL, "ssim_sum: images must be at least 7x7");
  KDIP_CUDA(cudaMemsetAsync(out, 0, (size_t)B * sizeof(double), (cudaStream_t)s));
  const int tiles = ((W - SS_W + 1 + SS_T - 1) / SS_T) * ((H - SS_W + 1 + SS_T - 1) / SS_T);
  ssim_kernel<<<dim3(tiles, 3, B), IMG_THREADS, 0, (cudaStream_t)s>>>(a, b, out, H, W);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}
