// C-ABI wrappers of the UNet building blocks (parity tests) and layout converters.
#include "unet_kernels.cuh"

using namespace kdip;

namespace kdip {
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, int HW, bf16* __restrict__ dst, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t np = i / C;
    const size_t n = np / HW, p = np % HW;
    dst[i] = __float2bfloat16(src[(n * C + c) * HW + p]);
  }
}
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ src, int C, int HW, float* __restrict__ dst, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i % HW;
    const size_t nc = i / HW;
    const size_t n = nc / C, c = nc % C;
    dst[i] = __bfloat162float(src[(n * HW + p) * C + c]);
  }
}
}  // namespace kdip

extern "C" int kdip_nchw_f32_to_nhwc_bf16(const float* src, int N, int C, int H, int W, void* dst, kdip_stream_t s) {
  size_t total = (size_t)N * C * H * W;
  nchw_to_nhwc_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)s>>>(src, C, H * W, (bf16*)dst, total);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}
extern "C" int kdip_nhwc_bf16_to_nchw_f32(const void* src, int N, int C, int H, int W, float* dst, kdip_stream_t s) {
  size_t total = (size_t)N * C * H * W;
  nhwc_to_nchw_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)s>>>((const bf16*)src, C, H * W, dst, total);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_layer_chan_stats(const void* x, int N, int P, int C, float* stats, kdip_stream_t s) {
  return launch_chan_stats((const bf16*)x, N, P, C, stats, (cudaStream_t)s);
}
extern "C" int kdip_layer_gn_finalize(const float* stats0, int C0, const float* stats1, int C1, int N, int P, const float* gamma,
                                      const float* beta, const float* film, int film_stride, int film_off, float* ab, float* mr,
                                      kdip_stream_t s) {
  return launch_gn_finalize(stats0, C0, stats1, C1, N, P, gamma, beta, film, film_stride, film_off, ab, mr, (cudaStream_t)s);
}
extern "C" int kdip_layer_gn_apply(const void* src0, int C0, const void* src1, int C1, int N, int H, int W, const float* ab,
                                   int act_silu, int resample, void* out, kdip_stream_t s) {
  return launch_gn_apply((const bf16*)src0, C0, (const bf16*)src1, C1, N, H, W, ab, act_silu, resample, (bf16*)out, (cudaStream_t)s);
}
extern "C" int kdip_layer_gn_bwd(const void* src0, int C0, const void* src1, int C1, int N, int H, int W, const float* ab,
                                 const float* mr, int act_silu, int resample, const void* gy, const void* extra, int extra_mode,
                                 float* red_zeroed, float* k_scratch, void* dst0, void* dst1, kdip_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  int rc = launch_gn_bwd_reduce((const bf16*)src0, C0, (const bf16*)src1, C1, N, H, W, ab, act_silu, resample, (const bf16*)gy, red_zeroed, st);
  if (rc) return rc;
  rc = launch_gn_bwd_finalize(red_zeroed, ab, mr, nullptr, N, C0 + C1, H * W, nullptr, 0, 0, k_scratch, st);
  if (rc) return rc;
  return launch_gn_bwd_apply((const bf16*)src0, C0, (const bf16*)src1, C1, N, H, W, ab, k_scratch, act_silu, resample, (const bf16*)gy,
                             (const bf16*)extra, extra_mode, (bf16*)dst0, (bf16*)dst1, st);
}
extern "C" int kdip_layer_conv_small_cin(const float* in, const float* in_scale, const float* w_oihw, const float* bias, int N,
                                         int O, int I, int flip, int H, int W, float* w_scratch, void* out, kdip_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  int rc = launch_pack_small(w_oihw, O, I, flip, w_scratch, st);
  if (rc) return rc;
  const int CIN = flip ? O : I, Cout = flip ? I : O;
  return launch_conv_small_cin(in, in_scale, w_scratch, bias, N, CIN, H, W, Cout, (bf16*)out, st);
}
extern "C" int kdip_layer_time_embed(const float* t, int N, int mc, const float* w1, const float* b1, const float* w2,
                                     const float* b2, float* semb, kdip_stream_t s) {
  return launch_time_embed(t, N, mc, w1, b1, w2, b2, semb, (cudaStream_t)s);
}
extern "C" int kdip_layer_emb_proj(const float* semb, int N, int ted, const float* wall, const float* ball, int R, float* out,
                                   kdip_stream_t s) {
  return launch_emb_proj(semb, nullptr, N, ted, wall, ball, R, out, (cudaStream_t)s);
}
extern "C" int kdip_layer_attention_fwd(const void* qkv, int N, int T, int heads, void* out, float* lse, kdip_stream_t s) {
  return launch_attention_fwd((const bf16*)qkv, N, T, heads, 64, (bf16*)out, lse, (cudaStream_t)s);
}
extern "C" int kdip_layer_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, int N, int T,
                                        int heads, void* dqkv, kdip_stream_t s) {
  return launch_attention_bwd((const bf16*)qkv, (const bf16*)out, (const bf16*)d_out, lse, N, T, heads, 64, (bf16*)dqkv, (cudaStream_t)s);
}
