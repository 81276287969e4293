// Weight packing: fp32 parameter tensors (the reference's state_dict layout, OIHW / OI / OI1) -> bf16 K-major slabs
// [taps*rows_pad][cols_pad] consumed by conv_gemm.cu.  flip_transpose builds the input-gradient ("dgrad") operator of a
// stride-1, pad-1 convolution: W'[tap'][ci][co] = W[co][ci][taps-1-tap'] (180-degree rotated, in/out swapped).
#include "kdip_common.cuh"

namespace kdip {

// Cin_total = input channels of the source tensor; only the sub-range [ci_off, ci_off + Cin) is packed (the 1x1 skip conv
// of an output-stage ResBlock is split by concat source, unet.py:222,662).
__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin_total, int ci_off, int Cin, int taps,
                                   int rows_pad, int cols_pad, int flip, __nv_bfloat16* __restrict__ dst) {
  const size_t total = (size_t)taps * rows_pad * cols_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % cols_pad);
    const int row = (int)((i / cols_pad) % rows_pad);
    const int tap = (int)(i / ((size_t)cols_pad * rows_pad));
    float v = 0.f;
    if (!flip) {
      if (row < Cout && col < Cin) v = w[((size_t)row * Cin_total + ci_off + col) * taps + tap];
    } else {
      if (row < Cin && col < Cout) v = w[((size_t)col * Cin_total + ci_off + row) * taps + (taps - 1 - tap)];
    }
    dst[i] = __float2bfloat16(v);
  }
}

int pack_weight_ex(const float* w, int Cout, int Cin_total, int ci_off, int Cin_sub, int taps, int rows_pad, int cols_pad,
                   int flip, void* dst, cudaStream_t s) {
  const int rows = flip ? Cin_sub : Cout, cols = flip ? Cout : Cin_sub;
  KDIP_REQUIRE(rows_pad >= rows && cols_pad >= cols, KDIP_ESHAPE, "pack_weight: padding smaller than tensor");
  size_t total = (size_t)taps * rows_pad * cols_pad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  pack_weight_kernel<<<blocks, 256, 0, s>>>(w, Cout, Cin_total, ci_off, Cin_sub, taps, rows_pad, cols_pad, flip,
                                            reinterpret_cast<__nv_bfloat16*>(dst));
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

}  // namespace kdip

extern "C" int kdip_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int taps, int rows_pad, int cols_pad,
                                     int flip_transpose, void* dst_bf16, kdip_stream_t s) {
  KDIP_REQUIRE(w_oihw && dst_bf16, KDIP_EINVAL, "pack_conv_weight: null pointer");
  KDIP_REQUIRE(taps == 1 || taps == 9, KDIP_EINVAL, "pack_conv_weight: taps must be 1 or 9");
  return kdip::pack_weight_ex(w_oihw, Cout, Cin, 0, Cin, taps, rows_pad, cols_pad, flip_transpose, dst_bf16, (cudaStream_t)s);
}
