// Spectral building blocks (fft.cu) shared by the operator / solver entry points (operators.cu).
#pragma once
#include "kdip_common.cuh"

namespace kdip {

// pointwise op applied between the forward and inverse column FFTs (cols_kernel)
enum {
  SPEC_FORWARD_ONLY = 0,  // forward column FFT only (OTF construction): out = full 2-D half spectrum
  SPEC_MULT = 2,          // v *= FB (A) or conj(FB) (A^T)                       measurements.py:141,152,180,192
  SPEC_DIV_CONJ = 3,      // v = v / (sigma_s^2 + theta[img]*|FB|^2) * conj(FB)    condition.py:357
  SPEC_DIV_TABLE = 5      // v = v / (sigma_s^2 + theta[img]*table)                condition.py:409-410 (table = invW)
};

struct SpecOp {
  int mode;
  int planes_per_image;   // 3: theta is indexed by plane / planes_per_image
  const float2* otf;      // [S][S/2+1] half spectrum of the PSF (FB of utils_sisr.pre_calculate)
  int conj_otf;           // SPEC_MULT: multiply by conj(FB) (A^T) instead of FB (A)
  const float* table;     // SPEC_DIV_TABLE: real [S][S/2+1]
  const float* theta;     // [images] scalar x0 variance (device)
  float sigma_s2;
};

int check_fft_size(int S, int planes);
// x real [planes][S][S] -> row-transformed half spectrum [planes][S][S/2+1]
int launch_rows_r2c(const float* x, float2* out, int planes, int S, cudaStream_t s);
// inverse of the above with epilogue out = alpha*res*(mul?mul:1) + beta*(add?add:0); alpha should carry 1/(S*S)
int launch_rows_c2r(const float2* in, float* out, int planes, int S, float alpha, const float* mul, float beta, const float* add,
                    cudaStream_t s);
int launch_cols(const float2* in, float2* out, int planes, int S, const SpecOp& op, cudaStream_t s);

// S = 256: register radix-16 x 16 transforms (fft256.cu); the launchers above dispatch to them
int launch_rows_r2c_256(const float* x, float2* out, int planes, cudaStream_t s);
int launch_rows_c2r_256(const float2* in, float* out, int planes, float alpha, const float* mul, float beta, const float* add, cudaStream_t s);
int launch_cols_256(const float2* in, float2* out, int planes, const SpecOp& op, cudaStream_t s);
// S = 256, SPEC_MULT / SPEC_DIV_CONJ: the whole application out = alpha * IFFT2(op(FFT2 x)) * mul + beta * add in ONE launch
// (8-CTA cluster per plane, spectrum resident in distributed shared memory)
int launch_spec_filter_256(const float* x, float* out, int planes, const SpecOp& op, float alpha, const float* mul, float beta,
                           const float* add, cudaStream_t s);
// dispatcher: the fused kernel when it applies (KDIP_FFT_FUSED=0 keeps the three passes), else rows_r2c -> cols(op) -> rows_c2r
// through the two caller-provided half-spectrum buffers
int launch_spec_filter(const float* x, float* out, int planes, int S, const SpecOp& op, float alpha, const float* mul, float beta,
                       const float* add, float2* specA, float2* specB, cudaStream_t s);

}  // namespace kdip
