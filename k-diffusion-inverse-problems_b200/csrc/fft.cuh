// Spectral building blocks (fft.cu) shared by the operator / solver entry points (operators.cu).
#pragma once
#include "kdip_common.cuh"

namespace kdip {

enum { SPEC_FORWARD_ONLY = 0, SPEC_INVERSE_ONLY = 1, SPEC_MULT = 2, SPEC_BLUR_CLOSED = 3, SPEC_RESIDUAL = 4 };

// pointwise op applied between the forward and inverse column FFTs (cols_kernel)
struct SpecOp {
  int mode;
  int planes_per_image;   // 3: theta / per-image scalars are indexed by plane / planes_per_image
  const float2* otf;      // [S][S/2+1] half spectrum of the PSF (FB of utils_sisr.pre_calculate)
  int conj_otf;           // SPEC_MULT: multiply by conj(FB) (A^T) instead of FB (A)
  const float2* fy;       // [planes][S][S/2+1] spectrum of the measurement
  const float* theta;     // [images] scalar x0 variance
  float sigma_s2;
};

int check_fft_size(int S, int planes);
// x real [planes][S][S] (optionally * premul elementwise) -> row-transformed half spectrum [planes][S][S/2+1]
int launch_rows_r2c(const float* x, const float* premul, float2* out, int planes, int S, cudaStream_t s);
// inverse of the above with epilogue out = alpha*res*(mul?mul:1) + beta*(add?add:0); alpha should carry 1/(S*S)
int launch_rows_c2r(const float2* in, float* out, int planes, int S, float alpha, const float* mul, float beta, const float* add,
                    cudaStream_t s);
int launch_cols(const float2* in, float2* out, int planes, int S, const SpecOp& op, cudaStream_t s);

}  // namespace kdip
