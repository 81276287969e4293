// Launchers shared by the measurement-operator / mat-solver entry points (operators.cu): spectral blocks (fft.cu) and
// orthogonal transforms (transforms.cu).  All tensors fp32, planes = images * 3, each plane [S][S] row-major (NCHW).
#pragma once
#include "fft.cuh"

namespace kdip {

// Haar level-3 DWT packed like pywt.coeffs_to_array (condition/utils.py:116-132).  forward: out = mul .* DWT(x) (mul may be
// NULL; it is indexed modulo mul_planes planes so one map can serve a whole batch); inverse: out = IDWT(x).
int launch_dwt(const float* x, const float* mul, int mul_planes, float* out, int planes, int S, int inverse, cudaStream_t s);
// out = W (theta .* W^T x) (Haar level 3): one fused pass at S = 256, else forward * theta -> tmp -> inverse
int launch_dwt_cov(const float* x, const float* theta, int theta_planes, float* tmp, float* out, int planes, int S, cudaStream_t s);

// Orthonormal DCT-II over (C=3, H, W) of each image (condition/utils.py:91-103).  ws: dct_workspace_bytes(planes, S).
size_t dct_workspace_bytes(int planes, int S);
int launch_dct(const float* x, const float* mul, int mul_images, float* out, int images, int S, int inverse, void* ws,
               cudaStream_t s);

}  // namespace kdip
