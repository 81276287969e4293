// Batched 2-D real FFT kernels for the spectral operators of the path: circular blur A / A^T
// (condition/measurements.py:139-156,178-196), OTF construction (condition/diffpir_utils/utils_sisr.py:22-41,79-96) and the
// closed-form mat solvers (condition/condition.py:356-357,408-410).  Replaces torch.fft / cuFFT + ATen complex pointwise.
//
// A real S x S plane (S = 16..256, power of two) is transformed as
//   rows:  two real rows packed into one complex length-S FFT in shared memory  -> half spectrum [S][S/2+1]
//   cols:  8 spectrum columns per CTA staged in shared memory (64-byte row segments), length-S complex FFT along y,
//          an optional POINTWISE SPECTRAL OP, and (optionally) the inverse column FFT — all in one kernel, so a complete
//          solve is rows_r2c -> cols(op) -> rows_c2r = 3 kernels, each streaming the plane once.
//   rows:  inverse complex-to-real with a fused epilogue  out = alpha * res * mul + beta * add.
// fp32 radix-2 in shared memory; twiddles from sincospif (exact argument reduction).
#include <stdlib.h>

#include "kdip_common.cuh"
#include "fft.cuh"

namespace kdip {

static constexpr int FFT_THREADS = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ int bitrev(int v, int logS) { return (int)(__brev((unsigned)v) >> (32 - logS)); }

// tw[k] = exp(-2*pi*i*k/S), k < S/2
__device__ __forceinline__ void make_twiddles(float2* tw, int S) {
  for (int k = threadIdx.x; k < S / 2; k += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)k / (float)S, &s, &c);
    tw[k] = make_float2(c, -s);
  }
}

// In-place radix-2 DIT FFTs of T sequences of length S stored at buf[t*ld + i] in BIT-REVERSED order on entry,
// natural order on exit.  inverse: conjugated twiddles (no scaling).  All threads of the block participate.
__device__ __forceinline__ void fft_smem(float2* buf, int ld, int T, int S, int logS, const float2* tw, bool inverse) {
  const int halfS = S >> 1;
  for (int s = 1; s <= logS; ++s) {
    const int half = 1 << (s - 1);
    const int tstep = S >> s;
    __syncthreads();
    for (int b = threadIdx.x; b < T * halfS; b += blockDim.x) {
      const int t = b / halfS, bb = b - t * halfS;
      const int j = bb & (half - 1);
      const int i0 = ((bb >> (s - 1)) << s) + j;
      float2 w = tw[j * tstep];
      if (inverse) w.y = -w.y;
      float2* p = buf + t * ld;
      const float2 u = p[i0], v = cmul(p[i0 + half], w);
      p[i0] = make_float2(u.x + v.x, u.y + v.y);
      p[i0 + half] = make_float2(u.x - v.x, u.y - v.y);
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
// rows, real -> half spectrum.  grid = P * S / (2T) blocks; block b handles rows [b*2T, b*2T + 2T) of the flattened [P*S] rows
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FFT_THREADS) rows_r2c_kernel(const float* __restrict__ x, float2* __restrict__ out, int S, int logS,
                                                               int T) {
  extern __shared__ float2 sm[];
  float2* tw = sm;               // [S/2]
  float2* buf = sm + S / 2;      // [T][S+1]
  const int ld = S + 1;
  const int Sh = S / 2 + 1;
  make_twiddles(tw, S);
  const size_t row0 = (size_t)blockIdx.x * 2 * T;
  for (int i = threadIdx.x; i < 2 * T * S; i += blockDim.x) {
    const int r = i / S, c = i - r * S;
    const float v = x[(row0 + r) * S + c];
    float* dst = reinterpret_cast<float*>(&buf[(r >> 1) * ld + bitrev(c, logS)]);
    dst[r & 1] = v;
  }
  fft_smem(buf, ld, T, S, logS, tw, false);
  for (int i = threadIdx.x; i < T * Sh; i += blockDim.x) {
    const int t = i / Sh, k = i - t * Sh;
    const float2 z = buf[t * ld + k], zc = cconj(buf[t * ld + ((S - k) & (S - 1))]);
    // Xa = (Z[k] + conj(Z[S-k]))/2 ; Xb = (Z[k] - conj(Z[S-k]))/(2i)
    const float2 xa = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y + zc.y));
    const float2 d = make_float2(0.5f * (z.x - zc.x), 0.5f * (z.y - zc.y));
    const float2 xb = make_float2(d.y, -d.x);
    out[(row0 + 2 * t) * Sh + k] = xa;
    out[(row0 + 2 * t + 1) * Sh + k] = xb;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// rows, half spectrum -> real, with fused epilogue  out = alpha * res * (mul ? mul : 1) + beta * (add ? add : 0)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FFT_THREADS) rows_c2r_kernel(const float2* __restrict__ in, float* __restrict__ out, int S, int logS,
                                                               int T, float alpha, const float* __restrict__ mul, float beta,
                                                               const float* __restrict__ add) {
  extern __shared__ float2 sm[];
  float2* tw = sm;
  float2* buf = sm + S / 2;
  const int ld = S + 1;
  const int Sh = S / 2 + 1;
  make_twiddles(tw, S);
  const size_t row0 = (size_t)blockIdx.x * 2 * T;
  for (int i = threadIdx.x; i < T * S; i += blockDim.x) {
    const int t = i / S, k = i - t * S;
    float2 xa, xb;
    if (k < Sh) {
      xa = in[(row0 + 2 * t) * Sh + k];
      xb = in[(row0 + 2 * t + 1) * Sh + k];
    } else {
      xa = cconj(in[(row0 + 2 * t) * Sh + (S - k)]);
      xb = cconj(in[(row0 + 2 * t + 1) * Sh + (S - k)]);
    }
    // Z = Xa + i*Xb
    buf[t * ld + bitrev(k, logS)] = make_float2(xa.x - xb.y, xa.y + xb.x);
  }
  fft_smem(buf, ld, T, S, logS, tw, true);
  for (int i = threadIdx.x; i < 2 * T * S; i += blockDim.x) {
    const int r = i / S, c = i - r * S;
    const float2 z = buf[(r >> 1) * ld + c];
    float v = alpha * ((r & 1) ? z.y : z.x);
    const size_t o = (row0 + r) * S + c;
    if (mul) v *= mul[o];
    if (add) v += beta * add[o];
    out[o] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// columns: forward FFT along y, then (unless SPEC_FORWARD_ONLY) a pointwise spectral op and the inverse FFT along y
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FFT_THREADS) cols_kernel(const float2* __restrict__ in, float2* __restrict__ out, int S, int logS,
                                                           SpecOp op) {
  extern __shared__ float2 sm[];
  const int CW = 8;
  const int ld = S + 1;
  float2* tw = sm;                     // [S/2]
  float2* bufA = sm + S / 2;           // [CW][S+1]
  float2* bufB = bufA + CW * ld;       // [CW][S+1]
  const int Sh = S / 2 + 1;
  const int groups = (Sh + CW - 1) / CW;
  const int p = blockIdx.x / groups;               // plane
  const int kx0 = (blockIdx.x - p * groups) * CW;
  make_twiddles(tw, S);
  const bool inv = op.mode != SPEC_FORWARD_ONLY;
  const float2* src = in + (size_t)p * S * Sh;
  for (int i = threadIdx.x; i < S * CW; i += blockDim.x) {
    const int ky = i / CW, c = i - ky * CW;
    const int kx = kx0 + c;
    float2 v = make_float2(0.f, 0.f);
    if (kx < Sh) v = src[(size_t)ky * Sh + kx];
    bufA[c * ld + bitrev(ky, logS)] = v;
  }
  fft_smem(bufA, ld, CW, S, logS, tw, false);
  // pointwise op in natural order
  float2* res = bufA;
  if (inv) {
    const int img = p / op.planes_per_image;
    for (int i = threadIdx.x; i < S * CW; i += blockDim.x) {
      const int ky = i / CW, c = i - ky * CW;
      const int kx = kx0 + c;
      float2 v = bufA[c * ld + ky];
      if (kx < Sh) {
        const size_t sidx = (size_t)ky * Sh + kx;
        if (op.mode == SPEC_MULT) {
          float2 m = op.otf[sidx];
          if (op.conj_otf) m.y = -m.y;
          v = cmul(v, m);
        } else if (op.mode == SPEC_DIV_CONJ) {
          // fft2(r) / (sigma_s^2 + theta*|FB|^2) * conj(FB)                         condition/condition.py:357
          const float2 fb = op.otf[sidx];
          const float den = op.sigma_s2 + op.theta[img] * (fb.x * fb.x + fb.y * fb.y);
          v = cmul(make_float2(v.x / den, v.y / den), cconj(fb));
        } else if (op.mode == SPEC_DIV_TABLE) {
          // fft2(r) / (sigma_s^2 + theta*invW)                                       condition/condition.py:409-410
          const float den = op.sigma_s2 + op.theta[img] * op.table[sidx];
          v = make_float2(v.x / den, v.y / den);
        }
      }
      bufB[c * ld + bitrev(ky, logS)] = v;
    }
    fft_smem(bufB, ld, CW, S, logS, tw, true);
    res = bufB;
  }
  float2* dst = out + (size_t)p * S * Sh;
  for (int i = threadIdx.x; i < S * CW; i += blockDim.x) {
    const int ky = i / CW, c = i - ky * CW;
    const int kx = kx0 + c;
    if (kx < Sh) dst[(size_t)ky * Sh + kx] = res[c * ld + ky];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------------
static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
// complex FFTs per CTA (2T real rows each); shrunk until it divides the row count
static int rows_T(int S, int planes) {
  int T = S >= 256 ? 8 : (S >= 64 ? 16 : 32);
  while (T > 1 && ((size_t)planes * S) % (2 * T) != 0) T >>= 1;
  return T;
}

int check_fft_size(int S, int planes) {
  KDIP_REQUIRE(S >= 16 && S <= 256 && (S & (S - 1)) == 0, KDIP_ESHAPE, "fft: size %d must be a power of two in [16, 256]", S);
  KDIP_REQUIRE(planes > 0, KDIP_ESHAPE, "fft: no planes");
  return KDIP_OK;
}

int launch_rows_r2c(const float* x, float2* out, int planes, int S, cudaStream_t s) {
  int rc = check_fft_size(S, planes);
  if (rc) return rc;
  if (S == 256 && !getenv("KDIP_FFT_RADIX2")) return launch_rows_r2c_256(x, out, planes, s);
  const int T = rows_T(S, planes);
  const size_t smem = (size_t)(S / 2 + T * (S + 1)) * sizeof(float2);
  rows_r2c_kernel<<<(unsigned)((size_t)planes * S / (2 * T)), FFT_THREADS, smem, s>>>(x, out, S, ilog2(S), T);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

int launch_rows_c2r(const float2* in, float* out, int planes, int S, float alpha, const float* mul, float beta, const float* add,
                    cudaStream_t s) {
  int rc = check_fft_size(S, planes);
  if (rc) return rc;
  if (S == 256 && !getenv("KDIP_FFT_RADIX2")) return launch_rows_c2r_256(in, out, planes, alpha, mul, beta, add, s);
  const int T = rows_T(S, planes);
  const size_t smem = (size_t)(S / 2 + T * (S + 1)) * sizeof(float2);
  rows_c2r_kernel<<<(unsigned)((size_t)planes * S / (2 * T)), FFT_THREADS, smem, s>>>(in, out, S, ilog2(S), T, alpha, mul, beta, add);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

int launch_cols(const float2* in, float2* out, int planes, int S, const SpecOp& op, cudaStream_t s) {
  int rc = check_fft_size(S, planes);
  if (rc) return rc;
  if (S == 256 && !getenv("KDIP_FFT_RADIX2")) return launch_cols_256(in, out, planes, op, s);
  const int Sh = S / 2 + 1, groups = (Sh + 7) / 8;
  const size_t smem = (size_t)(S / 2 + 2 * 8 * (S + 1)) * sizeof(float2);
  cols_kernel<<<(unsigned)(planes * groups), FFT_THREADS, smem, s>>>(in, out, S, ilog2(S), op);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

int launch_spec_filter(const float* x, float* out, int planes, int S, const SpecOp& op, float alpha, const float* mul, float beta,
                       const float* add, float2* specA, float2* specB, cudaStream_t s) {
  // Measured on B200 (tools/time_fft.py, one application, us): planes 3 / 12 / 24 / 48 / 72 / 96 / 192: fused 18.7 / 19.4 / 23.8 / 42.4 /
  // 58.6 / 64.8 / 125.5, three passes 29.2 / 29.0 / 29.8 / 37.4 / 48.4 / 62.3 / 116.1.  One cluster of 8 CTAs per plane: 37 clusters are
  // co-resident, so up to one wave of planes the single launch wins (latency-bound regime: configs[0], B <= 12); beyond it both run
  // at ~1 instruction per clock per SM (sync- and latency-bound register FFTs, profiles/r2_ncu_fft.txt) and the three passes, whose
  // CTAs are independent, pack the SMs slightly better.  KDIP_FFT_FUSED=1 / 0 forces one or the other.
  const int fused_mode = getenv("KDIP_FFT_FUSED") ? atoi(getenv("KDIP_FFT_FUSED")) : -1;
  const bool fused_on = !getenv("KDIP_FFT_RADIX2") && (fused_mode == 1 || (fused_mode < 0 && planes <= 36));
  if (S == 256 && fused_on && (op.mode == SPEC_MULT || op.mode == SPEC_DIV_CONJ)) {
    int rc = check_fft_size(S, planes);
    if (rc) return rc;
    return launch_spec_filter_256(x, out, planes, op, alpha, mul, beta, add, s);
  }
  int rc = launch_rows_r2c(x, specA, planes, S, s);
  if (rc) return rc;
  rc = launch_cols(specA, specB, planes, S, op, s);
  if (rc) return rc;
  return launch_rows_c2r(specB, out, planes, S, alpha, mul, beta, add, s);
}

}  // namespace kdip
