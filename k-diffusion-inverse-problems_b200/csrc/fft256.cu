// 256-point register FFTs for the spectral operators at the path's image size (S = 256): same three passes as fft.cu
// (rows real->half spectrum, columns forward + pointwise op + inverse, rows half spectrum->real with fused epilogue;
// condition/measurements.py:139-156,178-196, condition/diffpir_utils/utils_sisr.py:22-41, condition/condition.py:356-357,408-410),
// but each 256-point transform is 16 x 16: 16 threads hold 16 complex values each, a radix-16 butterfly in registers
// (two radix-4 layers), the W256 twiddles from a shared-memory table, ONE transposition through shared memory, a second radix-16.
// The radix-2 shared-memory kernels of fft.cu (8 barrier-separated stages per transform) remain for S < 256.
#include "kdip_common.cuh"
#include "fft.cuh"

namespace kdip {

static constexpr int F_THREADS = 256;     // 16 transforms x 16 threads
static constexpr int F_LD = 17;           // padded row of the per-transform 16 x 16 exchange tile
static constexpr int F_TILE = 16 * F_LD + 1;   // per-transform stride (odd: spreads the 16 transforms over the banks)

__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 c_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// 4-point DFT in place: forward W4 = -i, inverse W4 = +i
template <bool INV>
__device__ __forceinline__ void fft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = c_add(a, c), s1 = c_sub(a, c), s2 = c_add(b, d), s3 = c_sub(b, d);
  const float2 r = INV ? make_float2(-s3.y, s3.x) : make_float2(s3.y, -s3.x);   // -+ i * s3
  a = c_add(s0, s2);
  c = c_sub(s0, s2);
  b = c_add(s1, r);
  d = c_sub(s1, r);
}

// 16-point DFT, natural order in and out (n = 4a + b -> k = c + 4d; the register permutation is resolved at compile time)
template <bool INV>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
#pragma unroll
  for (int b = 0; b < 4; ++b) fft4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);
  // v[4c + b] *= W16^(b c), W16 = exp(-+ 2 pi i / 16)
  const float sg = INV ? 1.f : -1.f;
  const float2 w1 = make_float2(C1, sg * S1), w2 = make_float2(C2, sg * C2), w3 = make_float2(S1, sg * C1);
  const float2 w4 = make_float2(0.f, sg), w6 = make_float2(-C2, sg * C2), w9 = make_float2(-C1, -sg * S1);
  v[5] = c_mul(v[5], w1); v[6] = c_mul(v[6], w2); v[7] = c_mul(v[7], w3);
  v[9] = c_mul(v[9], w2); v[10] = c_mul(v[10], w4); v[11] = c_mul(v[11], w6);
  v[13] = c_mul(v[13], w3); v[14] = c_mul(v[14], w6); v[15] = c_mul(v[15], w9);
#pragma unroll
  for (int c = 0; c < 4; ++c) fft4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  // X[c + 4d] sits in v[4c + d]
  float2 t[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) t[k] = v[4 * (k & 3) + (k >> 2)];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = t[k];
}

// One 256-point transform spread over the 16 threads j = 0..15 of a group.  In: u[n1] = x[16 n1 + j].  Out: u[k2] = X[j + 16 k2].
// `tile` is this transform's exchange tile.  WARPSYNC = false: the group's threads sit in different warps - every thread of the CTA
// must call (two __syncthreads inside).  WARPSYNC = true: the group is half of one warp and the tile is private to it - only
// __syncwarp, the warps of a CTA run through their transforms independently; `tw` is then the [k1][j] table of make_tw256_kj
// (conflict-free: the 16 lanes read consecutive entries; the plain table's tw[j k1] walks collide gcd(k1, 16)-fold).
template <bool INV, bool WARPSYNC = false>
__device__ __forceinline__ void fft256(float2 (&u)[16], int j, float2* tile, const float2* tw) {
  fft16<INV>(u);                       // over n1: u[k1] = Y[k1][n2 = j]
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) {
    float2 w = WARPSYNC ? tw[k1 * 16 + j] : tw[j * k1];             // exp(-2 pi i j k1 / 256)
    if (INV) w.y = -w.y;
    u[k1] = c_mul(u[k1], w);
  }
  if (WARPSYNC) __syncwarp(); else __syncthreads();                  // the tile may still be read by the previous user
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) tile[k1 * F_LD + j] = u[k1];
  if (WARPSYNC) __syncwarp(); else __syncthreads();
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) u[n2] = tile[j * F_LD + n2];
  fft16<INV>(u);                       // over n2: u[k2] = X[j + 16 k2]
}

__device__ __forceinline__ void make_tw256(float2* tw) {
  float s, c;
  sincospif(2.0f * (float)threadIdx.x / 256.f, &s, &c);
  tw[threadIdx.x] = make_float2(c, -s);
  __syncthreads();
}
// tw[k1 * 16 + j] = exp(-2 pi i j k1 / 256) (threadIdx.x = k1 * 16 + j)
__device__ __forceinline__ void make_tw256_kj(float2* tw) {
  float s, c;
  sincospif(2.0f * (float)((threadIdx.x >> 4) * (threadIdx.x & 15)) / 256.f, &s, &c);
  tw[threadIdx.x] = make_float2(c, -s);
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
// rows, real -> half spectrum: 32 real rows (16 packed complex transforms) per CTA
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(F_THREADS) rows_r2c_256_kernel(const float* __restrict__ x, float2* __restrict__ out) {
  __shared__ float2 tw[256];
  extern __shared__ float2 tiles[];   // [16][F_TILE]
  constexpr int S = 256, Sh = 129;
  make_tw256_kj(tw);
  const int f = threadIdx.x >> 4, j = threadIdx.x & 15;
  const size_t row0 = (size_t)blockIdx.x * 32;
  const float* ra = x + (row0 + 2 * f) * S;
  const float* rb = ra + S;
  float2 u[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) u[n1] = make_float2(__ldg(ra + 16 * n1 + j), __ldg(rb + 16 * n1 + j));
  float2* tile = tiles + f * F_TILE;
  fft256<false, true>(u, j, tile, tw);
  __syncwarp();
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) tile[j * F_LD + k2] = u[k2];       // Z[j + 16 k2] at [j][k2]
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * Sh; i += F_THREADS) {
    const int t = i / Sh, k = i - t * Sh;
    const float2* tl = tiles + t * F_TILE;
    const int kc = (S - k) & (S - 1);
    const float2 z = tl[(k & 15) * F_LD + (k >> 4)];
    float2 zc = tl[(kc & 15) * F_LD + (kc >> 4)];
    zc.y = -zc.y;
    // Xa = (Z[k] + conj(Z[S-k]))/2 ; Xb = (Z[k] - conj(Z[S-k]))/(2i)
    const float2 xa = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y + zc.y));
    const float2 d = make_float2(0.5f * (z.x - zc.x), 0.5f * (z.y - zc.y));
    out[(row0 + 2 * t) * Sh + k] = xa;
    out[(row0 + 2 * t + 1) * Sh + k] = make_float2(d.y, -d.x);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// rows, half spectrum -> real with the fused epilogue out = alpha * res * (mul ? mul : 1) + beta * (add ? add : 0)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(F_THREADS) rows_c2r_256_kernel(const float2* __restrict__ in, float* __restrict__ out, float alpha,
                                                                 const float* __restrict__ mul, float beta,
                                                                 const float* __restrict__ add) {
  __shared__ float2 tw[256];
  extern __shared__ float2 tiles[];
  constexpr int S = 256, Sh = 129;
  make_tw256_kj(tw);
  const size_t row0 = (size_t)blockIdx.x * 32;
  // Each bin of the half spectrum is read ONCE: half-warp task (t, d) loads bins k = 16 d + j of the rows 2t, 2t+1 (128-byte
  // segments) and writes Z[k] = Xa + i Xb as well as its Hermitian image Z[S - k] = conj(Xa) + i conj(Xb); all sixteen loads of a
  // thread are issued before the first use.  (The first version walked k = 0..255 and fetched every bin twice, 32 loads per
  // thread at 100 registers: 63 % of its stall samples were L2 round trips.)
  {
    const int f = threadIdx.x >> 4, j = threadIdx.x & 15;
    float2 xa[8], xb[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int task = f + 16 * q, t = task >> 3, d = task & 7;
      const float2* pa = in + (row0 + 2 * t) * Sh + 16 * d + j;
      xa[q] = pa[0];
      xb[q] = pa[Sh];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int task = f + 16 * q, t = task >> 3, d = task & 7;
      const int k = 16 * d + j;
      float2* tl = tiles + t * F_TILE;
      tl[j * F_LD + d] = make_float2(xa[q].x - xb[q].y, xa[q].y + xb[q].x);      // element k = 16 n1 + n2 lives at [n2][n1]
      if (k != 0) {
        const int kc = S - k;
        tl[(kc & 15) * F_LD + (kc >> 4)] = make_float2(xa[q].x + xb[q].y, xb[q].x - xa[q].y);
      }
    }
    if (threadIdx.x < 16) {   // Nyquist bins
      const float2* pa = in + (row0 + 2 * threadIdx.x) * Sh + 128;
      const float2 a = pa[0], b = pa[Sh];
      tiles[threadIdx.x * F_TILE + 0 * F_LD + 8] = make_float2(a.x - b.y, a.y + b.x);
    }
  }
  __syncthreads();
  const int f = threadIdx.x >> 4, j = threadIdx.x & 15;
  float2* tile = tiles + f * F_TILE;
  float2 u[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) u[n1] = tile[j * F_LD + n1];
  fft256<true, true>(u, j, tile, tw);
  const size_t oa = (row0 + 2 * f) * S, ob = oa + S;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const int c = j + 16 * k2;
    float va = alpha * u[k2].x, vb = alpha * u[k2].y;
    if (mul) { va *= mul[oa + c]; vb *= mul[ob + c]; }
    if (add) { va += beta * add[oa + c]; vb += beta * add[ob + c]; }
    out[oa + c] = va;
    out[ob + c] = vb;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// columns: 16 spectrum columns of one plane per CTA; thread (j, c): column c, member j of that column's 16-thread group
// (tid = 16 j + c, so that a warp touches two 128-byte row segments of the [ky][kx] arrays)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(F_THREADS, 3) cols_256_kernel(const float2* __restrict__ in, float2* __restrict__ out, SpecOp op) {
  __shared__ float2 tw[256];
  extern __shared__ float2 tiles[];         // [16][F_TILE] exchange tiles, then the parked OTF values [256][16]
  constexpr int S = 256, Sh = 129, G = 9;   // column groups per plane
  float2* otf_s = tiles + 16 * F_TILE;
  make_tw256(tw);
  const int p = blockIdx.x / G, kx0 = (blockIdx.x - p * G) * 16;
  const int j = threadIdx.x >> 4, c = threadIdx.x & 15;
  const int kx = kx0 + c;
  const bool live = kx < Sh;
  const bool use_otf = op.mode == SPEC_MULT || op.mode == SPEC_DIV_CONJ;
  const float2* src = in + (size_t)p * S * Sh + kx;
  float2* tile = tiles + c * F_TILE;
  float2 u[16];
  // The sixteen OTF bins this thread will need after the forward transform (ky = j + 16 k2) are fetched together with its data and
  // parked in shared memory (private slots: no synchronisation): inside the filter loop their L2 round trips serialised.
  if (use_otf && live) {
    float2 fb[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) fb[k2] = __ldg(op.otf + (size_t)(j + 16 * k2) * Sh + kx);
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) u[n1] = src[(size_t)(16 * n1 + j) * Sh];
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) otf_s[(j + 16 * k2) * 16 + c] = fb[k2];
  } else {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) u[n1] = live ? src[(size_t)(16 * n1 + j) * Sh] : make_float2(0.f, 0.f);
  }
  fft256<false>(u, j, tile, tw);            // u[k2] = X[ky = j + 16 k2]
  if (op.mode != SPEC_FORWARD_ONLY) {
    const int img = p / op.planes_per_image;
    if (live) {
      const float theta = (op.mode == SPEC_DIV_CONJ || op.mode == SPEC_DIV_TABLE) ? op.theta[img] : 0.f;
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const size_t sidx = (size_t)(j + 16 * k2) * Sh + kx;
        float2 v = u[k2];
        if (op.mode == SPEC_MULT) {
          float2 m = otf_s[(j + 16 * k2) * 16 + c];
          if (op.conj_otf) m.y = -m.y;
          v = c_mul(v, m);
        } else if (op.mode == SPEC_DIV_CONJ) {
          const float2 fb = otf_s[(j + 16 * k2) * 16 + c];
          const float den = op.sigma_s2 + theta * (fb.x * fb.x + fb.y * fb.y);
          v = c_mul(make_float2(v.x / den, v.y / den), make_float2(fb.x, -fb.y));
        } else if (op.mode == SPEC_DIV_TABLE) {
          const float den = op.sigma_s2 + theta * op.table[sidx];
          v = make_float2(v.x / den, v.y / den);
        }
        u[k2] = v;
      }
    }
    // inverse along ky: the value for input index m = 16 k2 + j is already in u[n1 = k2] of thread n2 = j
    fft256<true>(u, j, tile, tw);
  }
  if (live) {
    float2* dst = out + (size_t)p * S * Sh + kx;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) dst[(size_t)(j + 16 * k2) * Sh] = u[k2];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused 2-D spectral filter, ONE launch per application: out = alpha * IFFT2( op( FFT2(x) ) ) * mul + beta * add.
// A cluster of 8 CTAs owns one plane (32 rows each); the row-transformed half spectrum [32][129] stays in each CTA's shared
// memory, the column pass gathers / scatters its columns through distributed shared memory (ld/st.shared::cluster), and the
// inverse row pass reads the local rows again: the plane is read once and written once, no spectrum ever goes to L2 / HBM
// (the three-pass version spends 46-63 % of its issue slots waiting on those L2 round trips, profiles/r2_ncu_fft.txt).
// Columns: CTA c, group f transforms kx = 16 c + f; the 129th column rides with column 0 - both are REAL sequences after a
// real row transform (DC and Nyquist), so z = c0 + i c128 is one complex transform, separated in the spectrum by the
// Hermitian split (partner bin 256 - ky lives in lane 16 - j: one shuffle per value), filtered with their own OTF columns and
// merged back before the inverse.
// ---------------------------------------------------------------------------------------------------------------------
static constexpr int SF_CL = 8;              // CTAs per plane
static constexpr int SF_ROWS = 256 / SF_CL;  // rows per CTA
static constexpr int SF_LD = 17;             // column buffer row: 16 owned columns + the Nyquist column (CTA 0); odd -> conflict-free column walks

__device__ __forceinline__ uint32_t sf_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float2 sf_ld(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sf_st(uint32_t addr, float2 v) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// the pointwise spectral op of SpecOp on one bin; fb = the OTF value of that bin
__device__ __forceinline__ float2 sf_apply(const SpecOp& op, float2 v, float2 fb, float theta) {
  if (op.mode == SPEC_MULT) {
    if (op.conj_otf) fb.y = -fb.y;
    return c_mul(v, fb);
  }
  // SPEC_DIV_CONJ
  const float den = op.sigma_s2 + theta * (fb.x * fb.x + fb.y * fb.y);
  return c_mul(make_float2(v.x / den, v.y / den), make_float2(fb.x, -fb.y));
}

// Data flow (all exchanges are 128-byte row segments, the only shape distributed shared memory moves at speed: a first version
// that gathered columns with one 8-byte remote load per thread ran at 100 us per application, slower than the three passes):
//   phase 1  row transforms of the CTA's 32 rows; every half-warp pushes 16 consecutive bins of a row to the CTA that owns those
//            columns: colbuf[R][0..15] of CTA kx / 16 (st.shared::cluster); bin 128 goes to colbuf[R][16] of CTA 0
//   phase 2  column transforms entirely in local shared memory (colbuf[16 n1 + j][f], row stride 17 -> conflict-free), in place
//   phase 3  every half-warp pulls 16 consecutive bins of one of the CTA's rows back (ld.shared::cluster), rebuilds the packed
//            two-row spectrum with its Hermitian half, inverse row transforms, epilogue.
#ifndef KDIP_SF_MINB
#define KDIP_SF_MINB 2
#endif
__global__ void __cluster_dims__(SF_CL, 1, 1) __launch_bounds__(F_THREADS, KDIP_SF_MINB)
spec_filter_256_kernel(const float* __restrict__ x, float* __restrict__ out, SpecOp op, float alpha, const float* __restrict__ mul,
                       float beta, const float* __restrict__ add) {
  __shared__ float2 tw[256];
  extern __shared__ float2 sm[];
  constexpr int S = 256, Sh = 129;
  float2* colbuf = sm;                        // [256][SF_LD]: this CTA's 16 columns (+ Nyquist column in CTA 0) of all rows
  float2* tiles = sm + S * SF_LD;             // [16][F_TILE]
  // every CTA of the cluster must be running before its shared memory is written remotely: arrive now, wait before the first push
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  make_tw256_kj(tw);
  const uint32_t c = cluster_ctarank();
  const int p = blockIdx.x / SF_CL;
  const int f = threadIdx.x >> 4, j = threadIdx.x & 15;
  float2* tile = tiles + f * F_TILE;
  const uint32_t colbuf_u32 = smem_u32(colbuf);
  float2 u[16];

  // ---- phase 1 ----
  {
    const size_t row0 = (size_t)p * S + (size_t)c * SF_ROWS;
    const float* ra = x + (row0 + 2 * f) * S;
    const float* rb = ra + S;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) u[n1] = make_float2(__ldg(ra + 16 * n1 + j), __ldg(rb + 16 * n1 + j));
    fft256<false, true>(u, j, tile, tw);
    __syncwarp();
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) tile[j * F_LD + k2] = u[k2];       // Z[j + 16 k2] at [j][k2]
    __syncthreads();                                                   // the push below reads every group's tile
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    // half-warp task (t = row pair, d = destination CTA): lane j owns bin k = 16 d + j of rows 2t, 2t+1
    for (int task = f; task < 16 * SF_CL; task += 16) {
      const int t = task >> 3, d = task & 7;
      const float2* tl = tiles + t * F_TILE;
      const int k = 16 * d + j, kc = (S - k) & (S - 1);
      const float2 z = tl[j * F_LD + d];
      float2 zc = tl[(kc & 15) * F_LD + (kc >> 4)];
      zc.y = -zc.y;
      const float2 dd = make_float2(0.5f * (z.x - zc.x), 0.5f * (z.y - zc.y));
      const int R = (int)c * SF_ROWS + 2 * t;
      const uint32_t a = sf_mapa(colbuf_u32 + (uint32_t)((R * SF_LD + j) * 8), (uint32_t)d);
      sf_st(a, make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y + zc.y)));
      sf_st(a + SF_LD * 8, make_float2(dd.y, -dd.x));
    }
    if (threadIdx.x < 16) {   // Nyquist bins (real) of the 32 rows -> column 16 of CTA 0
      const float2 z = tiles[threadIdx.x * F_TILE + 0 * F_LD + 8];    // Z[128] = [128 & 15][128 >> 4]
      const int R = (int)c * SF_ROWS + 2 * threadIdx.x;
      const uint32_t a = sf_mapa(colbuf_u32 + (uint32_t)((R * SF_LD + 16) * 8), 0u);
      sf_st(a, make_float2(z.x, 0.f));                 // Xa[128] = Re Z[128]
      sf_st(a + SF_LD * 8, make_float2(z.y, 0.f));     // Xb[128] = Im Z[128]
    }
  }
  // the OTF column of phase 2 is fetched before the barrier: sixteen independent L2 round trips overlap the wait (issued inside the
  // filter loop they serialised - 17 % of the kernel's stall samples sat on them)
  const int kx = 16 * (int)c + f;
  const bool packed = (kx == 0);              // column 0 carries column 128 in its imaginary part
  float2 fb[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) fb[k2] = __ldg(op.otf + (size_t)(j + 16 * k2) * Sh + kx);
  cluster_sync_all();

  // ---- phase 2: column kx = 16 c + f, local ----
  {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      u[n1] = colbuf[(16 * n1 + j) * SF_LD + f];
      if (packed) u[n1].y = colbuf[(16 * n1 + j) * SF_LD + 16].x;
    }
    fft256<false, true>(u, j, tile, tw);            // u[k2] = Z[ky = j + 16 k2]
    const int img = p / op.planes_per_image;
    const float theta = (op.mode == SPEC_DIV_CONJ) ? op.theta[img] : 0.f;
    if (c == 0 && threadIdx.x < 32) {
      // warp 0 of CTA 0: lanes 0-15 = the packed column pair.  Partner bin 256 - ky: lane (16 - j) & 15, register 15 - k2
      // (lane 0: its own register (16 - k2) & 15).
      float2 zm[16];
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const int src = (threadIdx.x & 16) | ((16 - j) & 15);
        zm[k2].x = __shfl_sync(0xffffffffu, u[15 - k2].x, src);
        zm[k2].y = __shfl_sync(0xffffffffu, u[15 - k2].y, src);
      }
      if (j == 0) {
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) zm[k2] = u[(16 - k2) & 15];
      }
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const int ky = j + 16 * k2;
        if (packed) {
          const float2 z = u[k2], zc = make_float2(zm[k2].x, -zm[k2].y);
          const float2 a0 = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y + zc.y));
          const float2 d = make_float2(0.5f * (z.x - zc.x), 0.5f * (z.y - zc.y));
          const float2 a1 = make_float2(d.y, -d.x);
          const float2 r0 = sf_apply(op, a0, fb[k2], theta);
          const float2 r1 = sf_apply(op, a1, __ldg(op.otf + (size_t)ky * Sh + 128), theta);
          u[k2] = make_float2(r0.x - r1.y, r0.y + r1.x);
        } else {
          u[k2] = sf_apply(op, u[k2], fb[k2], theta);
        }
      }
    } else {
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) u[k2] = sf_apply(op, u[k2], fb[k2], theta);
    }
    // inverse along ky: the value for input index m = 16 k2 + j is already in u[n1 = k2] of thread n2 = j
    fft256<true, true>(u, j, tile, tw);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      float2* dst = colbuf + (j + 16 * k2) * SF_LD;
      if (packed) {
        dst[0] = make_float2(u[k2].x, 0.f);
        dst[16] = make_float2(u[k2].y, 0.f);
      } else {
        dst[f] = u[k2];
      }
    }
  }
  cluster_sync_all();

  // ---- phase 3 ----
  {
    // half-warp task (t, d): bins k = 16 d + j of rows 2t, 2t+1 from CTA d; Z = Xa + i Xb at k and its Hermitian image at 256 - k.
    // All sixteen remote loads of a thread are issued before the first use.
    float2 xa[8], xb[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int task = f + 16 * q, t = task >> 3, d = task & 7;
      const int R = (int)c * SF_ROWS + 2 * t;
      const uint32_t a = sf_mapa(colbuf_u32 + (uint32_t)((R * SF_LD + j) * 8), (uint32_t)d);
      xa[q] = sf_ld(a);
      xb[q] = sf_ld(a + SF_LD * 8);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int task = f + 16 * q, t = task >> 3, d = task & 7;
      const int k = 16 * d + j;
      float2* tl = tiles + t * F_TILE;
      tl[j * F_LD + d] = make_float2(xa[q].x - xb[q].y, xa[q].y + xb[q].x);
      if (k != 0) {   // bin 256 - k: conj(Xa) + i conj(Xb)
        const int kc = S - k;
        tl[(kc & 15) * F_LD + (kc >> 4)] = make_float2(xa[q].x + xb[q].y, xb[q].x - xa[q].y);
      }
    }
    if (threadIdx.x < 16) {
      const int R = (int)c * SF_ROWS + 2 * threadIdx.x;
      const uint32_t a = sf_mapa(colbuf_u32 + (uint32_t)((R * SF_LD + 16) * 8), 0u);
      const float2 xa = sf_ld(a), xb = sf_ld(a + SF_LD * 8);
      tiles[threadIdx.x * F_TILE + 0 * F_LD + 8] = make_float2(xa.x - xb.y, xa.y + xb.x);
    }
    __syncthreads();
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) u[n1] = tile[j * F_LD + n1];
    fft256<true, true>(u, j, tile, tw);
    const size_t oa = ((size_t)p * S + (size_t)c * SF_ROWS + 2 * f) * S, ob = oa + S;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const int cc = j + 16 * k2;
      float va = alpha * u[k2].x, vb = alpha * u[k2].y;
      if (mul) { va *= mul[oa + cc]; vb *= mul[ob + cc]; }
      if (add) { va += beta * add[oa + cc]; vb += beta * add[ob + cc]; }
      out[oa + cc] = va;
      out[ob + cc] = vb;
    }
  }
  // a CTA must not exit while its column buffer can still be read by a peer's phase 3
  cluster_sync_all();
}

static constexpr size_t kTilesBytes = (size_t)16 * F_TILE * sizeof(float2);

int launch_rows_r2c_256(const float* x, float2* out, int planes, cudaStream_t s) {
  rows_r2c_256_kernel<<<(unsigned)((size_t)planes * 256 / 32), F_THREADS, kTilesBytes, s>>>(x, out);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}
int launch_rows_c2r_256(const float2* in, float* out, int planes, float alpha, const float* mul, float beta, const float* add, cudaStream_t s) {
  rows_c2r_256_kernel<<<(unsigned)((size_t)planes * 256 / 32), F_THREADS, kTilesBytes, s>>>(in, out, alpha, mul, beta, add);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}
// x -> out through the fused cluster kernel (SPEC_MULT / SPEC_DIV_CONJ only)
int launch_spec_filter_256(const float* x, float* out, int planes, const SpecOp& op, float alpha, const float* mul, float beta,
                           const float* add, cudaStream_t s) {
  static bool attr_set = false;
  const size_t smem = ((size_t)256 * SF_LD + (size_t)16 * F_TILE) * sizeof(float2);
  if (!attr_set) {
    KDIP_CUDA(cudaFuncSetAttribute(spec_filter_256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  spec_filter_256_kernel<<<(unsigned)(planes * SF_CL), F_THREADS, smem, s>>>(x, out, op, alpha, mul, beta, add);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}
int launch_cols_256(const float2* in, float2* out, int planes, const SpecOp& op, cudaStream_t s) {
  static bool attr_set = false;
  const size_t smem = kTilesBytes + (size_t)256 * 16 * sizeof(float2);
  if (!attr_set) {
    KDIP_CUDA(cudaFuncSetAttribute(cols_256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  cols_256_kernel<<<(unsigned)(planes * 9), F_THREADS, smem, s>>>(in, out, op);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

}  // namespace kdip
