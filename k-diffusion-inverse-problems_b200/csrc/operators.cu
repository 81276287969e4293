// Measurement operators A / A^T and the posterior-covariance "mat" solvers  v = A^T (sigma_s^2 I + A Sigma A^T)^-1 (y - A x0)
// of the guided-sampling path, as device-resident handles behind the C ABI (include/kdip.h).
//
// Reference: condition/measurements.py:86-244 (SuperResolution / MotionBlur / GaussialBlur / Inpainting operators),
// condition/diffpir_utils/utils_sisr.py:9-96 (p2o, splits, upsample, downsample, pre_calculate),
// condition/dps_utils/resizer.py:8-198 (antialiased bicubic Resizer), condition/condition.py:317-439 (mat solvers; the
// scipy-CG-on-the-host path of :325-346,359-384,412-437 becomes a batched on-device CG with per-image convergence).
//
// Everything here is HBM-bound fp32 work on [B*3] planes of S x S: spectral passes come from fft.cu (3 kernels per
// A / A^T application), transforms from transforms.cu; the kernels below are the gather / scatter / residual glue and
// the CG vector updates.  The caller provides all workspace (kdip_op_workspace_bytes).
// NVCC_FLAGS: -fmad=false
#include <functional>
#include <vector>

#include "ops.cuh"

namespace kdip {

#define OP_THREADS 256
static inline int grid1d(size_t n) {
  size_t b = (n + OP_THREADS - 1) / OP_THREADS;
  size_t cap = (size_t)num_sms() * 16;
  if (b > cap) b = cap;
  return (int)(b ? b : 1);
}
#define GS_LOOP(i, n) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

// ---- PSF -> canvas (utils_sisr.py:22-38: zero-pad to S x S, roll by -floor(k/2) on both axes) ------------------------
__global__ void psf_canvas_kernel(const float* __restrict__ psf, int k, int S, float* __restrict__ canvas) {
  GS_LOOP(i, (size_t)S * S) canvas[i] = 0.f;
}
__global__ void psf_place_kernel(const float* __restrict__ psf, int k, int S, float* __restrict__ canvas) {
  GS_LOOP(i, (size_t)k * k) {
    const int r = (int)(i / k), c = (int)(i % k);
    const int rr = ((r - k / 2) % S + S) % S, cc = ((c - k / 2) % S + S) % S;
    canvas[(size_t)rr * S + cc] = psf[i];
  }
}
// invW = mean over the sf x sf aliases of |FB|^2 (utils_sisr.py:9-19 splits + condition.py:409), half-spectrum table [s][s/2+1]
__global__ void invw_kernel(const float2* __restrict__ otf, int S, int sf, float* __restrict__ invW) {
  const int s = S / sf, sh = s / 2 + 1, Sh = S / 2 + 1;
  GS_LOOP(i, (size_t)s * sh) {
    const int u = (int)(i / sh), v = (int)(i % sh);
    float acc = 0.f;
    for (int a = 0; a < sf; ++a)
      for (int b = 0; b < sf; ++b) {
        int ky = u + s * a, kx = v + s * b;
        if (kx > S / 2) { kx = S - kx; ky = (S - ky) % S; }   // Hermitian symmetry of the real PSF's spectrum
        const float2 f = otf[(size_t)ky * Sh + kx];
        acc += f.x * f.x + f.y * f.y;
      }
    invW[i] = acc / (float)(sf * sf);
  }
}
// full complex FB [S][S] from the half spectrum (for the reference's `pre_calculated` attribute)
__global__ void otf_full_kernel(const float2* __restrict__ otf, int S, float2* __restrict__ full) {
  const int Sh = S / 2 + 1;
  GS_LOOP(i, (size_t)S * S) {
    const int ky = (int)(i / S), kx = (int)(i % S);
    float2 v;
    if (kx <= S / 2) v = otf[(size_t)ky * Sh + kx];
    else { v = otf[(size_t)((S - ky) % S) * Sh + (S - kx)]; v.y = -v.y; }
    full[i] = v;
  }
}

// batched variant: [planes][S][S/2+1] half spectra -> [planes][S][S] full spectra (Hermitian completion)
__global__ void spec_full_kernel(const float2* __restrict__ half, int S, float2* __restrict__ full, size_t total) {
  const int Sh = S / 2 + 1;
  GS_LOOP(i, total) {
    const int kx = (int)(i % S), ky = (int)((i / S) % S);
    const size_t p = i / ((size_t)S * S);
    const float2* h = half + p * S * Sh;
    float2 v;
    if (kx <= S / 2) v = h[(size_t)ky * Sh + kx];
    else { v = h[(size_t)((S - ky) % S) * Sh + (S - kx)]; v.y = -v.y; }
    full[i] = v;
  }
}

// ---- super-resolution glue (utils_sisr.py:44-61) ----------------------------------------------------------------------
// r[p][i][j] = y[p][i][j] - full[p][sf*i][sf*j]
__global__ void sr_residual_kernel(const float* __restrict__ y, const float* __restrict__ full, float* __restrict__ r, int s,
                                   int sf, size_t total) {
  const int S = s * sf;
  GS_LOOP(i, total) {
    const int j = (int)(i % s), ii = (int)((i / s) % s);
    const size_t p = i / ((size_t)s * s);
    r[i] = y[i] - full[(p * S + (size_t)ii * sf) * S + (size_t)j * sf];
  }
}
// up = zero-filled upsample of q
__global__ void upsample_zero_kernel(const float* __restrict__ q, float* __restrict__ up, int s, int sf, size_t total_full) {
  const int S = s * sf;
  GS_LOOP(i, total_full) {
    const int x = (int)(i % S), yy = (int)((i / S) % S);
    const size_t p = i / ((size_t)S * S);
    float v = 0.f;
    if (x % sf == 0 && yy % sf == 0) v = q[(p * s + yy / sf) * s + x / sf];
    up[i] = v;
  }
}
// q = sigma2*u + full[::sf, ::sf]
__global__ void sr_matvec_tail_kernel(const float* __restrict__ u, const float* __restrict__ full, float sigma2,
                                      float* __restrict__ q, int s, int sf, size_t total) {
  const int S = s * sf;
  GS_LOOP(i, total) {
    const int j = (int)(i % s), ii = (int)((i / s) % s);
    const size_t p = i / ((size_t)s * s);
    q[i] = sigma2 * u[i] + full[(p * S + (size_t)ii * sf) * S + (size_t)j * sf];
  }
}

// ---- Resizer (resizer.py:55-74): separable weighted gather, H pass then W pass; and its adjoint -------------------------
// out[p][o][x] = sum_t w[o][t] * in[p][idx[o][t]][x]           (rows: Sin -> So, width Wd)
__global__ void resize_rows_kernel(const float* __restrict__ in, const float* __restrict__ w, const int* __restrict__ idx, int taps,
                                   int Sin, int So, int Wd, float* __restrict__ out, size_t total) {
  GS_LOOP(i, total) {
    const int x = (int)(i % Wd), o = (int)((i / Wd) % So);
    const size_t p = i / ((size_t)Wd * So);
    float acc = 0.f;
    for (int t = 0; t < taps; ++t) acc += in[(p * Sin + idx[o * taps + t]) * Wd + x] * w[o * taps + t];
    out[i] = acc;
  }
}
// out[p][y][o] = sum_t w[o][t] * in[p][y][idx[o][t]]  (+ sigma*noise)   (cols: Sin -> So, Hd rows)
__global__ void resize_cols_kernel(const float* __restrict__ in, const float* __restrict__ w, const int* __restrict__ idx, int taps,
                                   int Sin, int So, int Hd, const float* __restrict__ noise, float sigma, float* __restrict__ out,
                                   size_t total) {
  GS_LOOP(i, total) {
    const int o = (int)(i % So);
    const size_t py = i / So;
    float acc = 0.f;
    for (int t = 0; t < taps; ++t) acc += in[py * Sin + idx[o * taps + t]] * w[o * taps + t];
    if (noise) acc = acc + sigma * noise[i];
    out[i] = acc;
  }
}
// adjoint along columns: out[p][y][x] = sum_{e in CSR(x)} w_e * g[p][y][o_e]     (So -> Sin)
__global__ void resize_cols_adj_kernel(const float* __restrict__ g, const int* __restrict__ ptr, const int* __restrict__ oidx,
                                       const float* __restrict__ w, int Sin, int So, float* __restrict__ out, size_t total) {
  GS_LOOP(i, total) {
    const int x = (int)(i % Sin);
    const size_t py = i / Sin;
    float acc = 0.f;
    for (int e = ptr[x]; e < ptr[x + 1]; ++e) acc += g[py * So + oidx[e]] * w[e];
    out[i] = acc;
  }
}
// adjoint along rows: out[p][y][x] = sum_{e in CSR(y)} w_e * t[p][o_e][x]
__global__ void resize_rows_adj_kernel(const float* __restrict__ t, const int* __restrict__ ptr, const int* __restrict__ oidx,
                                       const float* __restrict__ w, int Sin, int So, int Wd, float* __restrict__ out, size_t total) {
  GS_LOOP(i, total) {
    const int x = (int)(i % Wd), yy = (int)((i / Wd) % Sin);
    const size_t p = i / ((size_t)Wd * Sin);
    float acc = 0.f;
    for (int e = ptr[yy]; e < ptr[yy + 1]; ++e) acc += t[(p * So + oidx[e]) * Wd + x] * w[e];
    out[i] = acc;
  }
}

// ---- small vector kernels -------------------------------------------------------------------------------------------
__global__ void sub_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, size_t n) {
  GS_LOOP(i, n) o[i] = a[i] - b[i];
}
__global__ void copy_kernel(const float* __restrict__ a, float* __restrict__ o, size_t n) { GS_LOOP(i, n) o[i] = a[i]; }
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, size_t n) {
  GS_LOOP(i, n) o[i] = a[i] * b[i];
}
// r = mask*y - mask*x0  (mask shared by the batch, CHW elements)                                  condition.py:340
__global__ void mask_residual_kernel(const float* __restrict__ y, const float* __restrict__ x0, const float* __restrict__ mask,
                                     float* __restrict__ r, size_t CHW, size_t n) {
  GS_LOOP(i, n) { const float m = mask[i % CHW]; r[i] = m * y[i] - m * x0[i]; }
}
// q = sigma2*m + mask * t   (t = Sigma m)                                                         condition.py:338
__global__ void mask_matvec_kernel(const float* __restrict__ m, const float* __restrict__ t, const float* __restrict__ theta,
                                   const float* __restrict__ mask, float sigma2, float* __restrict__ q, size_t CHW, size_t n) {
  GS_LOOP(i, n) {
    const float tv = theta ? theta[i] * m[i] : t[i];
    q[i] = sigma2 * m[i] + mask[i % CHW] * tv;
  }
}
// y = x * mask                                                                                        measurements.py:215
__global__ void mask_mul_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ o, size_t CHW, size_t n) {
  GS_LOOP(i, n) o[i] = x[i] * mask[i % CHW];
}

// ---- batched CG (scipy.sparse.linalg.cg semantics: x0 = 0, no preconditioner, stop when ||r|| < tol*||b||) ---------------
static constexpr int CG_NBLK = 64;   // partial sums per image

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;   // valid in warp 0
}

// partial[b][blk] = sum over this block's slice of a*b (bvec may be NULL -> a*a)
__global__ void __launch_bounds__(OP_THREADS) dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ bvec,
                                                                  float* __restrict__ part, size_t n) {
  const size_t img = blockIdx.y;
  const float4* ap = reinterpret_cast<const float4*>(a + img * n);
  const float4* bp = bvec ? reinterpret_cast<const float4*>(bvec + img * n) : ap;
  float acc = 0.f;
  const size_t n4 = n >> 2;       // n is a multiple of 4 (3 S^2, S >= 16); 16-byte loads, 4 elements per thread and trip
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 u = ap[i], v = bp[i];
    acc += (u.x * v.x + u.y * v.y) + (u.z * v.z + u.w * v.w);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) part[img * CG_NBLK + blockIdx.x] = acc;
}

// Per-image CG scalars, all resident on the device.  rho / done are double-buffered by iteration parity so that the kernel that
// publishes iteration k's values (cg_p_kernel, one thread per image) never races with the CTAs of the same launch still reading
// iteration k-1's: alpha and beta are formed inside the vector kernels from the reduction partials - no scalar kernels, no host.
struct CgState {
  float *rho[2];      // [B] ||r||^2 entering the iteration
  int *done[2];       // [B] converged (frozen) flags
  float *atol2;       // [B] (tol ||b||)^2
  int *iters;         // [B]
};

__device__ __forceinline__ float sum_partials(const float* part, int img) {
  float s = 0.f;
  for (int i = 0; i < CG_NBLK; ++i) s += part[(size_t)img * CG_NBLK + i];
  return s;
}

// x = 0 ; p = r ; partial ||r||^2
__global__ void __launch_bounds__(OP_THREADS) cg_init_kernel(const float* __restrict__ r, float* __restrict__ x, float* __restrict__ p,
                                                              float* __restrict__ part, size_t n) {
  const size_t img = blockIdx.y;
  float acc = 0.f;
  const float4* r4 = reinterpret_cast<const float4*>(r + img * n);
  float4* x4 = reinterpret_cast<float4*>(x + img * n);
  float4* p4 = reinterpret_cast<float4*>(p + img * n);
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = r4[i];
    x4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    p4[i] = v;
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) part[img * CG_NBLK + blockIdx.x] = acc;
}
// after ||b||^2 partials: rho, tolerance, b == 0 -> x = 0 (scipy returns immediately)
__global__ void cg_begin_kernel(const float* __restrict__ part, CgState st, float tol, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float v = sum_partials(part, b);
  st.rho[0][b] = v;
  st.atol2[b] = tol * tol * v;
  st.done[0][b] = (v == 0.f) ? 1 : 0;
  st.iters[b] = 0;
}
// alpha = rho / (p.q) from the partials of the preceding dot; x += alpha p ; r -= alpha q ; partial ||r||^2 (frozen once converged)
__global__ void __launch_bounds__(OP_THREADS) cg_update_kernel(float* __restrict__ x, float* __restrict__ r, const float* __restrict__ p,
                                                                const float* __restrict__ q, CgState st, int cur,
                                                                const float* __restrict__ part_pq, float* __restrict__ part_rr, size_t n) {
  const size_t img = blockIdx.y;
  float acc = 0.f;
  if (!st.done[cur][img]) {
    const float alpha = st.rho[cur][img] / sum_partials(part_pq, (int)img);
    float4* x4 = reinterpret_cast<float4*>(x + img * n);
    float4* r4 = reinterpret_cast<float4*>(r + img * n);
    const float4* p4 = reinterpret_cast<const float4*>(p + img * n);
    const float4* q4 = reinterpret_cast<const float4*>(q + img * n);
    const size_t n4 = n >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      float4 xv = x4[i], rv = r4[i];
      const float4 pv = p4[i], qv = q4[i];
      xv.x = xv.x + alpha * pv.x; xv.y = xv.y + alpha * pv.y; xv.z = xv.z + alpha * pv.z; xv.w = xv.w + alpha * pv.w;
      rv.x = rv.x - alpha * qv.x; rv.y = rv.y - alpha * qv.y; rv.z = rv.z - alpha * qv.z; rv.w = rv.w - alpha * qv.w;
      x4[i] = xv;
      r4[i] = rv;
      acc += (rv.x * rv.x + rv.y * rv.y) + (rv.z * rv.z + rv.w * rv.w);
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) part_rr[img * CG_NBLK + blockIdx.x] = acc;
}
// beta = ||r_new||^2 / rho ; p = r + beta p ; block 0 of each image publishes rho / done / iters of the next iteration
__global__ void __launch_bounds__(OP_THREADS) cg_p_kernel(float* __restrict__ p, const float* __restrict__ r, CgState st, int cur,
                                                           const float* __restrict__ part_rr, size_t n) {
  const size_t img = blockIdx.y;
  const int nxt = cur ^ 1;
  if (st.done[cur][img]) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { st.done[nxt][img] = 1; st.rho[nxt][img] = st.rho[cur][img]; }
    return;
  }
  const float rr = sum_partials(part_rr, (int)img);
  const float beta = rr / st.rho[cur][img];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st.rho[nxt][img] = rr;
    st.iters[img] += 1;
    st.done[nxt][img] = (rr < st.atol2[img]) ? 1 : 0;
  }
  float4* p4 = reinterpret_cast<float4*>(p + img * n);
  const float4* r4 = reinterpret_cast<const float4*>(r + img * n);
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 rv = r4[i];
    float4 pv = p4[i];
    pv.x = rv.x + beta * pv.x; pv.y = rv.y + beta * pv.y; pv.z = rv.z + beta * pv.z; pv.w = rv.w + beta * pv.w;
    p4[i] = pv;
  }
}
// norm[b] = sqrt(sum partials)
__global__ void norm_from_partials_kernel(const float* __restrict__ part, float* __restrict__ norm, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) norm[b] = sqrtf(sum_partials(part, b));
}

// ---- workspace arena ---------------------------------------------------------------------------------------------------
struct Arena {
  char* base;
  size_t cur = 0, cap;
  Arena(void* b, size_t c) : base((char*)b), cap(c) {}
  template <typename T>
  T* take(size_t count) {
    size_t o = cur;
    cur += (count * sizeof(T) + 255) & ~(size_t)255;
    return base ? reinterpret_cast<T*>(base + o) : nullptr;
  }
  bool ok() const { return base == nullptr || cur <= cap; }
};

}  // namespace kdip

using namespace kdip;

struct kdip_op {
  int kind, S, sf, s;
  float sigma_s;
  float2* otf = nullptr;     // [S][S/2+1]
  float* invW = nullptr;     // SR: [s][s/2+1]
  float* mask = nullptr;     // inpainting: [3][S][S]
  // Resizer tables (forward gather + CSR of the adjoint), identical for both axes of a square image
  float* rs_w = nullptr; int* rs_idx = nullptr; int rs_taps = 0;
  int* rt_ptr = nullptr; int* rt_o = nullptr; float* rt_w = nullptr;
  // CG convergence poll: pinned host snapshots of the device flags (two slots: the host reads iteration k-1's while iteration k
  // is already queued) + the iteration counts, and the events that guard them.  Allocated by kdip_op_create, never afterwards.
  int* done_host = nullptr;  // [3][kCgMaxBatch]
  cudaEvent_t poll_ev[2] = {nullptr, nullptr};
  std::vector<void*> owned;
};

static constexpr int kCgMaxBatch = 4096;

static bool is_blur(const kdip_op* op) { return op->kind == KDIP_OP_GAUSSIAN_BLUR || op->kind == KDIP_OP_MOTION_BLUR; }

extern "C" void kdip_op_destroy(kdip_op* op) {
  if (!op) return;
  for (void* p : op->owned) cudaFree(p);
  if (op->done_host) cudaFreeHost(op->done_host);
  for (int i = 0; i < 2; ++i)
    if (op->poll_ev[i]) cudaEventDestroy(op->poll_ev[i]);
  delete op;
}

static int op_alloc(kdip_op* op, size_t bytes, void** out) {
  void* p = nullptr;
  KDIP_CUDA(cudaMalloc(&p, bytes));
  op->owned.push_back(p);
  *out = p;
  return KDIP_OK;
}

extern "C" int kdip_op_create(const kdip_op_desc* d, kdip_op** out) {
  KDIP_REQUIRE(d && out, KDIP_EINVAL, "op_create: null argument");
  KDIP_REQUIRE(d->kind >= KDIP_OP_INPAINTING && d->kind <= KDIP_OP_SUPER_RESOLUTION, KDIP_EINVAL, "op_create: unknown kind %d", d->kind);
  KDIP_REQUIRE(d->S >= 16 && d->S <= 256 && (d->S & (d->S - 1)) == 0, KDIP_ESHAPE, "op_create: image side %d must be a power of two in [16,256]", d->S);
  kdip_op* op = new kdip_op();
  op->kind = d->kind; op->S = d->S; op->sigma_s = d->sigma_s;
  op->sf = (d->kind == KDIP_OP_SUPER_RESOLUTION) ? d->sf : 1;
  op->s = op->S / (op->sf > 0 ? op->sf : 1);
  cudaStream_t st = 0;
  auto fail = [&](int code) { kdip_op_destroy(op); return code; };
#define TRY(x) do { int _r = (x); if (_r != KDIP_OK) return fail(_r); } while (0)
#define TRYC(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return fail(::kdip::cuda_fail(_e, #x, __FILE__, __LINE__)); } while (0)
  const int S = op->S, Sh = S / 2 + 1;
  TRYC(cudaMallocHost((void**)&op->done_host, (size_t)3 * kCgMaxBatch * sizeof(int)));
  for (int i = 0; i < 2; ++i) TRYC(cudaEventCreateWithFlags(&op->poll_ev[i], cudaEventDisableTiming));
  if (d->kind == KDIP_OP_INPAINTING) {
    if (!d->mask) { set_error("op_create: inpainting needs a mask"); return fail(KDIP_EINVAL); }
    TRY(op_alloc(op, (size_t)3 * S * S * 4, (void**)&op->mask));
    TRYC(cudaMemcpy(op->mask, d->mask, (size_t)3 * S * S * 4, cudaMemcpyHostToDevice));
  } else {
    if (!d->psf || d->ksize < 1 || d->ksize > S) { set_error("op_create: blur / SR need a PSF with 1 <= ksize <= S (got %d)", d->ksize); return fail(KDIP_EINVAL); }
    if (d->kind == KDIP_OP_SUPER_RESOLUTION) {
      if (op->sf < 2 || S % op->sf != 0 || op->s < 16 || (op->s & (op->s - 1)) != 0) {
        set_error("op_create: scale factor %d unsupported for S=%d", op->sf, S);
        return fail(KDIP_ESHAPE);
      }
      if (!d->rs_w || !d->rs_idx || d->rs_taps < 1) { set_error("op_create: super_resolution needs the Resizer tables"); return fail(KDIP_EINVAL); }
    }
    // OTF = fft2(rolled, zero-padded PSF): our own FFT kernels on one plane
    float *psf_d, *canvas;
    float2* half;
    TRY(op_alloc(op, (size_t)d->ksize * d->ksize * 4, (void**)&psf_d));
    TRY(op_alloc(op, (size_t)S * S * 4, (void**)&canvas));
    TRY(op_alloc(op, (size_t)S * Sh * 8, (void**)&half));
    TRY(op_alloc(op, (size_t)S * Sh * 8, (void**)&op->otf));
    TRYC(cudaMemcpy(psf_d, d->psf, (size_t)d->ksize * d->ksize * 4, cudaMemcpyHostToDevice));
    psf_canvas_kernel<<<grid1d((size_t)S * S), OP_THREADS, 0, st>>>(psf_d, d->ksize, S, canvas);
    psf_place_kernel<<<grid1d((size_t)d->ksize * d->ksize), OP_THREADS, 0, st>>>(psf_d, d->ksize, S, canvas);
    TRY(launch_rows_r2c(canvas, half, 1, S, st));
    SpecOp so;
    memset(&so, 0, sizeof(so));
    so.mode = SPEC_FORWARD_ONLY; so.planes_per_image = 1;
    TRY(launch_cols(half, op->otf, 1, S, so, st));
    if (d->kind == KDIP_OP_SUPER_RESOLUTION) {
      const int s = op->s, sh = s / 2 + 1, taps = d->rs_taps;
      TRY(op_alloc(op, (size_t)s * sh * 4, (void**)&op->invW));
      invw_kernel<<<grid1d((size_t)s * sh), OP_THREADS, 0, st>>>(op->otf, S, op->sf, op->invW);
      TRY(op_alloc(op, (size_t)s * taps * 4, (void**)&op->rs_w));
      TRY(op_alloc(op, (size_t)s * taps * 4, (void**)&op->rs_idx));
      TRYC(cudaMemcpy(op->rs_w, d->rs_w, (size_t)s * taps * 4, cudaMemcpyHostToDevice));
      TRYC(cudaMemcpy(op->rs_idx, d->rs_idx, (size_t)s * taps * 4, cudaMemcpyHostToDevice));
      op->rs_taps = taps;
      // CSR of the adjoint: for every input index, the (output index, weight) pairs that read it, in forward order
      std::vector<int> ptr(S + 1, 0), oi((size_t)s * taps);
      std::vector<float> ww((size_t)s * taps);
      for (int o = 0; o < s; ++o)
        for (int t = 0; t < taps; ++t) {
          const int ix = d->rs_idx[o * taps + t];
          if (ix < 0 || ix >= S) { set_error("op_create: Resizer index %d out of range", ix); return fail(KDIP_EINVAL); }
          ptr[ix + 1]++;
        }
      for (int i = 0; i < S; ++i) ptr[i + 1] += ptr[i];
      std::vector<int> fill(ptr.begin(), ptr.end() - 1);
      for (int o = 0; o < s; ++o)
        for (int t = 0; t < taps; ++t) {
          const int ix = d->rs_idx[o * taps + t];
          oi[fill[ix]] = o;
          ww[fill[ix]] = d->rs_w[o * taps + t];
          fill[ix]++;
        }
      TRY(op_alloc(op, (size_t)(S + 1) * 4, (void**)&op->rt_ptr));
      TRY(op_alloc(op, (size_t)s * taps * 4, (void**)&op->rt_o));
      TRY(op_alloc(op, (size_t)s * taps * 4, (void**)&op->rt_w));
      TRYC(cudaMemcpy(op->rt_ptr, ptr.data(), (size_t)(S + 1) * 4, cudaMemcpyHostToDevice));
      TRYC(cudaMemcpy(op->rt_o, oi.data(), (size_t)s * taps * 4, cudaMemcpyHostToDevice));
      TRYC(cudaMemcpy(op->rt_w, ww.data(), (size_t)s * taps * 4, cudaMemcpyHostToDevice));
    }
    TRYC(cudaGetLastError());
  }
  TRYC(cudaStreamSynchronize(st));
#undef TRY
#undef TRYC
  *out = op;
  return KDIP_OK;
}

// ---- workspace plan (the same walk sizes and assigns) ------------------------------------------------------------------
struct OpWs {
  float2 *specA, *specB;
  float *full[6];        // [B*3][S][S] scratch planes
  float *small_[5];      // [B*3][s][s] (SR)
  float *part, *part2;   // [B][CG_NBLK] partial sums of p.q and of ||r||^2
  CgState st;
  void* dct;             // DCT matrix + temp
};

static size_t plan_ws(const kdip_op* op, int B, void* base, size_t cap, OpWs* w) {
  Arena a(base, cap);
  const size_t planes = (size_t)B * 3, S = op->S, Sh = S / 2 + 1, s = op->s;
  w->specA = a.take<float2>(planes * S * Sh);
  w->specB = a.take<float2>(planes * S * Sh);
  for (int i = 0; i < 6; ++i) w->full[i] = a.take<float>(planes * S * S);
  for (int i = 0; i < 5; ++i) w->small_[i] = a.take<float>(planes * s * s);
  w->part = a.take<float>((size_t)B * CG_NBLK);
  w->part2 = a.take<float>((size_t)B * CG_NBLK);
  for (int i = 0; i < 2; ++i) { w->st.rho[i] = a.take<float>(B); w->st.done[i] = a.take<int>(B); }
  w->st.atol2 = a.take<float>(B);
  w->st.iters = a.take<int>(B);
  w->dct = a.take<char>(dct_workspace_bytes((int)planes, (int)S));
  return a.cur;
}

extern "C" int kdip_op_side(const kdip_op* op) { return op ? op->S : 0; }

extern "C" int kdip_op_workspace_bytes(const kdip_op* op, int B, size_t* bytes) {
  KDIP_REQUIRE(op && bytes && B > 0, KDIP_EINVAL, "op_workspace_bytes: bad argument");
  OpWs w;
  *bytes = plan_ws(op, B, nullptr, 0, &w);
  return KDIP_OK;
}

static int get_ws(const kdip_op* op, int B, void* ws, size_t ws_bytes, OpWs* w) {
  KDIP_REQUIRE(ws != nullptr && ((uintptr_t)ws % 256) == 0, KDIP_EALIGN, "operator workspace must be 256-byte aligned");
  const size_t need = plan_ws(op, B, ws, ws_bytes, w);
  KDIP_REQUIRE(need <= ws_bytes, KDIP_ENOMEM, "operator workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  return KDIP_OK;
}

// out = sign/(S*S) * IFFT( OTF(or conj) * FFT(x) ) * (mul?) + beta*add        (A or A^T of the blur model at S x S)
static int blur_apply(const kdip_op* op, const OpWs& w, const float* x, int conj, float* out, int planes, float sign,
                      const float* mul, float beta, const float* add, cudaStream_t st) {
  const int S = op->S;
  SpecOp so;
  memset(&so, 0, sizeof(so));
  so.mode = SPEC_MULT; so.planes_per_image = 3; so.otf = op->otf; so.conj_otf = conj;
  return launch_spec_filter(x, out, planes, S, so, sign / ((float)S * (float)S), mul, beta, add, w.specA, w.specB, st);
}

static int resizer_forward(const kdip_op* op, const OpWs& w, const float* x, const float* noise, float sigma, float* y, int planes,
                           cudaStream_t st) {
  const int S = op->S, s = op->s;
  float* t = w.full[5];   // [planes][s][S]
  size_t n1 = (size_t)planes * s * S, n2 = (size_t)planes * s * s;
  resize_rows_kernel<<<grid1d(n1), OP_THREADS, 0, st>>>(x, op->rs_w, op->rs_idx, op->rs_taps, S, s, S, t, n1);
  KDIP_LAUNCH_CHECK();
  resize_cols_kernel<<<grid1d(n2), OP_THREADS, 0, st>>>(t, op->rs_w, op->rs_idx, op->rs_taps, S, s, s, noise, sigma, y, n2);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}
static int resizer_adjoint(const kdip_op* op, const OpWs& w, const float* g, float* x, int planes, cudaStream_t st) {
  const int S = op->S, s = op->s;
  float* t = w.full[5];   // [planes][s][S]
  size_t n1 = (size_t)planes * s * S, n2 = (size_t)planes * S * S;
  resize_cols_adj_kernel<<<grid1d(n1), OP_THREADS, 0, st>>>(g, op->rt_ptr, op->rt_o, op->rt_w, S, s, t, n1);
  KDIP_LAUNCH_CHECK();
  resize_rows_adj_kernel<<<grid1d(n2), OP_THREADS, 0, st>>>(t, op->rt_ptr, op->rt_o, op->rt_w, S, s, S, x, n2);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_op_otf(const kdip_op* op, float* fb_full, kdip_stream_t s) {
  KDIP_REQUIRE(op && fb_full && op->otf, KDIP_EINVAL, "op_otf: operator has no OTF");
  otf_full_kernel<<<grid1d((size_t)op->S * op->S), OP_THREADS, 0, (cudaStream_t)s>>>(op->otf, op->S, reinterpret_cast<float2*>(fb_full));
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_op_forward(const kdip_op* op, const float* x, const float* noise, float* y, int B, void* ws, size_t ws_bytes,
                               kdip_stream_t s) {
  KDIP_REQUIRE(op && x && y && B > 0, KDIP_EINVAL, "op_forward: bad argument");
  cudaStream_t st = (cudaStream_t)s;
  if (op->kind == KDIP_OP_INPAINTING) return kdip_inpaint_forward(x, noise, op->mask, op->sigma_s, y, B, 3 * op->S * op->S, s);
  OpWs w;
  int rc = get_ws(op, B, ws, ws_bytes, &w);
  if (rc) return rc;
  if (is_blur(op)) return blur_apply(op, w, x, 0, y, B * 3, 1.f, nullptr, noise ? op->sigma_s : 0.f, noise, st);
  return resizer_forward(op, w, x, noise, op->sigma_s, y, B * 3, st);
}

extern "C" int kdip_op_transpose(const kdip_op* op, const float* y, float* x, int B, void* ws, size_t ws_bytes, kdip_stream_t s) {
  KDIP_REQUIRE(op && x && y && B > 0, KDIP_EINVAL, "op_transpose: bad argument");
  cudaStream_t st = (cudaStream_t)s;
  const size_t n = (size_t)B * 3 * op->S * op->S;
  if (op->kind == KDIP_OP_INPAINTING) {   // measurements.py:228-238: identity on the dense layout
    copy_kernel<<<grid1d(n), OP_THREADS, 0, st>>>(y, x, n);
    KDIP_LAUNCH_CHECK();
    return KDIP_OK;
  }
  OpWs w;
  int rc = get_ws(op, B, ws, ws_bytes, &w);
  if (rc) return rc;
  if (is_blur(op)) return blur_apply(op, w, y, 1, x, B * 3, 1.f, nullptr, 0.f, nullptr, st);
  // SR: ifft2(conj(FB) * fft2(upsample(y)))                                   measurements.py:113-118
  upsample_zero_kernel<<<grid1d(n), OP_THREADS, 0, st>>>(y, w.full[0], op->s, op->sf, n);
  KDIP_LAUNCH_CHECK();
  return blur_apply(op, w, w.full[0], 1, x, B * 3, 1.f, nullptr, 0.f, nullptr, st);
}

// Adjoint of operator.forward(noiseless) - what torch.autograd gives the reference's LinearOperator.auto_transpose
// (measurements.py:48-52) and its DPS branch (condition.py:144-146): conj-OTF blur, the Resizer's adjoint (NOT the FFT-model
// transpose of measurements.py:113-118), or the mask.  g has y's shape; x [B,3,S,S].
extern "C" int kdip_op_forward_adjoint(const kdip_op* op, const float* g, float* x, int B, void* ws, size_t ws_bytes, kdip_stream_t s) {
  KDIP_REQUIRE(op && g && x && B > 0, KDIP_EINVAL, "op_forward_adjoint: bad argument");
  cudaStream_t st = (cudaStream_t)s;
  const size_t n = (size_t)B * 3 * op->S * op->S;
  if (op->kind == KDIP_OP_INPAINTING) {
    mask_mul_kernel<<<grid1d(n), OP_THREADS, 0, st>>>(g, op->mask, x, (size_t)3 * op->S * op->S, n);
    KDIP_LAUNCH_CHECK();
    return KDIP_OK;
  }
  OpWs w;
  int rc = get_ws(op, B, ws, ws_bytes, &w);
  if (rc) return rc;
  if (is_blur(op)) return blur_apply(op, w, g, 1, x, B * 3, 1.f, nullptr, 0.f, nullptr, st);
  return resizer_adjoint(op, w, g, x, B * 3, st);
}

// fft2 over the last two axes of x [B,3,S,S] as interleaved complex64 [B,3,S,S] (torch.fft.fftn(dim=(-2,-1)) of
// utils_sisr.py:91-95, used for the FBFy member of operator.pre_calculated)
extern "C" int kdip_op_fft2(const kdip_op* op, const float* x, float* out_full, int B, void* ws, size_t ws_bytes, kdip_stream_t s) {
  KDIP_REQUIRE(op && x && out_full && B > 0 && op->otf, KDIP_EINVAL, "op_fft2: needs a spectral operator (blur / super_resolution)");
  cudaStream_t st = (cudaStream_t)s;
  OpWs w;
  int rc = get_ws(op, B, ws, ws_bytes, &w);
  if (rc) return rc;
  const int planes = B * 3, S = op->S;
  rc = launch_rows_r2c(x, w.specA, planes, S, st);
  if (rc) return rc;
  SpecOp so;
  memset(&so, 0, sizeof(so));
  so.mode = SPEC_FORWARD_ONLY; so.planes_per_image = 3;
  rc = launch_cols(w.specA, w.specB, planes, S, so, st);
  if (rc) return rc;
  const size_t tot = (size_t)planes * S * S;
  spec_full_kernel<<<grid1d(tot), OP_THREADS, 0, st>>>(w.specB, S, reinterpret_cast<float2*>(out_full), tot);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

// r = y - A x0 in measurement space (blur: S x S, SR: s x s, inpainting: masked residual)
static int residual(const kdip_op* op, const OpWs& w, const float* y, const float* x0, float* r, int B, bool fft_model, cudaStream_t st) {
  const int planes = B * 3;
  const size_t n = (size_t)planes * op->S * op->S, ns = (size_t)planes * op->s * op->s;
  if (op->kind == KDIP_OP_INPAINTING) {
    mask_residual_kernel<<<grid1d(n), OP_THREADS, 0, st>>>(y, x0, op->mask, r, (size_t)3 * op->S * op->S, n);
    KDIP_LAUNCH_CHECK();
    return KDIP_OK;
  }
  if (is_blur(op)) return blur_apply(op, w, x0, 0, r, planes, -1.f, nullptr, 1.f, y, st);
  if (fft_model) {
    int rc = blur_apply(op, w, x0, 0, w.full[4], planes, 1.f, nullptr, 0.f, nullptr, st);
    if (rc) return rc;
    sr_residual_kernel<<<grid1d(ns), OP_THREADS, 0, st>>>(y, w.full[4], r, op->s, op->sf, ns);
    KDIP_LAUNCH_CHECK();
    return KDIP_OK;
  }
  int rc = resizer_forward(op, w, x0, nullptr, 0.f, w.small_[4], planes, st);
  if (rc) return rc;
  sub_kernel<<<grid1d(ns), OP_THREADS, 0, st>>>(y, w.small_[4], r, ns);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_mat_closed(const kdip_op* op, const float* y, const float* x0, const float* theta, float* mat, int B, void* ws,
                               size_t ws_bytes, kdip_stream_t s) {
  KDIP_REQUIRE(op && y && x0 && theta && mat && B > 0, KDIP_EINVAL, "mat_closed: bad argument");
  cudaStream_t st = (cudaStream_t)s;
  float sig = op->sigma_s < 1e-3f ? 1e-3f : op->sigma_s;                       // condition.py:321,353
  if (op->kind == KDIP_OP_INPAINTING) return kdip_inpaint_mat_scalar(y, x0, op->mask, theta, sig, mat, B, 3 * op->S * op->S, s);
  OpWs w;
  int rc = get_ws(op, B, ws, ws_bytes, &w);
  if (rc) return rc;
  const int planes = B * 3, S = op->S;
  SpecOp so;
  memset(&so, 0, sizeof(so));
  so.planes_per_image = 3; so.otf = op->otf; so.theta = theta;
  if (is_blur(op)) {
    // ifft2( fft2(y - A x0) / (sigma_s^2 + theta |FB|^2) * conj(FB) )          condition.py:357
    rc = residual(op, w, y, x0, w.full[0], B, true, st);
    if (rc) return rc;
    so.mode = SPEC_DIV_CONJ; so.sigma_s2 = sig * sig;
    return launch_spec_filter(w.full[0], mat, planes, S, so, 1.f / ((float)S * (float)S), nullptr, 0.f, nullptr, w.specA, w.specB, st);
  }
  // SR: ifft2( conj(FB) * tile( fft2(y - down(A x0)) / (sigma_s^2 + theta invW) ) )       condition.py:404-410
  if (sig < 1e-2f) sig = 1e-2f;
  const int s_ = op->s;
  rc = residual(op, w, y, x0, w.small_[0], B, true, st);
  if (rc) return rc;
  rc = launch_rows_r2c(w.small_[0], w.specA, planes, s_, st);
  if (rc) return rc;
  so.mode = SPEC_DIV_TABLE; so.table = op->invW; so.sigma_s2 = sig * sig;
  rc = launch_cols(w.specA, w.specB, planes, s_, so, st);
  if (rc) return rc;
  rc = launch_rows_c2r(w.specB, w.small_[1], planes, s_, 1.f / ((float)s_ * (float)s_), nullptr, 0.f, nullptr, st);
  if (rc) return rc;
  // the periodic tiling of a spectrum is the spectrum of the zero-filled upsample: go back through real space
  const size_t n = (size_t)planes * S * S;
  upsample_zero_kernel<<<grid1d(n), OP_THREADS, 0, st>>>(w.small_[1], w.full[0], s_, op->sf, n);
  KDIP_LAUNCH_CHECK();
  return blur_apply(op, w, w.full[0], 1, mat, planes, 1.f, nullptr, 0.f, nullptr, st);
}

// t_out = W( theta .* W^T t_in )   (Sigma applied in image space; condition/utils.py:146-163 semantics)
static int apply_cov(const OpWs& w, int ot, const float* theta, const float* t_in, float* tmp, float* t_out, int B, int S, cudaStream_t st) {
  const int planes = B * 3;
  const size_t n = (size_t)planes * S * S;
  if (ot == KDIP_OT_NONE) {
    mul_kernel<<<grid1d(n), OP_THREADS, 0, st>>>(theta, t_in, t_out, n);
    KDIP_LAUNCH_CHECK();
    return KDIP_OK;
  }
  if (ot == KDIP_OT_DWT) {
    return launch_dwt_cov(t_in, theta, planes, tmp, t_out, planes, S, st);
  }
  int rc = launch_dct(t_in, theta, B, tmp, B, S, 0, w.dct, st);
  if (rc) return rc;
  return launch_dct(tmp, nullptr, 0, t_out, B, S, 1, w.dct, st);
}

extern "C" int kdip_ortho(int ot, int inverse, const float* x, const float* mul, float* out, int B, int S, void* ws, size_t ws_bytes,
                          kdip_stream_t s) {
  KDIP_REQUIRE(x && out && B > 0, KDIP_EINVAL, "ortho: bad argument");
  KDIP_REQUIRE(!(inverse && mul), KDIP_EINVAL, "ortho: the multiplier applies to the forward transform only");
  cudaStream_t st = (cudaStream_t)s;
  if (ot == KDIP_OT_NONE) {
    const size_t n = (size_t)B * 3 * S * S;
    if (mul) mul_kernel<<<grid1d(n), OP_THREADS, 0, st>>>(x, mul, out, n);
    else copy_kernel<<<grid1d(n), OP_THREADS, 0, st>>>(x, out, n);
    KDIP_LAUNCH_CHECK();
    return KDIP_OK;
  }
  if (ot == KDIP_OT_DWT) return launch_dwt(x, mul, B * 3, out, B * 3, S, inverse, st);
  KDIP_REQUIRE(ot == KDIP_OT_DCT, KDIP_EINVAL, "ortho: unknown transform %d", ot);
  KDIP_REQUIRE(ws != nullptr && ((uintptr_t)ws % 256) == 0 && ws_bytes >= dct_workspace_bytes(B * 3, S), KDIP_ENOMEM,
               "ortho(dct): workspace of %zu bytes (256-byte aligned) required", dct_workspace_bytes(B * 3, S));
  KDIP_REQUIRE(x != out, KDIP_EINVAL, "ortho(dct): in-place transform unsupported");
  return launch_dct(x, mul, B, out, B, S, inverse, ws, st);
}

extern "C" int kdip_ortho_workspace_bytes(int ot, int B, int S, size_t* bytes) {
  KDIP_REQUIRE(bytes && B > 0, KDIP_EINVAL, "ortho_workspace_bytes: bad argument");
  *bytes = (ot == KDIP_OT_DCT) ? dct_workspace_bytes(B * 3, S) : 0;
  return KDIP_OK;
}

extern "C" int kdip_mat_cg(kdip_op* op, const float* y, const float* x0, const float* theta_map, int ot, float* mat, int B, float tol,
                           int maxiter, int* iters_out, void* ws, size_t ws_bytes, kdip_stream_t s) {
  KDIP_REQUIRE(op && y && x0 && theta_map && mat && B > 0 && maxiter > 0, KDIP_EINVAL, "mat_cg: bad argument");
  KDIP_REQUIRE(ot == KDIP_OT_NONE || ot == KDIP_OT_DCT || ot == KDIP_OT_DWT, KDIP_EINVAL, "mat_cg: unknown transform %d", ot);
  cudaStream_t st = (cudaStream_t)s;
  OpWs w;
  int rc = get_ws(op, B, ws, ws_bytes, &w);
  if (rc) return rc;
  KDIP_REQUIRE(B <= kCgMaxBatch, KDIP_ESHAPE, "mat_cg: batch %d exceeds the poll buffer (%d images)", B, kCgMaxBatch);
  const int planes = B * 3, S = op->S, s_ = op->s;
  const bool sr = op->kind == KDIP_OP_SUPER_RESOLUTION;
  float sig = op->sigma_s < 1e-3f ? 1e-3f : op->sigma_s;
  if (sr && sig < 1e-2f) sig = 1e-2f;
  const float sigma2 = sig * sig;
  const size_t nfull = (size_t)3 * S * S;                      // per image, image space
  const size_t n = sr ? (size_t)3 * s_ * s_ : nfull;           // per image, unknowns (measurement space)
  // CG vectors live in measurement space
  float *xs, *r, *p, *q;
  if (sr) { xs = w.small_[0]; r = w.small_[1]; p = w.small_[2]; q = w.small_[3]; }
  else { xs = w.full[0]; r = w.full[1]; p = w.full[2]; q = w.full[3]; }
  float *t1 = w.full[4], *t2 = w.full[5];
  // SR needs two more full-resolution temporaries than the four CG vectors leave free
  float *t3 = sr ? w.full[0] : nullptr, *t4 = sr ? w.full[1] : nullptr;

  // q = sigma_s^2 u + A Sigma A^T u                                   condition.py:338,366-377,419-430
  std::function<int(const float*, float*)> matvec;
  if (op->kind == KDIP_OP_INPAINTING) {
    matvec = [&](const float* u, float* out) -> int {
      const size_t tot = (size_t)B * nfull;
      if (ot == KDIP_OT_NONE) {
        mask_matvec_kernel<<<grid1d(tot), OP_THREADS, 0, st>>>(u, nullptr, theta_map, op->mask, sigma2, out, nfull, tot);
      } else {
        int r_ = apply_cov(w, ot, theta_map, u, t1, t2, B, S, st);
        if (r_) return r_;
        mask_matvec_kernel<<<grid1d(tot), OP_THREADS, 0, st>>>(u, t2, nullptr, op->mask, sigma2, out, nfull, tot);
      }
      KDIP_LAUNCH_CHECK();
      return KDIP_OK;
    };
  } else if (is_blur(op)) {
    matvec = [&](const float* u, float* out) -> int {
      int r_;
      if (ot == KDIP_OT_NONE) {
        r_ = blur_apply(op, w, u, 1, t1, planes, 1.f, theta_map, 0.f, nullptr, st);            // theta .* A^T u
        if (r_) return r_;
        return blur_apply(op, w, t1, 0, out, planes, 1.f, nullptr, sigma2, u, st);             // A(.) + sigma^2 u
      }
      r_ = blur_apply(op, w, u, 1, t1, planes, 1.f, nullptr, 0.f, nullptr, st);
      if (r_) return r_;
      r_ = apply_cov(w, ot, theta_map, t1, t2, t1, B, S, st);
      if (r_) return r_;
      return blur_apply(op, w, t1, 0, out, planes, 1.f, nullptr, sigma2, u, st);
    };
  } else {
    matvec = [&](const float* u, float* out) -> int {
      const size_t tot_full = (size_t)B * nfull, tot = (size_t)B * n;
      upsample_zero_kernel<<<grid1d(tot_full), OP_THREADS, 0, st>>>(u, t3, s_, op->sf, tot_full);
      KDIP_LAUNCH_CHECK();
      int r_ = blur_apply(op, w, t3, 1, t1, planes, 1.f, ot == KDIP_OT_NONE ? theta_map : nullptr, 0.f, nullptr, st);
      if (r_) return r_;
      if (ot != KDIP_OT_NONE) {
        r_ = apply_cov(w, ot, theta_map, t1, t2, t1, B, S, st);
        if (r_) return r_;
      }
      r_ = blur_apply(op, w, t1, 0, t4, planes, 1.f, nullptr, 0.f, nullptr, st);
      if (r_) return r_;
      sr_matvec_tail_kernel<<<grid1d(tot), OP_THREADS, 0, st>>>(u, t4, sigma2, out, s_, op->sf, tot);
      KDIP_LAUNCH_CHECK();
      return KDIP_OK;
    };
  }

  // b = y - A x0 (FFT model for SR, as the reference's solver uses; condition.py:431)
  if (sr) {
    // residual() uses full[4] as scratch and writes r (small_)
    rc = residual(op, w, y, x0, r, B, true, st);
  } else {
    rc = residual(op, w, y, x0, r, B, true, st);
  }
  if (rc) return rc;
  dim3 g(CG_NBLK, B);
  cg_init_kernel<<<g, OP_THREADS, 0, st>>>(r, xs, p, w.part2, n);
  KDIP_LAUNCH_CHECK();
  cg_begin_kernel<<<(B + 127) / 128, 128, 0, st>>>(w.part2, w.st, tol, B);
  KDIP_LAUNCH_CHECK();
  // Iteration k reads rho / done slot k & 1 and publishes slot (k + 1) & 1.  The host never waits for the iteration it has just
  // queued: after queueing iteration k it snapshots that iteration's flags (async copy + event) and inspects the snapshot of
  // iteration k - 1, so the GPU always has one iteration of work queued behind the one being polled and at most one surplus
  // iteration (frozen images: its vector kernels return immediately) runs after the last image converged.
  bool all_done = false;
  int* flags[2] = {op->done_host, op->done_host + kCgMaxBatch};
  auto inspect = [&](int slot) {
    bool d = true;
    for (int b = 0; b < B; ++b) d = d && (flags[slot][b] != 0);
    return d;
  };
  int it = 0;
  for (; it < maxiter && !all_done; ++it) {
    const int cur = it & 1;
    rc = matvec(p, q);
    if (rc) return rc;
    dot_partial_kernel<<<g, OP_THREADS, 0, st>>>(p, q, w.part, n);
    KDIP_LAUNCH_CHECK();
    cg_update_kernel<<<g, OP_THREADS, 0, st>>>(xs, r, p, q, w.st, cur, w.part, w.part2, n);
    KDIP_LAUNCH_CHECK();
    cg_p_kernel<<<g, OP_THREADS, 0, st>>>(p, r, w.st, cur, w.part2, n);
    KDIP_LAUNCH_CHECK();
    KDIP_CUDA(cudaMemcpyAsync(flags[cur], w.st.done[cur ^ 1], (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    KDIP_CUDA(cudaEventRecord(op->poll_ev[cur], st));
    if (it > 0) {
      KDIP_CUDA(cudaEventSynchronize(op->poll_ev[cur ^ 1]));
      all_done = inspect(cur ^ 1);
    }
  }
  if (!all_done && it > 0) {      // the last queued iteration has not been inspected yet
    const int last = (it - 1) & 1;
    KDIP_CUDA(cudaEventSynchronize(op->poll_ev[last]));
    all_done = inspect(last);
  }
  if (iters_out) {
    int* ih = op->done_host + 2 * kCgMaxBatch;
    KDIP_CUDA(cudaMemcpyAsync(ih, w.st.iters, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    KDIP_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b) iters_out[b] = ih[b];
  }
  // mat = A^T u (inpainting: the solution itself, condition.py:346)
  const size_t tot = (size_t)B * nfull;
  if (op->kind == KDIP_OP_INPAINTING) {
    copy_kernel<<<grid1d(tot), OP_THREADS, 0, st>>>(xs, mat, tot);
    KDIP_LAUNCH_CHECK();
  } else if (is_blur(op)) {
    rc = blur_apply(op, w, xs, 1, mat, planes, 1.f, nullptr, 0.f, nullptr, st);
    if (rc) return rc;
  } else {
    upsample_zero_kernel<<<grid1d(tot), OP_THREADS, 0, st>>>(xs, t1, s_, op->sf, tot);
    KDIP_LAUNCH_CHECK();
    rc = blur_apply(op, w, t1, 1, mat, planes, 1.f, nullptr, 0.f, nullptr, st);
    if (rc) return rc;
  }
  if (!all_done) {
    set_error("CG not converge.");   // the reference's warning text (condition.py:345)
    return KDIP_ENOTCONV;
  }
  return KDIP_OK;
}

// DPS (condition.py:140-148): r = y - A(x0) with A = operator.forward(noiseless); norm[b] = ||r_b||_2 ; v = A^T r
// (the adjoint of that same forward: conj-OTF blur, Resizer adjoint, or the mask).
extern "C" int kdip_dps_grad(const kdip_op* op, const float* y, const float* x0, float* v, float* norm, int B, void* ws,
                             size_t ws_bytes, kdip_stream_t s) {
  KDIP_REQUIRE(op && y && x0 && v && norm && B > 0, KDIP_EINVAL, "dps_grad: bad argument");
  cudaStream_t st = (cudaStream_t)s;
  OpWs w;
  int rc = get_ws(op, B, ws, ws_bytes, &w);
  if (rc) return rc;
  const int planes = B * 3, S = op->S;
  const bool sr = op->kind == KDIP_OP_SUPER_RESOLUTION;
  const size_t n = sr ? (size_t)3 * op->s * op->s : (size_t)3 * S * S;
  float* r = sr ? w.small_[0] : w.full[0];
  if (op->kind == KDIP_OP_INPAINTING) {
    // forward(noiseless) = x*mask; y is already masked: r = y - mask*x0 = mask*(y - x0) up to rounding
    const size_t tot = (size_t)B * n;
    mask_mul_kernel<<<grid1d(tot), OP_THREADS, 0, st>>>(x0, op->mask, w.full[1], n, tot);
    KDIP_LAUNCH_CHECK();
    sub_kernel<<<grid1d(tot), OP_THREADS, 0, st>>>(y, w.full[1], r, tot);
    KDIP_LAUNCH_CHECK();
  } else {
    rc = residual(op, w, y, x0, r, B, false, st);
    if (rc) return rc;
  }
  dim3 g(CG_NBLK, B);
  dot_partial_kernel<<<g, OP_THREADS, 0, st>>>(r, nullptr, w.part, n);
  KDIP_LAUNCH_CHECK();
  norm_from_partials_kernel<<<(B + 127) / 128, 128, 0, st>>>(w.part, norm, B);
  KDIP_LAUNCH_CHECK();
  if (op->kind == KDIP_OP_INPAINTING) {
    const size_t tot = (size_t)B * n;
    mask_mul_kernel<<<grid1d(tot), OP_THREADS, 0, st>>>(r, op->mask, v, n, tot);
    KDIP_LAUNCH_CHECK();
    return KDIP_OK;
  }
  if (is_blur(op)) return blur_apply(op, w, r, 1, v, planes, 1.f, nullptr, 0.f, nullptr, st);
  return resizer_adjoint(op, w, r, v, planes, st);
}
