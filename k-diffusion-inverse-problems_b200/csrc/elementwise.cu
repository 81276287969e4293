// HBM-bound elementwise kernels of the guided-sampling loop: sampler updates (k_diffusion/sampling.py:118-184),
// p_mean_variance epilogue + Convert variance (gaussian_diffusion.py:262-311, condition/condition.py:241-248),
// its VJP seed, guidance combine (condition.py:131-173) and the inpainting operator / closed-form solve
// (measurements.py:211-238, condition.py:317-323).  All fp32, float4-vectorised, grid = k * #SMs, no smem needed
// (every element is touched once — see DESIGN.md for bytes per element of each kernel).
// Compiled without FMA contraction: the reference evaluates these expressions as separately rounded ATen mul / add
// kernels, and the mask / inpainting ops are held to bit-exactness against it.
// NVCC_FLAGS: -fmad=false
#include "kdip_common.cuh"

namespace kdip {

static inline int ew_grid(size_t nvec, int threads) {
  size_t blocks = (nvec + threads - 1) / threads;
  size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
#define EW_THREADS 256
#define EW_LOOP(i, n) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

__device__ __forceinline__ float4 ld4(const float* p, size_t i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }
__device__ __forceinline__ void st4(float* p, size_t i, float4 v) { reinterpret_cast<float4*>(p)[i] = v; }

// ---- sampler -----------------------------------------------------------------------------------------------------
__global__ void churn_kernel(float* __restrict__ x, const float* __restrict__ noise, float a, size_t n4) {
  EW_LOOP(i, n4) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    float4 e = ld4(noise, i);
    v.x += e.x * a; v.y += e.y * a; v.z += e.z * a; v.w += e.w * a;
    st4(x, i, v);
  }
}

__global__ void euler_kernel(const float* __restrict__ x, const float* __restrict__ den, float sigma, float dt,
                             float* __restrict__ x_out, float* __restrict__ d_out, size_t n4) {
  EW_LOOP(i, n4) {
    float4 a = ld4(x, i), b = ld4(den, i), d, o;
    // same operation order as the reference: d = (x - denoised) / sigma ; x + d * dt
    d.x = (a.x - b.x) / sigma; d.y = (a.y - b.y) / sigma; d.z = (a.z - b.z) / sigma; d.w = (a.w - b.w) / sigma;
    o.x = a.x + d.x * dt; o.y = a.y + d.y * dt; o.z = a.z + d.z * dt; o.w = a.w + d.w * dt;
    st4(x_out, i, o);
    if (d_out) st4(d_out, i, d);
  }
}

__global__ void heun_kernel(const float* __restrict__ x, const float* __restrict__ d, const float* __restrict__ x2,
                            const float* __restrict__ den2, float sigma, float dt, float* __restrict__ x_out, size_t n4) {
  EW_LOOP(i, n4) {
    float4 a = ld4(x, i), d1 = ld4(d, i), b = ld4(x2, i), c = ld4(den2, i), o;
    o.x = a.x + (d1.x + (b.x - c.x) / sigma) / 2.f * dt;
    o.y = a.y + (d1.y + (b.y - c.y) / sigma) / 2.f * dt;
    o.z = a.z + (d1.z + (b.z - c.z) / sigma) / 2.f * dt;
    o.w = a.w + (d1.w + (b.w - c.w) / sigma) / 2.f * dt;
    st4(x_out, i, o);
  }
}

// out = a*x + b*y + c*z with host scalars, evaluated as ((a*x) + (b*y)) + (c*z), each product and sum rounded separately
// (this file is built with -fmad=false); y / z may be NULL.  The update of every other k-diffusion sampler is one or two of these.
template <int TERMS>
__global__ void lincomb3_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, float a,
                                float b, float c, float* __restrict__ out, size_t n4) {
  EW_LOOP(i, n4) {
    float4 o = ld4(x, i);
    o.x = a * o.x; o.y = a * o.y; o.z = a * o.z; o.w = a * o.w;
    if (TERMS >= 2) {
      const float4 v = ld4(y, i);
      o.x = o.x + b * v.x; o.y = o.y + b * v.y; o.z = o.z + b * v.z; o.w = o.w + b * v.w;
    }
    if (TERMS >= 3) {
      const float4 v = ld4(z, i);
      o.x = o.x + c * v.x; o.y = o.y + c * v.y; o.z = o.z + c * v.z; o.w = o.w + c * v.w;
    }
    st4(out, i, o);
  }
}

// ---- p_mean_variance epilogue ----------------------------------------------------------------------------------------
// grid.y = image; each thread handles 4 pixels of all 3 channels.
__global__ void pmv_kernel(const float* __restrict__ out, const float* __restrict__ x, const kdip_pmv_scalars* __restrict__ sc,
                           float* __restrict__ x0, float* __restrict__ var, int convert, int HW4) {
  const int b = blockIdx.y;
  const kdip_pmv_scalars s = sc[b];
  const float a = s.recip * s.c_in;
  const size_t HW = (size_t)HW4;  // in float4 units
  const float* ob = out + (size_t)b * 6 * HW * 4;
  const float* xb = x + (size_t)b * 3 * HW * 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * HW; i += (size_t)gridDim.x * blockDim.x) {
    float4 e = ld4(ob, i), xv = ld4(xb, i), r;
    // x0 = clamp(sqrt(1/abar)*(x*c_in) - sqrt(1/abar-1)*eps, -1, 1)
    r.x = fminf(fmaxf(s.recip * (xv.x * s.c_in) - s.recipm1 * e.x, -1.f), 1.f);
    r.y = fminf(fmaxf(s.recip * (xv.y * s.c_in) - s.recipm1 * e.y, -1.f), 1.f);
    r.z = fminf(fmaxf(s.recip * (xv.z * s.c_in) - s.recipm1 * e.z, -1.f), 1.f);
    r.w = fminf(fmaxf(s.recip * (xv.w * s.c_in) - s.recipm1 * e.w, -1.f), 1.f);
    (void)a;
    st4(x0 + (size_t)b * 3 * HW * 4, i, r);
    if (var) {
      float4 v = ld4(ob + 3 * HW * 4, i), o;
      // variance = exp(frac*max_log + (1-frac)*min_log), frac = (v+1)/2 ; Convert: clip((variance - beta~)/coef1^2, 1e-6)
      const float vv[4] = {v.x, v.y, v.z, v.w};
      float r4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float f = (vv[j] + 1.f) * 0.5f;
        const float mv = expf(f * s.max_log + (1.f - f) * s.min_log);          // model variance, gaussian_diffusion.py:271-276
        r4[j] = convert ? fmaxf((mv - s.post_var) / s.coef1_sq, 1e-6f) : mv;   // Eq. (22), condition.py:243-246
      }
      o = make_float4(r4[0], r4[1], r4[2], r4[3]);
      st4(var + (size_t)b * 3 * HW * 4, i, o);
    }
  }
}

__global__ void pmv_vjp_seed_kernel(const float* __restrict__ x0, const float* __restrict__ v,
                                    const kdip_pmv_scalars* __restrict__ sc, float* __restrict__ seed,
                                    float* __restrict__ direct, int HW4) {
  const int b = blockIdx.y;
  const kdip_pmv_scalars s = sc[b];
  const size_t HW = (size_t)HW4;
  const float ce = -s.recipm1, cd = s.recip * s.c_in;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * HW; i += (size_t)gridDim.x * blockDim.x) {
    float4 g = ld4(v + (size_t)b * 3 * HW * 4, i), se, di;
    if (x0 != nullptr) {
      // clamp backward: gradient passes where the clamp is inactive (x0 == NULL: the unclamped v2 denoiser, condition.py:291)
      const float4 m = ld4(x0 + (size_t)b * 3 * HW * 4, i);
      g.x = (m.x > -1.f && m.x < 1.f) ? g.x : 0.f;
      g.y = (m.y > -1.f && m.y < 1.f) ? g.y : 0.f;
      g.z = (m.z > -1.f && m.z < 1.f) ? g.z : 0.f;
      g.w = (m.w > -1.f && m.w < 1.f) ? g.w : 0.f;
    }
    se.x = ce * g.x; se.y = ce * g.y; se.z = ce * g.z; se.w = ce * g.w;
    di.x = cd * g.x; di.y = cd * g.y; di.z = cd * g.z; di.w = cd * g.w;
    st4(seed + (size_t)b * 6 * HW * 4, i, se);
    st4(seed + (size_t)b * 6 * HW * 4 + 3 * HW * 4, i, make_float4(0.f, 0.f, 0.f, 0.f));
    if (direct) st4(direct + (size_t)b * 3 * HW * 4, i, di);
  }
}

__global__ void combine_kernel(const float* __restrict__ x0, const float* __restrict__ g, const float* __restrict__ direct,
                               const float* __restrict__ coef, const float* __restrict__ c_in, float* __restrict__ hat, int CHW4) {
  const int b = blockIdx.y;
  const float k = coef[b];
  const float ci = c_in ? c_in[b] : 1.f;
  const size_t off = (size_t)b * CHW4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)CHW4; i += (size_t)gridDim.x * blockDim.x) {
    float4 m = ld4(x0, off + i), u = ld4(g, off + i), o;
    float4 d = direct ? ld4(direct, off + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x = fminf(fmaxf(m.x + k * (ci * u.x + d.x), -1.f), 1.f);
    o.y = fminf(fmaxf(m.y + k * (ci * u.y + d.y), -1.f), 1.f);
    o.z = fminf(fmaxf(m.z + k * (ci * u.z + d.z), -1.f), 1.f);
    o.w = fminf(fmaxf(m.w + k * (ci * u.w + d.w), -1.f), 1.f);
    st4(hat, off + i, o);
  }
}

// out = a[b]*x + c[b]*y (y may be NULL); no clipping (TMPD's Jacobian-diagonal proxy, condition.py:268-269)
__global__ void lincomb_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ a,
                               const float* __restrict__ c, float* __restrict__ out, int CHW4) {
  const int b = blockIdx.y;
  const float ka = a[b], kc = c ? c[b] : 0.f;
  const size_t off = (size_t)b * CHW4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)CHW4; i += (size_t)gridDim.x * blockDim.x) {
    float4 u = ld4(x, off + i), o;
    float4 v = y ? ld4(y, off + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x = ka * u.x + kc * v.x; o.y = ka * u.y + kc * v.y; o.z = ka * u.z + kc * v.z; o.w = ka * u.w + kc * v.w;
    st4(out, off + i, o);
  }
}

// v2 (DWT-Var) epilogue: x0 = eps*c_out + x, c_out = -sigma; variances exp(logvar)*c_out^2   (condition.py:287-300)
__global__ void v2_epilogue_kernel(const float* __restrict__ out6, const float* __restrict__ cov6, const float* __restrict__ x,
                                   const float* __restrict__ sigma, float* __restrict__ x0, float* __restrict__ var,
                                   float* __restrict__ var_ot, int HW4) {
  const int b = blockIdx.y;
  const float c_out = -sigma[b];
  const float c2 = c_out * c_out;
  const size_t HW = (size_t)HW4;
  const float* ob = out6 + (size_t)b * 6 * HW * 4;
  const float* cb = cov6 + (size_t)b * 6 * HW * 4;
  const size_t o3 = (size_t)b * 3 * HW * 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * HW; i += (size_t)gridDim.x * blockDim.x) {
    float4 e = ld4(ob, i), xv = ld4(x + o3, i), r;
    r.x = e.x * c_out + xv.x; r.y = e.y * c_out + xv.y; r.z = e.z * c_out + xv.z; r.w = e.w * c_out + xv.w;
    st4(x0 + o3, i, r);
    if (var) {
      float4 l = ld4(cb, i), lo = ld4(cb + 3 * HW * 4, i), v, vo;
      v.x = expf(l.x) * c2; v.y = expf(l.y) * c2; v.z = expf(l.z) * c2; v.w = expf(l.w) * c2;
      vo.x = expf(lo.x) * c2; vo.y = expf(lo.y) * c2; vo.z = expf(lo.z) * c2; vo.w = expf(lo.w) * c2;
      st4(var + o3, i, v);
      st4(var_ot + o3, i, vo);
    }
  }
}

// ---- inpainting ------------------------------------------------------------------------------------------------------
__global__ void inpaint_fwd_kernel(const float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ mask,
                                   float sigma_s, float* __restrict__ y, int CHW4) {
  const size_t off = (size_t)blockIdx.y * CHW4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)CHW4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = ld4(x, off + i), m = ld4(mask, i), o;
    if (noise) {
      float4 e = ld4(noise, off + i);
      a.x += sigma_s * e.x; a.y += sigma_s * e.y; a.z += sigma_s * e.z; a.w += sigma_s * e.w;
    }
    o.x = a.x * m.x; o.y = a.y * m.y; o.z = a.z * m.z; o.w = a.w * m.w;
    st4(y, off + i, o);
  }
}

__global__ void inpaint_mat_kernel(const float* __restrict__ y, const float* __restrict__ x0, const float* __restrict__ mask,
                                   const float* __restrict__ theta, float sigma_s2, float* __restrict__ mat, int CHW4) {
  const int b = blockIdx.y;
  const float den = sigma_s2 + theta[b];
  const size_t off = (size_t)b * CHW4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)CHW4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = ld4(y, off + i), c = ld4(x0, off + i), m = ld4(mask, i), o;
    // (mask*y - mask*x0) / (sigma_s^2 + theta): true division, same rounding as the reference
    o.x = (m.x * a.x - m.x * c.x) / den; o.y = (m.y * a.y - m.y * c.y) / den;
    o.z = (m.z * a.z - m.z * c.z) / den; o.w = (m.w * a.w - m.w * c.w) / den;
    st4(mat, off + i, o);
  }
}

__global__ void gather_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, float* __restrict__ dst, int CHW, int M) {
  const size_t b = blockIdx.y;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)M; i += (size_t)gridDim.x * blockDim.x)
    dst[b * M + i] = __ldg(src + b * CHW + __ldg(idx + i));
}
__global__ void scatter_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, float* __restrict__ dst, int CHW, int M) {
  const size_t b = blockIdx.y;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)M; i += (size_t)gridDim.x * blockDim.x)
    dst[b * CHW + __ldg(idx + i)] = __ldg(src + b * M + i);
}

}  // namespace kdip

using namespace kdip;

#define REQ_ALIGN16(p) KDIP_REQUIRE(((uintptr_t)(p) % 16) == 0, KDIP_EALIGN, #p " must be 16-byte aligned")
#define REQ_MULT4(n) KDIP_REQUIRE(((n) % 4) == 0, KDIP_ESHAPE, #n " must be a multiple of 4 (got %lld)", (long long)(n))

extern "C" int kdip_churn(float* x, const float* noise, float s_noise, float sigma, float sigma_hat, size_t n, kdip_stream_t s) {
  REQ_ALIGN16(x); REQ_ALIGN16(noise); REQ_MULT4(n);
  // eps = noise*s_noise ; x + eps*sqrt(sigma_hat^2 - sigma^2)   (host scalar math in fp32 like the reference's 0-dim tensors)
  float a = s_noise * sqrtf(sigma_hat * sigma_hat - sigma * sigma);
  churn_kernel<<<ew_grid(n / 4, EW_THREADS), EW_THREADS, 0, (cudaStream_t)s>>>(x, noise, a, n / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_euler_step(const float* x, const float* denoised, float sigma_hat, float dt, float* x_out, float* d_out,
                               size_t n, kdip_stream_t s) {
  REQ_ALIGN16(x); REQ_ALIGN16(denoised); REQ_ALIGN16(x_out); REQ_ALIGN16(d_out); REQ_MULT4(n);
  KDIP_REQUIRE(sigma_hat > 0.f, KDIP_EINVAL, "euler_step: sigma_hat must be > 0");
  euler_kernel<<<ew_grid(n / 4, EW_THREADS), EW_THREADS, 0, (cudaStream_t)s>>>(x, denoised, sigma_hat, dt, x_out, d_out, n / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_heun_step(const float* x, const float* d, const float* x2, const float* denoised2, float sigma_next,
                              float dt, float* x_out, size_t n, kdip_stream_t s) {
  REQ_ALIGN16(x); REQ_ALIGN16(d); REQ_ALIGN16(x2); REQ_ALIGN16(denoised2); REQ_ALIGN16(x_out); REQ_MULT4(n);
  KDIP_REQUIRE(sigma_next > 0.f, KDIP_EINVAL, "heun_step: sigma_next must be > 0");
  heun_kernel<<<ew_grid(n / 4, EW_THREADS), EW_THREADS, 0, (cudaStream_t)s>>>(x, d, x2, denoised2, sigma_next, dt, x_out, n / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_lincomb3(const float* x, const float* y, const float* z, float a, float b, float c, float* out, size_t n,
                             kdip_stream_t s) {
  REQ_ALIGN16(x); REQ_ALIGN16(y); REQ_ALIGN16(z); REQ_ALIGN16(out); REQ_MULT4(n);
  KDIP_REQUIRE(x != nullptr && out != nullptr, KDIP_EINVAL, "lincomb3: x and out are required");
  KDIP_REQUIRE(y != nullptr || z == nullptr, KDIP_EINVAL, "lincomb3: z without y");
  const dim3 g = ew_grid(n / 4, EW_THREADS);
  if (z) lincomb3_kernel<3><<<g, EW_THREADS, 0, (cudaStream_t)s>>>(x, y, z, a, b, c, out, n / 4);
  else if (y) lincomb3_kernel<2><<<g, EW_THREADS, 0, (cudaStream_t)s>>>(x, y, z, a, b, c, out, n / 4);
  else lincomb3_kernel<1><<<g, EW_THREADS, 0, (cudaStream_t)s>>>(x, y, z, a, b, c, out, n / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

static inline dim3 grid_by(int per_image_vec, int B) {
  int gx = (per_image_vec + EW_THREADS - 1) / EW_THREADS;
  int cap = (num_sms() * 16 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3(gx, B);
}

extern "C" int kdip_pmv_epilogue(const float* unet_out, const float* x, const kdip_pmv_scalars* sc, float* x0_mean,
                                 float* x0_var, int var_mode, int B, int HW, kdip_stream_t s) {
  REQ_ALIGN16(unet_out); REQ_ALIGN16(x); REQ_ALIGN16(x0_mean); REQ_ALIGN16(x0_var); REQ_MULT4(HW);
  KDIP_REQUIRE(B > 0, KDIP_ESHAPE, "pmv_epilogue: B must be > 0");
  KDIP_REQUIRE(x0_var == nullptr || var_mode == KDIP_VAR_MODEL || var_mode == KDIP_VAR_CONVERT, KDIP_EINVAL, "pmv_epilogue: bad var_mode %d", var_mode);
  pmv_kernel<<<grid_by(3 * HW / 4, B), EW_THREADS, 0, (cudaStream_t)s>>>(unet_out, x, sc, x0_mean, x0_var,
                                                                         var_mode == KDIP_VAR_CONVERT ? 1 : 0, HW / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_pmv_vjp_seed(const float* x0_mean, const float* v, const kdip_pmv_scalars* sc, float* seed, float* direct,
                                 int B, int HW, kdip_stream_t s) {
  REQ_ALIGN16(x0_mean); REQ_ALIGN16(v); REQ_ALIGN16(seed); REQ_ALIGN16(direct); REQ_MULT4(HW);
  KDIP_REQUIRE(B > 0, KDIP_ESHAPE, "pmv_vjp_seed: B must be > 0");
  pmv_vjp_seed_kernel<<<grid_by(3 * HW / 4, B), EW_THREADS, 0, (cudaStream_t)s>>>(x0_mean, v, sc, seed, direct, HW / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_guidance_combine(const float* x0_mean, const float* unet_grad, const float* direct, const float* coef,
                                     const float* c_in, float* hat_x0, int B, int CHW, kdip_stream_t s) {
  REQ_ALIGN16(x0_mean); REQ_ALIGN16(unet_grad); REQ_ALIGN16(direct); REQ_ALIGN16(hat_x0); REQ_MULT4(CHW);
  KDIP_REQUIRE(B > 0, KDIP_ESHAPE, "guidance_combine: B must be > 0");
  combine_kernel<<<grid_by(CHW / 4, B), EW_THREADS, 0, (cudaStream_t)s>>>(x0_mean, unet_grad, direct, coef, c_in, hat_x0, CHW / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_lincomb(const float* x, const float* y, const float* a, const float* c, float* out, int B, int CHW,
                            kdip_stream_t s) {
  REQ_ALIGN16(x); REQ_ALIGN16(y); REQ_ALIGN16(out); REQ_MULT4(CHW);
  KDIP_REQUIRE(B > 0 && a != nullptr && (y == nullptr || c != nullptr), KDIP_EINVAL, "lincomb: bad argument");
  lincomb_kernel<<<grid_by(CHW / 4, B), EW_THREADS, 0, (cudaStream_t)s>>>(x, y, a, c, out, CHW / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_v2_epilogue(const float* unet_out, const float* cov_out, const float* x, const float* sigma, float* x0_mean,
                                float* x0_var, float* theta0_var, int B, int HW, kdip_stream_t s) {
  REQ_ALIGN16(unet_out); REQ_ALIGN16(cov_out); REQ_ALIGN16(x); REQ_ALIGN16(x0_mean); REQ_ALIGN16(x0_var); REQ_ALIGN16(theta0_var);
  REQ_MULT4(HW);
  KDIP_REQUIRE(B > 0 && sigma != nullptr, KDIP_EINVAL, "v2_epilogue: bad argument");
  KDIP_REQUIRE((x0_var == nullptr) == (theta0_var == nullptr), KDIP_EINVAL, "v2_epilogue: x0_var and theta0_var go together");
  KDIP_REQUIRE(x0_var == nullptr || cov_out != nullptr, KDIP_EINVAL, "v2_epilogue: variances need cov_out");
  v2_epilogue_kernel<<<grid_by(3 * HW / 4, B), EW_THREADS, 0, (cudaStream_t)s>>>(unet_out, cov_out, x, sigma, x0_mean, x0_var,
                                                                                   theta0_var, HW / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_inpaint_forward(const float* x, const float* noise, const float* mask, float sigma_s, float* y, int B,
                                    int CHW, kdip_stream_t s) {
  REQ_ALIGN16(x); REQ_ALIGN16(noise); REQ_ALIGN16(mask); REQ_ALIGN16(y); REQ_MULT4(CHW);
  KDIP_REQUIRE(B > 0, KDIP_ESHAPE, "inpaint_forward: B must be > 0");
  inpaint_fwd_kernel<<<grid_by(CHW / 4, B), EW_THREADS, 0, (cudaStream_t)s>>>(x, noise, mask, sigma_s, y, CHW / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_inpaint_mat_scalar(const float* y, const float* x0, const float* mask, const float* theta, float sigma_s,
                                       float* mat, int B, int CHW, kdip_stream_t s) {
  REQ_ALIGN16(y); REQ_ALIGN16(x0); REQ_ALIGN16(mask); REQ_ALIGN16(mat); REQ_MULT4(CHW);
  KDIP_REQUIRE(B > 0, KDIP_ESHAPE, "inpaint_mat_scalar: B must be > 0");
  inpaint_mat_kernel<<<grid_by(CHW / 4, B), EW_THREADS, 0, (cudaStream_t)s>>>(y, x0, mask, theta, sigma_s * sigma_s, mat, CHW / 4);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_gather(const float* src, const int32_t* idx, float* dst, int B, int CHW, int M, kdip_stream_t s) {
  KDIP_REQUIRE(B > 0 && M >= 0, KDIP_ESHAPE, "gather: bad sizes");
  if (M == 0) return KDIP_OK;
  gather_kernel<<<grid_by(M, B), EW_THREADS, 0, (cudaStream_t)s>>>(src, idx, dst, CHW, M);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}
extern "C" int kdip_scatter(const float* src, const int32_t* idx, float* dst, int B, int CHW, int M, kdip_stream_t s) {
  KDIP_REQUIRE(B > 0 && M >= 0, KDIP_ESHAPE, "scatter: bad sizes");
  KDIP_CUDA(cudaMemsetAsync(dst, 0, (size_t)B * CHW * sizeof(float), (cudaStream_t)s));
  if (M == 0) return KDIP_OK;
  scatter_kernel<<<grid_by(M, B), EW_THREADS, 0, (cudaStream_t)s>>>(src, idx, dst, CHW, M);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}
