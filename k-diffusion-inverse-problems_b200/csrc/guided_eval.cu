// kdip_guided_eval: ONE call = one guided model evaluation of the sampling loop (condition/condition.py:83-174 on top of
// ConditionOpenAIDenoiser.uncond_pred, :231-274) for the branches that need no iterative solver:
//     UNet forward -> p_mean_variance epilogue -> mat (closed form, scalar x0 variance) or DPS residual gradient
//     -> clamp / scaling VJP seed -> UNet input-VJP -> hat_x0 = clip(x0 + coef (c_in g + direct), -1, 1)
// Everything is enqueued on the caller's stream from the caller's workspace: no allocation, no synchronisation.  The call has two
// halves, also exported on their own: kdip_guided_eval_set broadcasts the evaluation's scalars (passed BY VALUE as kernel
// arguments - no host memory is read after the call returns) into per-image device arrays of the workspace, and
// kdip_guided_eval_run does everything else from device-resident data only, so _run can be captured ONCE into a CUDA graph per
// (guidance, batch) and replayed for every sigma of the schedule after an eager _set.  (A first version copied a pinned host
// struct inside the graph; a replay then reads the struct when the GPU gets there, by which time the host may already have
// written the next evaluation's scalars - measured as 2-18 % trajectory errors as soon as the host ran ahead.)
// The per-pixel-covariance / CG branch (condition.py:325-346,359-384,412-437) polls convergence on the host and stays with the
// individually exported pieces (kdip_mat_cg etc.), which is also what user-registered operators / mat solvers use.
#include "kdip_common.cuh"

namespace kdip {

struct GeWs {
  void *unet_ws, *op_ws;
  size_t unet_bytes, op_bytes;
  float *out6, *x0, *mat, *seed, *direct, *grad;
  kdip_guided_cfg* cfg;
  kdip_pmv_scalars* sc;
  float *c_in, *t, *theta, *coef, *norm;
};

struct Bump {
  char* base;
  size_t cur = 0;
  void* take(size_t bytes) {
    const size_t o = cur;
    cur += (bytes + 255) & ~(size_t)255;
    return base ? (void*)(base + o) : nullptr;
  }
};

static int plan(kdip_unet* u, const kdip_op* op, int B, void* base, size_t* total, GeWs* w) {
  size_t ub = 0, ob = 0;
  int rc = kdip_unet_workspace_bytes(u, B, &ub);
  if (rc) return rc;
  rc = kdip_op_workspace_bytes(op, B, &ob);
  if (rc) return rc;
  const int S = kdip_op_side(op);
  const size_t plane = (size_t)B * 3 * S * S * sizeof(float);
  Bump a{(char*)base};
  w->unet_bytes = ub; w->op_bytes = ob;
  w->unet_ws = a.take(ub);
  w->op_ws = a.take(ob);
  w->out6 = (float*)a.take(2 * plane);
  w->x0 = (float*)a.take(plane);
  w->mat = (float*)a.take(plane);
  w->seed = (float*)a.take(2 * plane);
  w->direct = (float*)a.take(plane);
  w->grad = (float*)a.take(plane);
  w->cfg = (kdip_guided_cfg*)a.take(sizeof(kdip_guided_cfg));
  w->sc = (kdip_pmv_scalars*)a.take((size_t)B * sizeof(kdip_pmv_scalars));
  w->c_in = (float*)a.take((size_t)B * 4);
  w->t = (float*)a.take((size_t)B * 4);
  w->theta = (float*)a.take((size_t)B * 4);
  w->coef = (float*)a.take((size_t)B * 4);
  w->norm = (float*)a.take((size_t)B * 4);
  *total = a.cur;
  return KDIP_OK;
}

// broadcast the evaluation's scalars (uniform over the batch inside a sampler call) into the per-image device arrays
__global__ void ge_fill_kernel(const kdip_guided_cfg c, kdip_guided_cfg* __restrict__ cfg_dev, kdip_pmv_scalars* __restrict__ sc,
                               float* __restrict__ c_in, float* __restrict__ t, float* __restrict__ theta, float* __restrict__ coef, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (b == 0) *cfg_dev = c;
  sc[b] = c.sc;
  c_in[b] = c.sc.c_in;
  t[b] = c.t_model;
  theta[b] = c.theta;
  const float s2 = c.sigma * c.sigma;
  float k = 0.f;
  if (c.guidance == KDIP_GUIDE_TYPE_I) k = s2;                                   // condition.py:173
  else if (c.guidance == KDIP_GUIDE_PGDM) k = s2 * c.theta;                      // :155-156 (theta = r^2)
  else if (c.guidance == KDIP_GUIDE_DIFFPIR) k = c.theta;                        // :164
  coef[b] = k;
}
// DPS: coef[b] = sigma^2 zeta / ||r_b||                                          condition.py:145-147
__global__ void ge_dps_coef_kernel(const kdip_guided_cfg* __restrict__ cfg, const float* __restrict__ norm, float* __restrict__ coef, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) coef[b] = cfg->sigma * cfg->sigma * cfg->zeta / norm[b];
}

}  // namespace kdip

using namespace kdip;

extern "C" int kdip_guided_eval_workspace_bytes(kdip_unet* u, const kdip_op* op, int B, size_t* bytes) {
  KDIP_REQUIRE(u && op && bytes && B > 0, KDIP_EINVAL, "guided_eval_workspace_bytes: bad argument");
  GeWs w;
  return plan(u, op, B, nullptr, bytes, &w);
}

static int get_ws(kdip_unet* u, const kdip_op* op, int B, void* ws, size_t ws_bytes, GeWs* w) {
  KDIP_REQUIRE(ws != nullptr && ((uintptr_t)ws % 256) == 0, KDIP_EALIGN, "guided_eval: workspace must be 256-byte aligned");
  size_t need = 0;
  int rc = plan(u, op, B, ws, &need, w);
  if (rc) return rc;
  KDIP_REQUIRE(need <= ws_bytes, KDIP_ENOMEM, "guided_eval: workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  return KDIP_OK;
}

extern "C" int kdip_guided_eval_set(kdip_unet* u, const kdip_op* op, const kdip_guided_cfg* cfg, int B, void* ws, size_t ws_bytes,
                                    kdip_stream_t stream) {
  KDIP_REQUIRE(u && op && cfg && B > 0, KDIP_EINVAL, "guided_eval_set: bad argument");
  KDIP_REQUIRE(cfg->guidance >= KDIP_GUIDE_UNCOND && cfg->guidance <= KDIP_GUIDE_DIFFPIR, KDIP_EINVAL, "guided_eval: unknown guidance %d",
               cfg->guidance);
  GeWs w;
  int rc = get_ws(u, op, B, ws, ws_bytes, &w);
  if (rc) return rc;
  ge_fill_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*cfg, w.cfg, w.sc, w.c_in, w.t, w.theta, w.coef, B);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

extern "C" int kdip_guided_eval_run(kdip_unet* u, kdip_op* op, int guidance, const float* x, const float* y, float* hat_x0, int B,
                                    void* ws, size_t ws_bytes, kdip_stream_t stream) {
  KDIP_REQUIRE(u && op && x && y && hat_x0 && B > 0, KDIP_EINVAL, "guided_eval_run: bad argument");
  KDIP_REQUIRE(guidance >= KDIP_GUIDE_UNCOND && guidance <= KDIP_GUIDE_DIFFPIR, KDIP_EINVAL, "guided_eval: unknown guidance %d", guidance);
  GeWs w;
  int rc = get_ws(u, op, B, ws, ws_bytes, &w);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = kdip_op_side(op), HW = S * S, CHW = 3 * HW;
  const int g = guidance;
  rc = kdip_unet_forward(u, x, w.c_in, w.t, B, w.out6, nullptr, w.unet_ws, w.unet_bytes, stream);
  if (rc) return rc;
  rc = kdip_pmv_epilogue(w.out6, x, w.sc, w.x0, nullptr, 0, B, HW, stream);
  if (rc) return rc;
  if (g == KDIP_GUIDE_UNCOND) return kdip_guidance_combine(w.x0, w.x0, nullptr, w.coef, nullptr, hat_x0, B, CHW, stream);
  if (g == KDIP_GUIDE_DPS) {
    rc = kdip_dps_grad(op, y, w.x0, w.mat, w.norm, B, w.op_ws, w.op_bytes, stream);
    if (rc) return rc;
    ge_dps_coef_kernel<<<(B + 127) / 128, 128, 0, st>>>(w.cfg, w.norm, w.coef, B);
    KDIP_LAUNCH_CHECK();
  } else {
    rc = kdip_mat_closed(op, y, w.x0, w.theta, w.mat, B, w.op_ws, w.op_bytes, stream);
    if (rc) return rc;
    if (g == KDIP_GUIDE_DIFFPIR) return kdip_guidance_combine(w.x0, w.mat, nullptr, w.coef, nullptr, hat_x0, B, CHW, stream);
  }
  rc = kdip_pmv_vjp_seed(w.x0, w.mat, w.sc, w.seed, w.direct, B, HW, stream);
  if (rc) return rc;
  rc = kdip_unet_vjp(u, w.seed, B, w.grad, w.unet_ws, w.unet_bytes, stream);
  if (rc) return rc;
  return kdip_guidance_combine(w.x0, w.grad, w.direct, w.coef, w.c_in, hat_x0, B, CHW, stream);
}

extern "C" int kdip_guided_eval(kdip_unet* u, kdip_op* op, const kdip_guided_cfg* cfg, const float* x, const float* y, float* hat_x0,
                                int B, void* ws, size_t ws_bytes, kdip_stream_t stream) {
  int rc = kdip_guided_eval_set(u, op, cfg, B, ws, ws_bytes, stream);
  if (rc) return rc;
  return kdip_guided_eval_run(u, op, cfg->guidance, x, y, hat_x0, B, ws, ws_bytes, stream);
}
