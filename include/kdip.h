/* libkdip — C ABI of the B200-native guided-sampling hot path (drop-in boundary).
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (pass 0 for the legacy default
 * stream); none allocates device memory after *_create, none synchronises, all are safe under CUDA-graph
 * capture unless stated.  The caller owns every I/O buffer and workspace.  All image tensors at this boundary
 * are NCHW contiguous fp32, exactly what the reference's Python passes around.
 *
 * Return value: 0 on success, <0 = KDIP_E*; kdip_last_error() returns a thread-local message.
 *
 * Each function cites the reference interface (file:line under /root/reference) it replaces.
 */
#ifndef KDIP_H_
#define KDIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDIP_OK 0
#define KDIP_EINVAL (-1)   /* bad argument / unsupported configuration            */
#define KDIP_ESHAPE (-2)   /* shape not supported by the kernels                   */
#define KDIP_EALIGN (-3)   /* pointer alignment                                    */
#define KDIP_ECUDA (-4)    /* CUDA runtime / driver error (message has details)    */
#define KDIP_ENOTCONV (-5) /* CG hit maxiter (reference: warnings.warn, non-fatal) */
#define KDIP_ENOMEM (-6)   /* caller workspace too small                           */

typedef void* kdip_stream_t; /* cudaStream_t */

const char* kdip_last_error(void);
int kdip_version(void);
/* Number of CUDA kernels this library has launched in the process so far (bench.py reports the delta as gpu_launches). */
unsigned long long kdip_launch_count(void);
/* A caller that replays a captured CUDA graph of library calls bypasses the library's own launch path: it reports the kernel
 * nodes of each replay here (the count kdip_launch_count advanced by while the graph was captured). */
void kdip_launch_count_add(size_t n);
/* Diagnostics: a conv_gemm_kernel wait that times out (a pipeline-protocol bug; the kernel then traps and the context is lost)
 * first leaves a record in host-mapped memory.  Copies up to n (<= 64) words: out[0] = number of timed-out waits, then from
 * out[4] on four words per wait (source line in csrc/conv_gemm.cu, block, thread | parity << 16, barrier shared-memory address).
 * Returns out[0]; 0 = no record.  Usable after the CUDA context has failed. */
int kdip_conv_trap_read(unsigned int* out, int n);
/* Device sanity: returns KDIP_OK when the current device is sm_100 (B200); fills sm_count if non-null. */
int kdip_device_check(int* sm_count);

/* ------------------------------------------------------------------------------------------------------------
 * Sampler elementwise updates — k_diffusion/sampling.py:118-184 (sample_euler / sample_heun), to_d :46-48.
 * x, denoised, d, noise: [B,3,H,W] fp32, n = B*3*H*W elements.  Sigmas are host scalars (the schedule lives on
 * the host, sampling.py:17-23), so no device->host sync is needed for the branch decisions.
 * ------------------------------------------------------------------------------------------------------------ */
/* x += noise * s_noise * sqrt(sigma_hat^2 - sigma^2)                                  (sampling.py:124-127,165-168) */
int kdip_churn(float* x, const float* noise, float s_noise, float sigma, float sigma_hat, size_t n, kdip_stream_t s);
/* d = (x - denoised)/sigma_hat ; x_out = x + d*dt ; d_out may be NULL (Euler)          (sampling.py:129-134,170-179) */
int kdip_euler_step(const float* x, const float* denoised, float sigma_hat, float dt, float* x_out, float* d_out,
                    size_t n, kdip_stream_t s);
/* out = a*x + b*y + c*z, host scalars, products and sums rounded separately; y / z may be NULL; out may alias an input.
 * The state update of the remaining k_diffusion/sampling.py samplers (sample_euler_ancestral :139-156, sample_dpm_2 :187-215,
 * sample_dpm_2_ancestral :218-248, sample_lms :259-275, sample_dpmpp_2s_ancestral :507-538, sample_dpmpp_2m :583-606).       */
int kdip_lincomb3(const float* x, const float* y, const float* z, float a, float b, float c, float* out, size_t n,
                  kdip_stream_t s);
/* d2 = (x2 - denoised2)/sigma_next ; x_out = x + (d + d2)/2 * dt                       (sampling.py:180-183) */
int kdip_heun_step(const float* x, const float* d, const float* x2, const float* denoised2, float sigma_next, float dt,
                   float* x_out, size_t n, kdip_stream_t s);

/* ------------------------------------------------------------------------------------------------------------
 * p_mean_variance epilogue + x0-covariance — guided_diffusion/gaussian_diffusion.py:262-276,296-297,305-311,
 * condition/condition.py:232-248.  Per-image scalar tables (host floats, length B): the caller extracts them
 * from the float64 schedule at integer t (gaussian_diffusion.py:895-908).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  float c_in;          /* 1/sqrt(sigma^2+1)                      k_diffusion/external.py:97-100 */
  float recip;         /* sqrt(1/abar_t)                          gaussian_diffusion.py:145      */
  float recipm1;       /* sqrt(1/abar_t - 1)                      gaussian_diffusion.py:146      */
  float min_log;       /* posterior_log_variance_clipped[t]       :154-158                       */
  float max_log;       /* log(beta_t)                             :271                           */
  float post_var;      /* posterior_variance[t]  (Convert Eq.22)  condition.py:244               */
  float coef1_sq;      /* posterior_mean_coef1[t]^2 (fp32 square) condition.py:245               */
} kdip_pmv_scalars;

#define KDIP_VAR_MODEL 1   /* p_mean_variance's 'variance' (LEARNED_RANGE), gaussian_diffusion.py:271-276 */
#define KDIP_VAR_CONVERT 2 /* Convert x0 variance, Eq. (22): clip((variance - post_var)/coef1_sq, 1e-6), condition.py:243-246 */
/* unet_out [B,6,HW], x [B,3,HW] (UNSCALED x_t; the kernel applies c_in) ->
 *   x0_mean [B,3,HW] = clamp(recip*c_in*x - recipm1*eps, -1, 1)
 *   x0_var  [B,3,HW] (may be NULL) = variance = exp(frac*max_log+(1-frac)*min_log), frac = (v+1)/2   (KDIP_VAR_MODEL)
 *                                    or its Eq. (22) conversion                                    (KDIP_VAR_CONVERT)
 * sc: device array of B kdip_pmv_scalars. */
int kdip_pmv_epilogue(const float* unet_out, const float* x, const kdip_pmv_scalars* sc, float* x0_mean, float* x0_var,
                      int var_mode, int B, int HW, kdip_stream_t s);
/* VJP seed for the UNet output given v = d(loss)/d(x0_mean): writes seed [B,6,HW] = (-recipm1*m*v, 0) with
 * m = 1 where the clamp is inactive (x0_mean strictly inside (-1,1) reproduces torch.clamp's gradient), and
 * direct [B,3,HW] = recip*c_in*m*v (the d/dx of the explicit recip*c_in*x term; NULL to skip).
 * x0_mean == NULL: no clamp (m = 1) - the v2 denoiser x0 = x - sigma*eps (condition/condition.py:287-291) with
 * sc = {c_in 1, recip 1, recipm1 sigma}: seed = (-sigma*v, 0), direct = v. */
int kdip_pmv_vjp_seed(const float* x0_mean, const float* v, const kdip_pmv_scalars* sc, float* seed,
                      float* direct, int B, int HW, kdip_stream_t s);

/* ------------------------------------------------------------------------------------------------------------
 * Guidance combine — condition/condition.py:131,137,147,156,164,173:
 *   hat_x0 = clip(x0_mean + coef[b] * (c_in[b]*unet_grad + direct), -1, 1)   (types I / PiGDM / DPS)
 *   hat_x0 = clip(x0_mean + coef[b] * mat, -1, 1)                            (DiffPIR: pass unet_grad=mat, direct=NULL)
 * coef, c_in: device arrays of B floats (c_in may be NULL -> 1).
 * ------------------------------------------------------------------------------------------------------------ */
int kdip_guidance_combine(const float* x0_mean, const float* unet_grad, const float* direct, const float* coef,
                          const float* c_in, float* hat_x0, int B, int CHW, kdip_stream_t s);

/* out = a[b]*x + c[b]*y, per-image device scalars, no clipping (y, c may be NULL).  TMPD covariance
 * sigma^2 * grad_x sum(x0_mean) = (sigma^2 c_in) * unet_grad + sigma^2 * direct                 condition.py:268-269 */
int kdip_lincomb(const float* x, const float* y, const float* a, const float* c, float* out, int B, int CHW, kdip_stream_t s);

/* v2 (DWT-Var) epilogue — condition/condition.py:287-300, k_diffusion/external.py:161-169:
 *   x0_mean = unet_out[:, :3]*c_out + x, c_out = -sigma[b];  x0_var = exp(cov_out[:, :3])*c_out^2 ;
 *   theta0_var = exp(cov_out[:, 3:])*c_out^2 (both NULL to skip; cov_out may then be NULL).  sigma: device [B]. */
int kdip_v2_epilogue(const float* unet_out, const float* cov_out, const float* x, const float* sigma, float* x0_mean,
                     float* x0_var, float* theta0_var, int B, int HW, kdip_stream_t s);

/* ------------------------------------------------------------------------------------------------------------
 * Inpainting — condition/measurements.py:211-238, condition/condition.py:317-323.
 * mask: [3,HW] fp32 0/1 shared by the batch (measurements.py:209).
 * ------------------------------------------------------------------------------------------------------------ */
/* y = (x + sigma_s*noise) * mask ; noise may be NULL (noiseless)                         measurements.py:211-215 */
int kdip_inpaint_forward(const float* x, const float* noise, const float* mask, float sigma_s, float* y, int B, int CHW,
                         kdip_stream_t s);
/* closed form: mat = (mask*y - mask*x0)/(sigma_s^2 + theta[b]) ; sigma_s already clipped to >=1e-3 by the caller */
int kdip_inpaint_mat_scalar(const float* y, const float* x0, const float* mask, const float* theta, float sigma_s,
                            float* mat, int B, int CHW, kdip_stream_t s);
/* flatten: gather y at mask>0 positions (torch.where order: c,h,w) — idx [M] int32 ascending offsets into CHW  :217-219 */
int kdip_gather(const float* src, const int32_t* idx, float* dst, int B, int CHW, int M, kdip_stream_t s);
/* transpose(flatten=True): scatter into zeros                                                                :231-234 */
int kdip_scatter(const float* src, const int32_t* idx, float* dst, int B, int CHW, int M, kdip_stream_t s);

/* ------------------------------------------------------------------------------------------------------------
 * Data front end and evaluation reductions: the callers either side of the sampling path.
 * 8-bit images are interleaved HWC (PIL / PNG order), fp32 images NCHW in [-1, 1]; H*W must be a multiple of 4.
 */
/* Dataset transform of sample_condition_openai.py:140-144 (torchvision ToTensor, then x*2-1): dst = u8/255*2-1. Bit-exact. */
int kdip_images_u8_to_f32(const uint8_t* src_hwc, float* dst_nchw, int B, int H, int W, kdip_stream_t s);
/* k_diffusion/utils.py:24-31 to_pil_image: dst = trunc((clamp(x,-1,1)+1)/2*255) (torchvision mul(255).byte()). Bit-exact. */
int kdip_images_f32_to_u8(const float* src_nchw, uint8_t* dst_hwc, int B, int H, int W, kdip_stream_t s);
/* out[b] (fp64, device) = sum over the image of (a-b)^2.  to_eval_first != 0 maps both through (x/2+0.5).clip(0,1) and
 * subtracts in fp64 (PSNR of sample_condition_openai.py:41-44 = 10 log10(CHW / out[b])); 0 = fp32 differences
 * (analytic_variance.py:129).  Zeroes `out` itself. */
int kdip_sqerr_sum(const float* a, const float* b, int to_eval_first, double* out, int B, int CHW, kdip_stream_t s);
/* analytic_variance.py:128-129 fused: hat_x0 = x_noised + eps*(-sigma[b]) with eps = channels 0..2 of unet_out [B,6,H,W]
 * (OpenAIDenoiser.forward, k_diffusion/external.py:111-132); out[b] = sum (x0 - hat_x0)^2 (fp64); hat_x0 written if non-NULL. */
int kdip_denoise_sqerr(const float* unet_out, const float* x_noised, const float* x0, const float* sigma, double* out,
                       float* hat_x0, int B, int HW, kdip_stream_t s);
/* out[b] = sum over 3 channels and all window-valid centres of the SSIM map of to_eval(a), to_eval(b)
 * (skimage.metrics.structural_similarity(channel_axis=0, data_range=1), sample_condition_openai.py:45: 7x7 uniform window,
 * sample covariance, K1 0.01, K2 0.03); SSIM = out[b] / (3 (H-6) (W-6)). */
int kdip_ssim_sum(const float* a, const float* b, double* out, int B, int H, int W, kdip_stream_t s);

/* ------------------------------------------------------------------------------------------------------------
 * Measurement operators and mat solvers — condition/measurements.py:86-244 (operators), condition/condition.py:317-439
 * (inpainting_mat / gaussian_blur_mat / motion_blur_mat / super_resolution_mat), condition/utils.py:50-139
 * (OrthoTransform), condition/diffpir_utils/utils_sisr.py:9-96, condition/dps_utils/resizer.py:8-198.
 * Images are fp32 NCHW [B,3,S,S]; SR measurements [B,3,S/sf,S/sf].  The operator handle owns the OTF / mask / Resizer
 * tables (built at create from HOST arrays); all per-call scratch is caller workspace (kdip_op_workspace_bytes).
 * ------------------------------------------------------------------------------------------------------------ */
#define KDIP_OP_INPAINTING 0
#define KDIP_OP_GAUSSIAN_BLUR 1
#define KDIP_OP_MOTION_BLUR 2
#define KDIP_OP_SUPER_RESOLUTION 3
#define KDIP_OT_NONE 0 /* condition/utils.py:50-67: identity */
#define KDIP_OT_DCT 1  /* :89-103  scipy.fft.dctn(norm='ortho') over (C,H,W) */
#define KDIP_OT_DWT 2  /* :107-139 pywt haar level 3, coeffs_to_array packing  */

typedef struct {
  int kind;             /* KDIP_OP_*                                                                       */
  int S;                /* image side, power of two in [16,256]                                            */
  int sf;               /* super_resolution scale factor (measurements.py:89), else ignored               */
  float sigma_s;        /* measurement noise std (operator.sigma_s)                                        */
  const float* psf;     /* HOST fp32 [ksize][ksize]: blur kernel (get_kernel(), measurements.py:158,198) or the
                           bicubic SR kernel (measurements.py:95-97); NULL for inpainting                  */
  int ksize;
  const float* mask;    /* HOST fp32 [3][S][S] 0/1 (measurements.py:205-209); inpainting only              */
  const float* rs_w;    /* HOST Resizer weights [S/sf][rs_taps] (resizer.py:104-168), super_resolution only */
  const int32_t* rs_idx;/* HOST Resizer field of view [S/sf][rs_taps]                                      */
  int rs_taps;
} kdip_op_desc;

typedef struct kdip_op kdip_op;
int kdip_op_create(const kdip_op_desc* d, kdip_op** out);   /* allocates + synchronises (setup only) */
void kdip_op_destroy(kdip_op* op);
int kdip_op_workspace_bytes(const kdip_op* op, int B, size_t* bytes);
int kdip_op_side(const kdip_op* op);   /* image side S the operator was created for */
/* operator.forward(data, noiseless = (noise == NULL)): y = A x + sigma_s * noise          measurements.py:103-111,139-148,
 * 178-188,211-226.  noise has y's shape (drawn by the caller, torch.randn_like in the reference). */
int kdip_op_forward(const kdip_op* op, const float* x, const float* noise, float* y, int B, void* ws, size_t ws_bytes,
                    kdip_stream_t s);
/* operator.transpose(y): blur A^T y (conj OTF), SR ifft2(conj(FB) fft2(upsample(y))), inpainting identity
 *                                                                                       measurements.py:113-122,150-156,190-196,228-238 */
int kdip_op_transpose(const kdip_op* op, const float* y, float* x, int B, void* ws, size_t ws_bytes, kdip_stream_t s);
/* Adjoint of operator.forward(noiseless): what autograd computes in LinearOperator.auto_transpose (measurements.py:48-52) and in
 * the DPS branch (condition.py:144-146).  g has y's shape.  Blur: conj OTF; SR: the Resizer's adjoint; inpainting: mask. */
int kdip_op_forward_adjoint(const kdip_op* op, const float* g, float* x, int B, void* ws, size_t ws_bytes, kdip_stream_t s);
/* fft2 over (H, W) of x [B,3,S,S] -> interleaved complex64 [B,3,S,S]: the torch.fft.fftn of utils_sisr.py:91-95 behind the FBFy
 * member of operator.pre_calculated.  Spectral operators only. */
int kdip_op_fft2(const kdip_op* op, const float* x, float* out_full, int B, void* ws, size_t ws_bytes, kdip_stream_t s);
/* FB of pre_calculate (utils_sisr.py:79-96) as interleaved complex64 [S][S] (device), for operator.pre_calculated */
int kdip_op_otf(const kdip_op* op, float* fb_full, kdip_stream_t s);
/* closed-form mat for scalar x0 variance theta[b] (device, B floats)                   condition.py:322-323,356-357,408-410 */
int kdip_mat_closed(const kdip_op* op, const float* y, const float* x0, const float* theta, float* mat, int B, void* ws,
                    size_t ws_bytes, kdip_stream_t s);
/* CG mat for a per-element variance map theta_map [B,3,S,S] in the domain of transform `ot`  condition.py:325-346,359-384,412-437.
 * Batched on-device CG with scipy's semantics (x0 = 0, stop when ||r|| < tol*||b||, at most maxiter matvecs), per-image
 * convergence; alpha / beta / the convergence flags stay on the device.  iters_out: HOST int[B] or NULL.  The host polls the flags
 * with a one-iteration lag through pinned snapshots owned by the handle (event waits: not graph-capturable; at most one surplus
 * iteration is queued after the last image converged).  B <= 4096.
 * Returns KDIP_ENOTCONV (result still written, like the reference's warning) when an image hit maxiter. */
int kdip_mat_cg(kdip_op* op, const float* y, const float* x0, const float* theta_map, int ot, float* mat, int B, float tol,
                int maxiter, int* iters_out, void* ws, size_t ws_bytes, kdip_stream_t s);
/* DPS (condition.py:140-148): r = y - forward(x0, noiseless); norm[b] = ||r_b||_2 (device, B floats); v = A^T r [B,3,S,S] */
int kdip_dps_grad(const kdip_op* op, const float* y, const float* x0, float* v, float* norm, int B, void* ws,
                  size_t ws_bytes, kdip_stream_t s);
/* OrthoTransform: out = mul .* W^T x (inverse = 0; mul [B,3,S,S] or NULL) or out = W x (inverse = 1)   condition/utils.py:69-77 */
int kdip_ortho(int ot, int inverse, const float* x, const float* mul, float* out, int B, int S, void* ws, size_t ws_bytes,
               kdip_stream_t s);
int kdip_ortho_workspace_bytes(int ot, int B, int S, size_t* bytes);

/* ------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on the tcgen05 tensor cores (the UNet's conv3x3 / conv1x1 / qkv / proj and their
 * input-gradients) — replaces cuDNN under guided_diffusion/unet.py:185,211,222,287,295,617.
 * Activations are bf16 NHWC; weights bf16 [taps*Cout][Cin] (K-major), one row block per tap.
 * out[n,y,x,:] = bias + residual + sum_seg sum_tap W_seg[tap] . act_seg[n, y+dy(tap), x+dx(tap), :]
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* act; /* bf16 [N,H,W,C]                         */
  int C;           /* multiple of 64                          */
  const void* wgt; /* bf16 [taps*Cout_pad][C]                 */
  int taps;        /* 9 (3x3, pad 1, tap = ky*3+kx) or 1      */
} kdip_conv_seg;

typedef struct {
  int N, H, W;
  int Cout_pad;  /* GEMM N: multiple of 16 (<=256) or of 64; rows per tap in wgt                 */
  int Cout;      /* real output channels written (<= Cout_pad)                                   */
  int nseg;      /* 1..3 K-segments accumulated into the same output (e.g. conv3x3 + 1x1 skip)   */
  kdip_conv_seg seg[3];
  const float* bias;    /* [Cout_pad] fp32 or NULL                                               */
  const void* residual; /* bf16 NHWC with Cout channels or NULL                                  */
  int res_mode;         /* 0 none, 1 same resolution, 2 avg-pool 2x2 of [N,2H,2W,C], 3 nearest-up of [N,H/2,W/2,C] */
  void* out;            /* out_mode 0: bf16 [N,H,W,Cout]; 1: fp32 [N,Cout,H,W]                   */
  int out_mode;
  float out_scale;      /* multiplies the final value (out_mode 1 only; e.g. c_in for the input-VJP) */
  float* chan_stats;    /* optional fp32 [N,Cout,2] (sum, sum of squares) accumulated atomically; NULL to skip */
  /* Optional fused GroupNorm-backward reduction (autograd of nn.py:17-19 + SiLU; condition/condition.py:136,146,155,172,269).
   * The conv output g is the gradient wrt the normalised activation act(A x + B) of the tensor x = concat(gn_x0, gn_x1)
   * (bf16 NHWC at the output resolution, gn_C0 + C1 = Cout, both multiples of 64):
   *   gn_red[n][c][0] += sum_p g_u,  gn_red[n][c][1] += sum_p g_u x,  g_u = g act'(A x + B)   (g as stored, i.e. bf16-rounded)
   * Needs bf16 NHWC output, no residual, Cout a multiple of 64 and image sides that are multiples of the pixel tile. */
  const void* gn_x0;
  const void* gn_x1;    /* or NULL */
  int gn_C0;
  int gn_silu;          /* 1: act = SiLU, 0: identity */
  const float* gn_ab;   /* fp32 [N,Cout,2] folded affine (A, B) of the forward pass */
  float* gn_red;        /* fp32 [N,Cout,2], zeroed by the caller; NULL to skip */
  /* Optional fused GroupNorm apply on the operand path (forward of nn.py:17-19 + SiLU in front of the conv, unet.py:237-257 in_layers /
   * out_layers): segment s is consumed as act(A x + B) with (A, B) = in_ab[s][(n * in_ab_C + c) * 2 + {0,1}] for channel c of the
   * segment (in_ab[s] already points at the segment's first channel inside a [N, in_ab_C, 2] table; NULL = segment read as is).  The
   * values are rounded to bf16 exactly as kdip_gn_apply stores them and the conv pads the NORMALISED tensor with zeros.  Needs the
   * row-tile ("halo") pipeline: 3x3 first segment, W a multiple of 128, even H, bf16 NHWC output with Cout a multiple of 64. */
  const float* in_ab[3];
  int in_ab_C;
  int in_silu;          /* 1: act = SiLU, 0: identity */
} kdip_conv_desc;

typedef struct kdip_conv_plan kdip_conv_plan; /* encoded TMA descriptors + launch geometry */
int kdip_conv_plan_create(const kdip_conv_desc* d, kdip_conv_plan** out);
int kdip_conv_plan_run(const kdip_conv_plan* p, kdip_stream_t s);
void kdip_conv_plan_destroy(kdip_conv_plan* p);

/* fp32 OIHW (or [O][I] / [O][I][1]) -> packed bf16 [taps*rows_pad][cols_pad], zero padded.
 * flip_transpose=0: rows = output channels, cols = input channels (forward operator).
 * flip_transpose=1: rows = input channels, cols = output channels, taps reversed:
 *   W'[tap'][ci][co] = W[co][ci][taps-1-tap'] — the input-gradient (dgrad) operator of a stride-1 pad-1 conv. */
int kdip_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int taps, int rows_pad, int cols_pad,
                          int flip_transpose, void* dst_bf16, kdip_stream_t s);

/* ------------------------------------------------------------------------------------------------------------
 * ADM UNet denoiser — guided_diffusion/unet.py:398-668 (UNetModel), built as create_model does
 * (guided_diffusion/script_util.py:130-184) with condition/diffpir_utils/utils_model.py:353-387 defaults:
 * resblock_updown, use_scale_shift_norm, learn_sigma (6 output channels), legacy attention order, head width 64.
 * ------------------------------------------------------------------------------------------------------------ */
#define KDIP_PRECISION_BF16 0
#define KDIP_PRECISION_FP32 1
typedef struct {
  int image_size;         /* 256 (FFHQ / ImageNet) or 64 (test geometry)                      */
  int in_channels;        /* 3                                                                */
  int model_channels;     /* num_channels: 128 (FFHQ), 256 (ImageNet)                         */
  int out_channels;       /* 6 (learn_sigma)                                                  */
  int num_res_blocks;     /* 1 (FFHQ), 2 (ImageNet)                                           */
  int num_head_channels;  /* 64                                                               */
  int n_mult;             /* levels                                                           */
  float channel_mult[8];  /* script_util.py:148-160                                           */
  int n_att;
  int attention_ds[8];    /* downsample rates with attention, script_util.py:162-164          */
  int precision;          /* KDIP_PRECISION_BF16 (0): bf16 tcgen05 operands, fp32 accumulation - the fast path;
                             KDIP_PRECISION_FP32 (1): fp32 FMA on fp32 operands throughout, the reference's own arithmetic
                             (use_fp16=False, condition/diffpir_utils/utils_model.py:364) - CUDA cores, ~40x slower, for
                             tight-tolerance parity against the reference                  */
} kdip_unet_arch;

typedef struct kdip_unet kdip_unet;

/* names[i] / ptrs[i] / numels[i]: the reference state_dict (UNetModel.load_state_dict keys), fp32 tensors resident on
 * the current device.  The handle packs its own bf16 copies; the caller keeps ownership of the sources.  An optional
 * pair "out_cov.weight" [6,C,1,1] / "out_cov.bias" [6] enables the DWT-Var covariance head of
 * k_diffusion/external.py:141,163-165.  Allocates device memory (weights only); synchronises the legacy stream. */
int kdip_unet_create(const kdip_unet_arch* arch, int n_tensors, const char* const* names, const float* const* ptrs,
                     const int64_t* numels, kdip_unet** out);
void kdip_unet_destroy(kdip_unet* u);
/* Bytes of caller workspace needed for a batch of N images (activations kept for the VJP included). */
int kdip_unet_workspace_bytes(kdip_unet* u, int N, size_t* bytes);
/* UNetModel.forward (unet.py:636-668): x [N,3,S,S] fp32 NCHW, multiplied on load by x_scale[n] if non-NULL (c_in of
 * k_diffusion/external.py:97-100, condition/condition.py:238); t [N] fp32 timesteps (integers for v1, fractional for
 * v2); out [N,6,S,S] fp32.  cov_out [N,6,S,S] (or NULL) = out_cov(feature) (external.py:163-165).
 * Activations needed by kdip_unet_vjp stay in `workspace` until the next forward on it. */
int kdip_unet_forward(kdip_unet* u, const float* x, const float* x_scale, const float* t, int N, float* out,
                      float* cov_out, void* workspace, size_t ws_bytes, kdip_stream_t s);
/* Input-VJP of the preceding forward: grad_x [N,3,S,S] = d<seed, out>/d(x*x_scale)  (torch.autograd.grad sites
 * condition/condition.py:136,146,155,172,269).  seed [N,6,S,S] fp32. */
int kdip_unet_vjp(kdip_unet* u, const float* seed, int N, float* grad_x, void* workspace, size_t ws_bytes,
                  kdip_stream_t s);
/* Make (N, workspace) the handle's current launch plan WITHOUT launching anything (host-side only).  A caller that replays
 * a captured CUDA graph of kdip_unet_forward (the library is "safe under CUDA-graph capture", SURVEY.md 8(b) threading row)
 * bypasses the library, so the handle would still remember the batch of its last direct call; call this before such a replay
 * so that a following kdip_unet_vjp / kdip_unet_feature sees the batch whose activations really are in `workspace`. */
int kdip_unet_prepare(kdip_unet* u, int N, void* workspace, size_t ws_bytes);

/* Instrumented forward + input-VJP: the same launch lists with a CUDA-event pair around every step, summed by class.
 * Used by bench.py for the live roofline of the dominant kernel (conv_gemm_kernel); synchronises the stream.  One CUDA event is
 * recorded between consecutive launches (a step runs from the event before it to the event after it). */
typedef struct {
  float conv_ms;        /* device time inside tcgen05 implicit-GEMM conv launches                        */
  float other_ms;       /* GroupNorm / attention / small-conv / embedding kernels                        */
  float total_ms;       /* sum over the steps (= first launch to last, minus the two untimed flop-counting passes) */
  double conv_flops;    /* algorithmic FLOPs of those conv launches (2 x MAC, padded channels excluded)  */
  int conv_launches;
  int other_steps;
} kdip_unet_profile_t;
int kdip_unet_profile(kdip_unet* u, const float* x, const float* x_scale, const float* t, const float* seed, int N, float* out,
                      float* grad_x, void* workspace, size_t ws_bytes, kdip_stream_t s, kdip_unet_profile_t* prof);

/* State-dict schema of UNetModel for an architecture (names / shapes of unet.py:463-618), so the host-side module can
 * declare its parameters without restating the block walk.  shape_out: int64[4]. */
int kdip_unet_schema_count(const kdip_unet_arch* arch, int* n);
int kdip_unet_schema_entry(const kdip_unet_arch* arch, int index, char* name_out, int name_cap, int64_t* shape_out,
                           int* ndim);
/* Pre-head feature [N, C0, S, S] fp32 of the last forward (UNetModel.forward(return_feature=True), unet.py:665-666). */
int kdip_unet_feature(kdip_unet* u, int N, float* feat, kdip_stream_t s);

/* ------------------------------------------------------------------------------------------------------------
 * Fused guided model evaluation - the fast path of SURVEY.md 8(b) "Denoiser API": ConditionDenoiser.forward
 * (condition/condition.py:83-174) over ConditionOpenAIDenoiser.uncond_pred (:231-274) for every branch with a closed-form
 * mat solver: hat_x0 = clip(x0 + coef * J^T v, -1, 1).  One call enqueues UNet forward, p_mean_variance epilogue, mat solver
 * (or the DPS residual gradient), VJP seed, UNet input-VJP and the combine on `s` out of the caller's workspace: no allocation,
 * no synchronisation.  kdip_guided_eval = kdip_guided_eval_set (the scalars of `cfg` travel BY VALUE as kernel arguments into
 * per-image device arrays of the workspace; the struct is not read after the call returns) + kdip_guided_eval_run (device-resident
 * data only: capture it ONCE as a CUDA graph per (guidance, B), then per evaluation: _set eagerly, replay).  sigma is uniform over
 * the batch, as inside a sampler call.  The per-pixel covariance + CG branch (sigma < mle_sigma_thres with Convert / TMPD / DWT-Var)
 * is NOT covered: it polls convergence on the host - use kdip_unet_* + kdip_mat_cg + kdip_guidance_combine.
 * ------------------------------------------------------------------------------------------------------------ */
#define KDIP_GUIDE_UNCOND 0  /* hat_x0 = x0_mean                                                  condition.py:104-106 */
#define KDIP_GUIDE_TYPE_I 1  /* x0 + sigma^2 J^T mat(theta), scalar theta (Convert above thres, Analytic, ...)   :167-174 */
#define KDIP_GUIDE_PGDM 2    /* theta = r^2: x0 + sigma^2 r^2 J^T mat                                             :150-157 */
#define KDIP_GUIDE_DPS 3     /* x0 + sigma^2 zeta J^T A^T r / ||r||                                               :140-148 */
#define KDIP_GUIDE_DIFFPIR 4 /* x0 + theta mat(theta), theta = sigma^2 / lambda (no VJP)                          :159-165 */
typedef struct {
  int guidance;         /* KDIP_GUIDE_*                                                                                   */
  float sigma;          /* noise level of this evaluation                                                                 */
  float t_model;        /* timestep fed to the UNet: timestep_map[floor(sigma_to_t(sigma))] (condition.py:233, respace.py:123-128) */
  float theta;          /* scalar x0 variance handed to the mat solver (r^2, the Analytic table entry, sigma^2/lambda)    */
  float zeta;           /* DPS step size                                                                                  */
  kdip_pmv_scalars sc;  /* schedule constants at that integer timestep (c_in included)                                    */
} kdip_guided_cfg;
int kdip_guided_eval_workspace_bytes(kdip_unet* u, const kdip_op* op, int B, size_t* bytes);
/* x [B,3,S,S] (unscaled x_t), y: the measurement (operator.forward's shape) -> hat_x0 [B,3,S,S]. */
int kdip_guided_eval(kdip_unet* u, kdip_op* op, const kdip_guided_cfg* cfg, const float* x, const float* y, float* hat_x0, int B,
                     void* ws, size_t ws_bytes, kdip_stream_t s);
int kdip_guided_eval_set(kdip_unet* u, const kdip_op* op, const kdip_guided_cfg* cfg, int B, void* ws, size_t ws_bytes, kdip_stream_t s);
int kdip_guided_eval_run(kdip_unet* u, kdip_op* op, int guidance, const float* x, const float* y, float* hat_x0, int B, void* ws,
                         size_t ws_bytes, kdip_stream_t s);

/* ------------------------------------------------------------------------------------------------------------
 * LPIPS building blocks - sample_condition_openai.py:46,161 (`lpips.LPIPS(net='vgg')`; the `lpips` package v0.1 is a third-party
 * dependency, its published algorithm is restated in csrc/lpips.cu and oracle/lpips_ref.py).  The VGG16 convolutions go through
 * kdip_conv_plan_* / kdip_layer_conv_small_cin; these are the kernels around them.  Activations are bf16 NHWC.
 * ------------------------------------------------------------------------------------------------------------ */
/* in-place ReLU over n bf16 elements (n a multiple of 8, 16-byte aligned) */
int kdip_relu_bf16(void* x, size_t n, kdip_stream_t s);
/* nn.MaxPool2d(2, 2): [N,H,W,C] -> [N,H/2,W/2,C] (even H, W; C a multiple of 8) */
int kdip_maxpool2_bf16(const void* in, void* out, int N, int H, int W, int C, kdip_stream_t s);
/* out[n] += mean_p sum_c w[c] (f0[n,p,c] / (|f0[n,p,:]| + 1e-10) - f1[n,p,c] / (|f1[n,p,:]| + 1e-10))^2   (out: fp64 [N], caller zeroes;
 * lpips normalize_tensor + (diff)^2 + NetLinLayer 1x1 conv + spatial_average of one feature tap) */
int kdip_lpips_layer(const void* f0, const void* f1, const float* w, int N, int HW, int C, double* out, kdip_stream_t s);

/* ------------------------------------------------------------------------------------------------------------
 * UNet building blocks, exported for per-layer parity tests (tests/test_layers_gpu.py).  Activations are bf16 NHWC.
 * See csrc/unet_kernels.cuh for the full semantics of each argument.
 * ------------------------------------------------------------------------------------------------------------ */
int kdip_layer_chan_stats(const void* x, int N, int P, int C, float* stats, kdip_stream_t s);
int kdip_layer_gn_finalize(const float* stats0, int C0, const float* stats1, int C1, int N, int P, const float* gamma,
                           const float* beta, const float* film, int film_stride, int film_off, float* ab, float* mr,
                           kdip_stream_t s);
int kdip_layer_gn_apply(const void* src0, int C0, const void* src1, int C1, int N, int H, int W, const float* ab,
                        int act_silu, int resample, void* out, kdip_stream_t s);
int kdip_layer_gn_bwd(const void* src0, int C0, const void* src1, int C1, int N, int H, int W, const float* ab,
                      const float* mr, int act_silu, int resample, const void* gy, const void* extra, int extra_mode,
                      float* red_zeroed, float* k_scratch, void* dst0, void* dst1, kdip_stream_t s);
int kdip_layer_conv_small_cin(const float* in, const float* in_scale, const float* w_oihw, const float* bias, int N,
                              int O, int I, int flip, int H, int W, float* w_scratch, void* out, kdip_stream_t s);
int kdip_layer_time_embed(const float* t, int N, int mc, const float* w1, const float* b1, const float* w2,
                          const float* b2, float* semb, kdip_stream_t s);
int kdip_layer_emb_proj(const float* semb, int N, int ted, const float* wall, const float* ball, int R, float* out,
                        kdip_stream_t s);
int kdip_layer_attention_fwd(const void* qkv, int N, int T, int heads, void* out, float* lse, kdip_stream_t s);
int kdip_layer_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, int N, int T,
                             int heads, void* dqkv, kdip_stream_t s);
/* fp32 NCHW <-> bf16 NHWC converters (test plumbing and v2 feature export) */
int kdip_nchw_f32_to_nhwc_bf16(const float* src, int N, int C, int H, int W, void* dst, kdip_stream_t s);
int kdip_nhwc_bf16_to_nchw_f32(const void* src, int N, int C, int H, int W, float* dst, kdip_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* KDIP_H_ */
