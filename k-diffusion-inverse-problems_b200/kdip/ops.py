"""Thin Python handles over the libkdip C ABI for the guidance / operator / sampler kernels.

Every function here launches hand-written CUDA through ctypes on torch's current stream; torch is used only to own
device memory.  There is no CPU fallback: tensors must be CUDA fp32 contiguous (``_lib.ptr`` asserts it).
"""
import ctypes
import warnings

import numpy as np
import torch

from ._lib import KDIP_ENOTCONV, GuidedCfg, OpDesc, PmvScalars, check, lib, ptr, stream_ptr

OP_KIND = {"inpainting": 0, "gaussian_blur": 1, "motion_blur": 2, "super_resolution": 3}
OT_KIND = {None: 0, "dct": 1, "dwt": 2}


def _f32(t):
    return t.contiguous().float()


class Workspace:
    """A growable, 256-byte aligned device scratch buffer (a torch uint8 tensor)."""

    def __init__(self, device):
        self.device = device
        self._buf = None

    def get(self, nbytes):
        if self._buf is None or self._buf.numel() < nbytes + 256:
            self._buf = None
            self._buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        off = (-self._buf.data_ptr()) % 256
        return ctypes.c_void_p(self._buf.data_ptr() + off), self._buf.numel() - 256


class OperatorHandle:
    """Device-resident measurement operator (kdip_op_*): OTF / mask / Resizer tables live in the library."""

    def __init__(self, kind, S, sigma_s, device, psf=None, mask=None, sf=1, resizer=None):
        if not torch.cuda.is_available():
            raise RuntimeError("kdip operators need a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.kind, self.S, self.sf, self.device = kind, int(S), int(sf), torch.device(device)
        d = OpDesc()
        d.kind, d.S, d.sf, d.sigma_s = OP_KIND[kind], int(S), int(sf), float(sigma_s)
        keep = []
        if psf is not None:
            a = np.ascontiguousarray(psf, dtype=np.float32)
            keep.append(a)
            d.psf, d.ksize = a.ctypes.data, a.shape[-1]
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.float32).reshape(3, S, S)
            keep.append(m)
            d.mask = m.ctypes.data
        if resizer is not None:
            w, idx = resizer
            w = np.ascontiguousarray(w, dtype=np.float32)
            idx = np.ascontiguousarray(idx, dtype=np.int32)
            keep += [w, idx]
            d.rs_w, d.rs_idx, d.rs_taps = w.ctypes.data, idx.ctypes.data, w.shape[1]
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.kdip_op_create(ctypes.byref(d), ctypes.byref(h)))
        self._h = h
        self._ws = Workspace(self.device)
        self._ws_bytes = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:     # at interpreter shutdown the module globals may already be gone
            lib.kdip_op_destroy(h)
            self._h = None

    def _workspace(self, B):
        if B not in self._ws_bytes:
            n = ctypes.c_size_t()
            check(lib.kdip_op_workspace_bytes(self._h, B, ctypes.byref(n)))
            self._ws_bytes[B] = n.value
        return self._ws.get(self._ws_bytes[B])

    def out_hw(self):
        return self.S // self.sf

    def forward(self, x, noise=None):
        x = _f32(x)
        B = x.shape[0]
        y = torch.empty(B, 3, self.out_hw(), self.out_hw(), device=x.device, dtype=torch.float32)
        if noise is not None:
            noise = _f32(noise)
        ws, nb = self._workspace(B)
        check(lib.kdip_op_forward(self._h, ptr(x), ptr(noise), ptr(y), B, ws, nb, stream_ptr()))
        return y

    def transpose(self, y):
        y = _f32(y)
        B = y.shape[0]
        x = torch.empty(B, 3, self.S, self.S, device=y.device, dtype=torch.float32)
        ws, nb = self._workspace(B)
        check(lib.kdip_op_transpose(self._h, ptr(y), ptr(x), B, ws, nb, stream_ptr()))
        return x

    def forward_adjoint(self, g):
        """Adjoint of forward(noiseless) applied to g (y's shape) -> [B,3,S,S]."""
        g = _f32(g)
        B = g.shape[0]
        x = torch.empty(B, 3, self.S, self.S, device=g.device, dtype=torch.float32)
        ws, nb = self._workspace(B)
        check(lib.kdip_op_forward_adjoint(self._h, ptr(g), ptr(x), B, ws, nb, stream_ptr()))
        return x

    def fft2(self, x):
        """torch.fft.fftn(x, dim=(-2, -1)) for x [B,3,S,S] on libkdip's FFT kernels -> complex64 [B,3,S,S]."""
        x = _f32(x)
        B = x.shape[0]
        out = torch.empty(B, 3, self.S, self.S, 2, device=x.device, dtype=torch.float32)
        ws, nb = self._workspace(B)
        check(lib.kdip_op_fft2(self._h, ptr(x), ptr(out), B, ws, nb, stream_ptr()))
        return torch.view_as_complex(out)

    def otf(self):
        fb = torch.empty(self.S, self.S, 2, device=self.device, dtype=torch.float32)
        check(lib.kdip_op_otf(self._h, ptr(fb), stream_ptr()))
        return torch.view_as_complex(fb)

    def mat_closed(self, y, x0, theta):
        """theta: [B] device tensor of scalar variances."""
        y, x0, theta = _f32(y), _f32(x0), _f32(theta)
        B = x0.shape[0]
        mat = torch.empty_like(x0)
        ws, nb = self._workspace(B)
        check(lib.kdip_mat_closed(self._h, ptr(y), ptr(x0), ptr(theta), ptr(mat), B, ws, nb, stream_ptr()))
        return mat

    def mat_cg(self, y, x0, theta_map, ot=None, tol=1e-4, maxiter=1000):
        y, x0, theta_map = _f32(y), _f32(x0), _f32(theta_map)
        B = x0.shape[0]
        mat = torch.empty_like(x0)
        iters = (ctypes.c_int * B)()
        ws, nb = self._workspace(B)
        rc = lib.kdip_mat_cg(self._h, ptr(y), ptr(x0), ptr(theta_map), OT_KIND[ot], ptr(mat), B, tol, maxiter, iters, ws, nb,
                             stream_ptr())
        if rc == KDIP_ENOTCONV:
            warnings.warn("CG not converge.")        # condition/condition.py:344-345: non-fatal
        else:
            check(rc)
        self.last_cg_iters = list(iters)
        return mat

    def dps_grad(self, y, x0):
        """-> (v = A^T (y - A x0) [B,3,S,S], norm [B] = ||y - A x0||_2)."""
        y, x0 = _f32(y), _f32(x0)
        B = x0.shape[0]
        v = torch.empty_like(x0)
        norm = torch.empty(B, device=x0.device, dtype=torch.float32)
        ws, nb = self._workspace(B)
        check(lib.kdip_dps_grad(self._h, ptr(y), ptr(x0), ptr(v), ptr(norm), B, ws, nb, stream_ptr()))
        return v, norm


GUIDE = {"uncond": 0, "I": 1, "pgdm": 2, "dps": 3, "diffpir": 4}


class FusedGuidedEval:
    """kdip_guided_eval: one library call per guided model evaluation for the closed-form branches - and, after two eager calls,
    kdip_guided_eval_set (a one-CTA kernel that takes the evaluation's scalars by value) + ONE CUDA-graph replay of
    kdip_guided_eval_run, the same graph for every sigma of the schedule."""

    def __init__(self, engine, handle):
        self.engine, self.handle = engine, handle
        self.cfg = GuidedCfg()
        self._graphs = {}
        import os
        self._graphs_on = os.environ.get("KDIP_CUDA_GRAPH", "1") != "0"

    def _ws(self, B):
        n = ctypes.c_size_t()
        check(lib.kdip_guided_eval_workspace_bytes(self.engine._h, self.handle._h, B, ctypes.byref(n)))
        ws, _ = self.engine._workspace(B, at_least=n.value)
        return ws, n.value

    def _run(self, x, y, hat, B, ws, nb):
        check(lib.kdip_guided_eval_run(self.engine._h, self.handle._h, int(self.cfg.guidance), ptr(x), ptr(y), ptr(hat), B, ws, nb,
                                       stream_ptr()))

    def __call__(self, guidance, sigma, t_model, theta, zeta, sc_one, x, y):
        """sc_one: PmvScalars for this sigma (uniform over the batch).  x [B,3,S,S], y: measurement -> hat_x0."""
        c = self.cfg
        c.guidance, c.sigma, c.t_model, c.theta, c.zeta, c.sc = GUIDE[guidance], float(sigma), float(t_model), float(theta), float(zeta), sc_one
        x, y = _f32(x), _f32(y)
        B = x.shape[0]
        ws, nb = self._ws(B)
        hat = torch.empty_like(x)
        check(lib.kdip_guided_eval_set(self.engine._h, self.handle._h, ctypes.byref(c), B, ws, nb, stream_ptr()))
        key = (guidance, B, tuple(y.shape), ws.value)
        r = self._graphs.setdefault(key, {"calls": 0, "graph": None, "failed": False}) if self._graphs_on else None
        if r is not None and not r["failed"]:
            r["calls"] += 1
            if r["calls"] > 2 and r["graph"] is None:
                try:
                    r["x"], r["y"], r["hat"] = torch.zeros_like(x), torch.zeros_like(y), torch.zeros_like(x)
                    cur, side = torch.cuda.current_stream(), torch.cuda.Stream()
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        self._run(r["x"], r["y"], r["hat"], B, ws, nb)       # settles every lazily built plan
                    cur.wait_stream(side)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    n0 = lib.kdip_launch_count()
                    with torch.cuda.graph(g):
                        self._run(r["x"], r["y"], r["hat"], B, ws, nb)
                    r["graph"], r["kernels"] = g, int(lib.kdip_launch_count() - n0)
                except Exception as e:                                          # noqa: BLE001 - eager launches stay functional
                    warnings.warn(f"kdip: CUDA-graph capture of the fused guided evaluation failed ({e}); using eager launches")
                    r["failed"], r["graph"] = True, None
            if r["graph"] is not None:
                r["x"].copy_(x)
                r["y"].copy_(y)
                r["graph"].replay()
                lib.kdip_launch_count_add(r["kernels"])
                hat.copy_(r["hat"])
                self.engine._fwd_N = B
                self.engine.forward_token += 1
                return hat
        self._run(x, y, hat, B, ws, nb)
        self.engine._fwd_N = B
        self.engine.forward_token += 1
        return hat


_ortho_ws = {}


def ortho(ot, x, inverse=False, mul=None):
    """OrthoTransform kernels: out = mul .* W^T x (forward) or W x (inverse).  x [B,3,S,S]."""
    x = _f32(x)
    B, _, S, _ = x.shape
    out = torch.empty_like(x)
    if mul is not None:
        mul = _f32(mul.expand_as(x))
    n = ctypes.c_size_t()
    check(lib.kdip_ortho_workspace_bytes(OT_KIND[ot], B, S, ctypes.byref(n)))
    ws, nb = None, 0
    if n.value:
        w = _ortho_ws.setdefault(x.device, Workspace(x.device))
        ws, nb = w.get(n.value)
    check(lib.kdip_ortho(OT_KIND[ot], int(inverse), ptr(x), ptr(mul), ptr(out), B, S, ws, nb, stream_ptr()))
    return out


# ---- p_mean_variance epilogue / guidance combine -----------------------------------------------------------------------

def pmv_scalars_one(diffusion, t, c_in, dst=None):
    """kdip_pmv_scalars at integer timestep t from the float64 schedule (gaussian_diffusion.py:895-908: float64 -> fp32 at use)."""
    e = PmvScalars() if dst is None else dst
    tb = int(t)
    e.c_in = float(np.float32(c_in))
    e.recip = float(np.float32(diffusion.sqrt_recip_alphas_cumprod[tb]))
    e.recipm1 = float(np.float32(diffusion.sqrt_recipm1_alphas_cumprod[tb]))
    e.min_log = float(np.float32(diffusion.posterior_log_variance_clipped[tb]))
    e.max_log = float(np.float32(np.log(diffusion.betas[tb])))
    e.post_var = float(np.float32(diffusion.posterior_variance[tb]))
    c1 = np.float32(diffusion.posterior_mean_coef1[tb])
    e.coef1_sq = float(c1 * c1)
    return e


def pmv_scalars(diffusion, t, c_in, device):
    """Per-image scalar table for the epilogue kernels.  t: list[int], c_in: list[float]."""
    B = len(t)
    arr = (PmvScalars * B)()
    for b in range(B):
        pmv_scalars_one(diffusion, t[b], c_in[b], arr[b])
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device, non_blocking=False)


VAR_MODEL, VAR_CONVERT = 1, 2


def pmv_epilogue(unet_out, x, sc, var_mode=0):
    """var_mode: 0 no variance, VAR_MODEL = p_mean_variance's variance, VAR_CONVERT = Eq. (22) x0 variance."""
    B, _, H, W = x.shape
    x0 = torch.empty_like(x)
    var = torch.empty_like(x) if var_mode else None
    check(lib.kdip_pmv_epilogue(ptr(unet_out), ptr(x), ptr(sc), ptr(x0), ptr(var), int(var_mode), B, H * W, stream_ptr()))
    return x0, var


def v2_vjp_scalars(sigma_host, device):
    """kdip_pmv_scalars for the unclamped v2 denoiser x0 = x - sigma*eps: seed = (-sigma v, 0), direct = v."""
    B = len(sigma_host)
    arr = (PmvScalars * B)()
    for b in range(B):
        arr[b].c_in, arr[b].recip, arr[b].recipm1 = 1.0, 1.0, float(np.float32(sigma_host[b]))
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)


def pmv_vjp_seed(x0_mean, v, sc, like=None):
    """x0_mean None: no clamp mask (v2)."""
    B, _, H, W = (x0_mean if x0_mean is not None else like).shape
    seed = torch.empty(B, 6, H, W, device=v.device, dtype=torch.float32)
    direct = torch.empty(B, 3, H, W, device=v.device, dtype=torch.float32)
    v = _f32(v)
    check(lib.kdip_pmv_vjp_seed(ptr(x0_mean), ptr(v), ptr(sc), ptr(seed), ptr(direct), B, H * W, stream_ptr()))
    return seed, direct


def guidance_combine(x0_mean, g, direct, coef, c_in=None, out=None):
    B = x0_mean.shape[0]
    chw = x0_mean[0].numel()
    hat = torch.empty_like(x0_mean) if out is None else out
    g, coef = _f32(g), _f32(coef)          # bound to locals: a converted temporary must outlive the launch
    check(lib.kdip_guidance_combine(ptr(x0_mean), ptr(g), ptr(direct), ptr(coef), ptr(c_in), ptr(hat), B, chw,
                                    stream_ptr()))
    return hat


def lincomb(x, y, a, c):
    """out = a[b]*x + c[b]*y (per-image device scalars), unclipped."""
    B = x.shape[0]
    x, a = _f32(x), _f32(a)
    out = torch.empty_like(x)
    check(lib.kdip_lincomb(ptr(x), ptr(y), ptr(a), ptr(c), ptr(out), B, x[0].numel(), stream_ptr()))
    return out


def v2_epilogue(unet_out, cov_out, x, sigma_dev, want_var):
    """(x0_mean, x0_var, theta0_var) of ConditionOpenAIDenoiserV2.uncond_pred; variances None unless want_var."""
    B, _, H, W = x.shape
    x0 = torch.empty_like(x)
    var = torch.empty_like(x) if want_var else None
    var_ot = torch.empty_like(x) if want_var else None
    check(lib.kdip_v2_epilogue(ptr(unet_out), ptr(cov_out), ptr(x), ptr(sigma_dev), ptr(x0), ptr(var), ptr(var_ot), B, H * W,
                               stream_ptr()))
    return x0, var, var_ot


# ---- sampler updates -------------------------------------------------------------------------------------------------

def churn_(x, noise, s_noise, sigma, sigma_hat):
    noise = _f32(noise)
    check(lib.kdip_churn(ptr(x), ptr(noise), float(s_noise), float(sigma), float(sigma_hat), x.numel(), stream_ptr()))
    return x


def euler_step(x, denoised, sigma_hat, dt, want_d=False):
    x_out = torch.empty_like(x)
    d = torch.empty_like(x) if want_d else None
    denoised = _f32(denoised)
    check(lib.kdip_euler_step(ptr(x), ptr(denoised), float(sigma_hat), float(dt), ptr(x_out), ptr(d), x.numel(), stream_ptr()))
    return (x_out, d) if want_d else x_out


def heun_step(x, d, x2, denoised2, sigma_next, dt):
    x_out = torch.empty_like(x)
    denoised2 = _f32(denoised2)
    check(lib.kdip_heun_step(ptr(x), ptr(d), ptr(x2), ptr(denoised2), float(sigma_next), float(dt), ptr(x_out), x.numel(),
                             stream_ptr()))
    return x_out


def lincomb3(x, a, y=None, b=0.0, z=None, c=0.0, out=None):
    """a*x + b*y + c*z with host scalars (kdip_lincomb3); y / z optional."""
    x = _f32(x)
    out = torch.empty_like(x) if out is None else out
    y = _f32(y) if y is not None else None
    z = _f32(z) if z is not None else None
    check(lib.kdip_lincomb3(ptr(x), ptr(y), ptr(z), float(a), float(b), float(c), ptr(out), x.numel(), stream_ptr()))
    return out


def gather(src, idx):
    B, M = src.shape[0], idx.numel()
    dst = torch.empty(B, M, device=src.device, dtype=torch.float32)
    src = _f32(src)
    check(lib.kdip_gather(ptr(src), ptr(idx), ptr(dst), B, src[0].numel(), M, stream_ptr()))
    return dst


def scatter(src, idx, shape):
    B = src.shape[0]
    dst = torch.empty(B, *shape, device=src.device, dtype=torch.float32)
    src = _f32(src)
    check(lib.kdip_scatter(ptr(src), ptr(idx), ptr(dst), B, dst[0].numel(), idx.numel(), stream_ptr()))
    return dst


# ---- data front end / evaluation reductions (csrc/images.cu) ------------------------------------------------------------------

def images_u8_to_f32(u8_hwc):
    """[B,H,W,3] uint8 CUDA (PIL order) -> [B,3,H,W] fp32 in [-1,1]: ToTensor then x*2-1 (sample_condition_openai.py:140-144)."""
    assert u8_hwc.dtype == torch.uint8 and u8_hwc.dim() == 4 and u8_hwc.shape[-1] == 3
    B, H, W, _ = u8_hwc.shape
    out = torch.empty(B, 3, H, W, device=u8_hwc.device, dtype=torch.float32)
    check(lib.kdip_images_u8_to_f32(ptr(u8_hwc), ptr(out), B, H, W, stream_ptr()))
    return out


def images_f32_to_u8(x):
    """[B,3,H,W] fp32 -> [B,H,W,3] uint8 as k_diffusion.utils.to_pil_image would store it (utils.py:24-31)."""
    x = _f32(x)
    B, _, H, W = x.shape
    out = torch.empty(B, H, W, 3, device=x.device, dtype=torch.uint8)
    check(lib.kdip_images_f32_to_u8(ptr(x), ptr(out), B, H, W, stream_ptr()))
    return out


def sqerr_sum(a, b, to_eval_first=False):
    """Per-image sum of squared differences, fp64 [B] (device)."""
    a, b = _f32(a), _f32(b)
    B = a.shape[0]
    out = torch.empty(B, device=a.device, dtype=torch.float64)
    check(lib.kdip_sqerr_sum(ptr(a), ptr(b), int(to_eval_first), ptr(out), B, a[0].numel(), stream_ptr()))
    return out


def denoise_sqerr(unet_out, x_noised, x0, sigma_dev, want_hat=False):
    """analytic_variance.py:128-129: per-image sum (x0 - (x_noised - sigma*eps))^2 as fp64 [B] (and hat_x0 if asked)."""
    unet_out, x_noised, x0 = _f32(unet_out), _f32(x_noised), _f32(x0)
    B, _, H, W = x0.shape
    out = torch.empty(B, device=x0.device, dtype=torch.float64)
    hat = torch.empty_like(x0) if want_hat else None
    sigma_dev = _f32(sigma_dev)
    check(lib.kdip_denoise_sqerr(ptr(unet_out), ptr(x_noised), ptr(x0), ptr(sigma_dev), ptr(out), ptr(hat), B, H * W, stream_ptr()))
    return (out, hat) if want_hat else out


def psnr(x0, hat_x0):
    """peak_signal_noise_ratio(to_eval(x0), to_eval(hat_x0), data_range=1) per image, fp64 [B] (sample_condition_openai.py:41-44)."""
    n = x0[0].numel()
    return 10.0 * torch.log10(n / sqerr_sum(x0, hat_x0, to_eval_first=True))


def ssim(x0, hat_x0):
    """structural_similarity(to_eval(x0), to_eval(hat_x0), channel_axis=0, data_range=1) per image, fp64 [B] (:45)."""
    a, b = _f32(x0), _f32(hat_x0)
    B, C, H, W = a.shape
    assert C == 3
    out = torch.empty(B, device=a.device, dtype=torch.float64)
    check(lib.kdip_ssim_sum(ptr(a), ptr(b), ptr(out), B, H, W, stream_ptr()))
    return out / (3.0 * (H - 6) * (W - 6))
