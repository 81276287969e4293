"""DDPM schedule constants and p_mean_variance (guided_diffusion/gaussian_diffusion.py:18-35,118-169,232-333,895-908).

Training losses, ancestral / DDIM sampling loops (gaussian_diffusion.py:395-893) are outside the guided-sampling path
and are not provided.  The elementwise epilogue (eps -> clamped x0, learned-range variance) runs in the
``kdip_pmv_epilogue`` CUDA kernel; its backward (clamp mask, eps / direct terms) in ``kdip_pmv_vjp_seed``.
"""
import enum
import math

import numpy as np
import torch as th

from kdip import ops


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """gaussian_diffusion.py:18-44."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        n = num_diffusion_timesteps
        return np.array([min(1 - f((i + 1) / n) / f(i / n), 0.999) for i in range(n)])
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """gaussian_diffusion.py:895-908: numpy float64 table -> fp32 tensor broadcast to ``broadcast_shape``."""
    res = th.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


class _PmvEpilogue(th.autograd.Function):
    """(model_output [B,6,H,W], x [B,3,H,W]) -> (pred_xstart, variance); backward through the clamp."""

    @staticmethod
    def forward(ctx, model_output, x, sc, want_var):
        x0, var = ops.pmv_epilogue(model_output.contiguous().float(), x.contiguous().float(), sc,
                                   ops.VAR_MODEL if want_var else 0)
        ctx.sc = sc
        ctx.save_for_backward(x0)
        if var is None:
            var = x0.new_empty(0)
        ctx.mark_non_differentiable(var)
        return x0, var

    @staticmethod
    def backward(ctx, g_x0, g_var):
        (x0,) = ctx.saved_tensors
        seed, direct = ops.pmv_vjp_seed(x0, g_x0.contiguous(), ctx.sc)
        return seed, direct, None, None


class _Lazy(dict):
    """p_mean_variance result: 'pred_xstart' and 'variance' come from the fused kernel; the posterior 'mean' and
    'log_variance' (gaussian_diffusion.py:300-311), unused by the guidance path, are built on first access."""

    def __init__(self, diffusion, x, t, **kw):
        super().__init__(**kw)
        self._d, self._x, self._t = diffusion, x, t

    def __missing__(self, key):
        if key == "log_variance":
            v = th.log(self["variance"])
        elif key == "mean":
            d, x, t = self._d, self._x, self._t
            v = (_extract_into_tensor(d.posterior_mean_coef1, t, x.shape) * self["pred_xstart"]
                 + _extract_into_tensor(d.posterior_mean_coef2, t, x.shape) * x)
        else:
            raise KeyError(key)
        self[key] = v
        return v


class GaussianDiffusion:
    """Schedule constants (float64 numpy, gaussian_diffusion.py:133-169) + p_mean_variance."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def _predict_xstart_from_eps(self, x_t, t, eps):
        """gaussian_diffusion.py:328-333 (plain torch; the sampling path uses the fused kernel instead)."""
        return (_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def pmv_from_output(self, model_output, x, t, want_var=True, c_in=None, t_host=None):
        """Epilogue of p_mean_variance on an already computed UNet output.  ``x`` is the UNet input (x_t * c_in) when
        ``c_in`` is None, or the unscaled x_t with per-image ``c_in`` (list of floats) applied inside the kernel."""
        B = x.shape[0]
        if t_host is None:
            t_host = [int(v) for v in t.tolist()]
        sc = ops.pmv_scalars(self, t_host, c_in if c_in is not None else [1.0] * B, x.device)
        x0, var = _PmvEpilogue.apply(model_output, x, sc, want_var)
        return x0, (var if want_var else None), sc

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """gaussian_diffusion.py:232-326 for ModelMeanType.EPSILON + ModelVarType.LEARNED_RANGE with clip_denoised
        (the configuration of script_util.py:410-421 with learn_sigma=True)."""
        if model_kwargs is None:
            model_kwargs = {}
        if self.model_mean_type != ModelMeanType.EPSILON or self.model_var_type != ModelVarType.LEARNED_RANGE:
            raise NotImplementedError("kdip p_mean_variance covers the EPSILON / LEARNED_RANGE configuration of the guided-sampling path")
        if not clip_denoised or denoised_fn is not None:
            raise NotImplementedError("kdip p_mean_variance always clamps pred_xstart to [-1, 1] (clip_denoised=True, no denoised_fn)")
        B, C = x.shape[:2]
        assert t.shape == (B,)
        model_output = model(x, self._scale_timesteps(t), **model_kwargs)
        assert model_output.shape == (B, C * 2, *x.shape[2:])
        x0, var, _ = self.pmv_from_output(model_output, x, t, want_var=True)
        return _Lazy(self, x, t, variance=var, pred_xstart=x0)
