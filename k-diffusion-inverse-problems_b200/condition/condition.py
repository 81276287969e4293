"""ConditionDenoiser — E[x0 | x_t, y] under a Gaussian posterior with optimal covariance (condition/condition.py:41-439).

Same classes, constructor arguments, guidance names and mat-solver registry as the reference; batch-capable (the
reference asserts B == 1, condition.py:84; here every image of the batch carries its own measurement y and all
per-image scalars broadcast as [B]).  One model evaluation is:

    UNet forward (libkdip, bf16 tcgen05)  ->  p_mean_variance epilogue + x0 covariance (one fused kernel)
    ->  mat = A^T (sigma_s^2 I + A Sigma A^T)^-1 (y - A x0)   (closed form: 3-6 FFT kernels; per-pixel Sigma: on-device batched CG)
    ->  likelihood score = J^T mat through the clamp and the UNet (hand-written input-VJP)
    ->  hat_x0 = clip(x0 + sigma^2 * score, -1, 1)            (fused combine kernel)

sigma is uniform across the batch inside a sampler call; the sigma-dependent branches (MLE threshold, scalar vs.
per-pixel covariance) are decided on the host from the schedule value the sampler attaches to the sigma tensor, so a
model evaluation has no device->host synchronisation except the CG convergence poll.

'stsl' / 'stsl+mle' (condition.py:185-208) is composed from the same pieces: 1 + num_hutchinson_samples UNet forward + VJP pairs.
Out of scope: 'autoI' (a gpytorch Gaussian log-likelihood).
"""
from abc import abstractmethod
from warnings import warn  # noqa: F401  (mat solvers warn on CG non-convergence through kdip.ops)

import numpy as np
import torch
from torch import nn

from guided_diffusion.gaussian_diffusion import GaussianDiffusion, _extract_into_tensor  # noqa: F401
from guided_diffusion.unet import UNetModel
from k_diffusion.external import OpenAIDenoiser, OpenAIDenoiserV2
from kdip import ops

from .utils import OrthoTransform


def _host_sigma(sigma, B):
    """Per-image sigma as python floats.  The kdip samplers attach the host value; otherwise read it back (one sync)."""
    h = getattr(sigma, "_kdip_host", None)
    if h is not None:
        return [float(h)] * B
    v = [float(s) for s in sigma.detach().flatten().tolist()]
    return v * B if len(v) == 1 else v


def _uniform(sig):
    if any(s != sig[0] for s in sig):
        raise ValueError("kdip ConditionDenoiser: sigma must be identical across the batch (it is inside a sampler call)")
    return sig[0]


def _dev(vals, device):
    return torch.tensor(vals, device=device, dtype=torch.float32)


class ConditionDenoiser(nn.Module):
    '''Approximate E[x0|xt, y] given variational Gaussian posterior'''

    def __init__(self, operator, measurement, guidance, device='cpu', zeta=None, lambda_=None, eta=None,
                 num_hutchinson_samples=None, mle_sigma_thres=0.2, ortho_tf_type=None):
        super().__init__()
        self.operator = operator
        self.y, self.y_flatten = measurement
        self.guidance = guidance
        self.zeta = zeta
        self.lambda_ = lambda_
        self.eta = eta
        self.num_hutchinson_samples = num_hutchinson_samples
        self.mle_sigma_thres = mle_sigma_thres
        self.device = device
        self.ortho_tf_type = ortho_tf_type
        self.ortho_tf = OrthoTransform(ortho_tf_type)
        self.mat_solver = __MAT_SOLVER__[operator.name]
        self._ctx = None

    @abstractmethod
    def uncond_pred(self, x, sigma):
        raise NotImplementedError

    # ---- pieces the guidance implementations share --------------------------------------------------------------
    def _y_for(self, B):
        y = self.y
        if y.shape[0] != B:
            if y.shape[0] != 1:
                raise ValueError(f"measurement batch {y.shape[0]} does not match x batch {B}")
            y = y.expand(B, *y.shape[1:]).contiguous()
        return y

    def _score(self, x0_mean, v):
        """J^T v for J = d x0_mean / d x_t: clamp mask and scalings (kdip_pmv_vjp_seed) then the UNet input-VJP.
        Returns the two terms (g wrt the scaled UNet input, direct) that kdip_guidance_combine sums."""
        raise NotImplementedError

    def _data_grad(self, y, x0_mean):
        """(A^T r [B,3,H,W], ||r||_2 [B]) with r = y - operator.forward(x0, noiseless=True) (condition.py:144-146).  Built-in
        operators run kdip_dps_grad; an operator registered by the user without a kdip handle is differentiated through its own
        torch ``forward``, exactly as the reference's autograd does."""
        handle = getattr(self.operator, "handle", None)
        if handle is not None:
            return handle.dps_grad(y, x0_mean)
        with torch.enable_grad():
            x0 = x0_mean.detach().clone().requires_grad_()
            r = y - self.operator.forward(x0, noiseless=True)
            half_sq = 0.5 * r.pow(2).flatten(1).sum(1)
            (g,) = torch.autograd.grad(half_sq.sum(), x0)
        return (-g).contiguous().float(), (2.0 * half_sq.detach()).sqrt().float().contiguous()

    def _theta(self, x0_var, theta0_var):
        return x0_var if self.ortho_tf_type is None else theta0_var

    def forward(self, x, sigma):
        B = x.shape[0]
        sig = _uniform(_host_sigma(sigma, B))
        g = self.guidance
        if g in ("dps+mle", "pgdm+mle", "stsl+mle"):
            g = "I" if sig < self.mle_sigma_thres else g.split("+")[0]
        x = x.detach().contiguous().float()
        fused = self._fused_eval(g, x, sig)
        if fused is not None:
            return fused
        if g == "uncond":
            x0_mean = self.uncond_pred(x, sigma)[0]
            return ops.guidance_combine(x0_mean, x0_mean, None, _dev([0.0] * B, x.device))
        if g == "I":
            return self._type_I_guidance_impl(x, sigma)
        if g == "II":
            return self._type_II_guidance_impl(x, sigma)
        if g == "dps":
            return self._dps_guidance_impl(x, sigma)
        if g == "pgdm":
            return self._pgdm_guidance_impl(x, sigma)
        if g == "diffpir":
            return self._diffpir_guidance_impl(x, sigma)
        if g == "stsl":
            return self._stsl_guidance_impl(x, sigma)
        if g == "autoI":
            raise NotImplementedError(f"guidance '{self.guidance}' is outside the kdip hot path (SURVEY.md §2 #2)")
        raise ValueError(f"Invalid guidance type: '{self.guidance}'.")

    def _fused_eval(self, g, x, sig):
        """One kdip_guided_eval call for the branches it covers (subclasses decide); None -> the composed path below."""
        return None

    def _dps_guidance_impl(self, x, sigma):
        """condition.py:140-148: x0 - sigma^2 zeta grad_x ||y - A(x0)||_2  (per-image norm; equals the reference at B = 1)."""
        assert self.zeta is not None, "zeta must be specified for DPS guidance"
        B = x.shape[0]
        sig = _uniform(_host_sigma(sigma, B))
        x0_mean = self.uncond_pred(x, sigma)[0]
        v, norm = self._data_grad(self._y_for(B), x0_mean)
        g, direct = self._score(x0_mean, v)
        coef = float(np.float32(sig) ** 2 * np.float32(self.zeta)) / norm   # [B] device scalars: sigma^2 zeta / ||r_b||
        return ops.guidance_combine(x0_mean, g, direct, coef, self._ctx["c_in_dev"])

    def _hutchinson_eps(self, x):
        """One probe of the Hutchinson trace estimate (condition.py:198: torch.randn_like(x)); a hook so tests can inject the
        reference's CPU draws."""
        return torch.randn_like(x)

    def _stsl_guidance_impl(self, x, sigma):
        """condition.py:185-208.  loss = -zeta ||y - A x0(x)|| - (eta sigma^2 / (numel n)) sum_k <x0(x + eps_k) - x0(x), eps_k>, and
        hat_x0 = x0 + sigma^2 grad_x loss.  With J(.) = d x0 / d x the gradient is
            J(x)^T [ zeta A^T r / ||r||  +  w sum_k eps_k ]  -  w sum_k J(x + eps_k)^T eps_k,      w = eta sigma^2 / (numel n),
        i.e. one input-VJP at x (seeded with the data term plus the summed probes) and one forward + VJP per probe; numel is
        per image (the reference runs B = 1).  The probes are drawn first, in the reference's order."""
        assert self.zeta is not None and self.eta is not None and self.num_hutchinson_samples is not None, \
            "zeta, eta, and num_hutchinson_samples must be specified for STSL guidance"
        B, n = x.shape[0], int(self.num_hutchinson_samples)
        sig = np.float32(_uniform(_host_sigma(sigma, B)))
        numel = x[0].numel()
        w = float(np.float32(self.eta) / np.float32(numel) * sig ** 2 / np.float32(n)) if n > 0 else 0.0
        one, w_dev, mw_dev = _dev([1.0] * B, x.device), _dev([w] * B, x.device), _dev([-w] * B, x.device)
        x0_mean = self.uncond_pred(x, sigma)[0]
        c_in_dev = self._ctx["c_in_dev"]
        v, norm = self._data_grad(self._y_for(B), x0_mean)                     # A^T r, ||r|| per image
        eps = [self._hutchinson_eps(x).to(x.device, torch.float32).contiguous() for _ in range(n)]
        seed_vec = ops.lincomb(v, eps[0] if n else v, float(np.float32(self.zeta)) / norm, w_dev if n else _dev([0.0] * B, x.device))
        for k in range(1, n):
            seed_vec = ops.lincomb(seed_vec, eps[k], one, w_dev)
        g, direct = self._score(x0_mean, seed_vec)
        for k in range(n):
            x0_k = self.uncond_pred(ops.lincomb(x, eps[k], one, one), sigma)[0]
            g_k, direct_k = self._score(x0_k, eps[k])
            g, direct = ops.lincomb(g, g_k, one, mw_dev), ops.lincomb(direct, direct_k, one, mw_dev)
        return ops.guidance_combine(x0_mean, g, direct, _dev([float(sig ** 2)] * B, x.device), c_in_dev)

    def _pgdm_guidance_impl(self, x, sigma):
        """condition.py:150-157: theta = r^2 = sigma^2/(1+sigma^2); x0 + sigma^2 r^2 J^T mat."""
        B = x.shape[0]
        sig = np.float32(_uniform(_host_sigma(sigma, B)))
        x0_mean = self.uncond_pred(x, sigma)[0]
        r2 = sig ** 2 / (1 + sig ** 2)
        x0_var = _dev([float(r2)] * B, x.device)
        mat = self.mat_solver(self.operator, self._y_for(B), x0_mean, x0_var)
        g, direct = self._score(x0_mean, mat)
        return ops.guidance_combine(x0_mean, g, direct, _dev([float(sig ** 2 * r2)] * B, x.device), self._ctx["c_in_dev"])

    def _diffpir_guidance_impl(self, x, sigma):
        """condition.py:159-165: x0 + theta * mat with theta = sigma^2 / lambda (no VJP)."""
        assert self.lambda_ is not None, "lambda_ must be specified for DiffPIR guidance"
        B = x.shape[0]
        sig = np.float32(_uniform(_host_sigma(sigma, B)))
        x0_mean = self.uncond_pred(x, sigma)[0]
        x0_var = _dev([float(sig ** 2 / np.float32(self.lambda_))] * B, x.device)
        mat = self.mat_solver(self.operator, self._y_for(B), x0_mean, x0_var)
        return ops.guidance_combine(x0_mean, mat, None, x0_var)

    def _type_I_guidance_impl(self, x, sigma):
        """condition.py:167-174: x0 + sigma^2 J^T mat."""
        B = x.shape[0]
        sig = np.float32(_uniform(_host_sigma(sigma, B)))
        x0_mean, x0_var, theta0_var = self.uncond_pred(x, sigma)
        mat = self.mat_solver(self.operator, self._y_for(B), x0_mean, self._theta(x0_var, theta0_var), self.ortho_tf)
        g, direct = self._score(x0_mean, mat)
        return ops.guidance_combine(x0_mean, g, direct, _dev([float(sig ** 2)] * B, x.device), self._ctx["c_in_dev"])

    def _type_II_guidance_impl(self, x, sigma):
        """condition.py:176-183: x0 + W(theta .* W^T mat) (no VJP)."""
        B = x.shape[0]
        x0_mean, x0_var, theta0_var = self.uncond_pred(x, sigma)
        theta = self._theta(x0_var, theta0_var)
        mat = self.mat_solver(self.operator, self._y_for(B), x0_mean, theta, self.ortho_tf)
        if theta.dim() <= 1:                                   # scalar variance: W(theta W^T mat) = theta * mat
            return ops.guidance_combine(x0_mean, mat, None, theta.reshape(-1).expand(B).contiguous())
        upd = ops.ortho(self.ortho_tf_type, ops.ortho(self.ortho_tf_type, mat, mul=theta), inverse=True)
        return ops.guidance_combine(x0_mean, upd, None, _dev([1.0] * B, x.device))


class ConditionOpenAIDenoiser(ConditionDenoiser):
    """condition.py:211-274 (v1: OpenAI UNet + GaussianDiffusion, integer t, clamped x0)."""

    def __init__(self, inner_model, diffusion: GaussianDiffusion, x0_cov_type, recon_mse, **kwargs):
        super().__init__(**kwargs)
        if not isinstance(inner_model, UNetModel):
            raise TypeError("kdip ConditionOpenAIDenoiser needs the kdip guided_diffusion.unet.UNetModel as inner_model")
        self.inner_model = inner_model
        self.diffusion = diffusion
        self.denoiser = OpenAIDenoiser(inner_model, diffusion, device=self.device)
        self.x0_cov_type = x0_cov_type
        self.recon_mse = recon_mse
        if recon_mse is not None:
            for key in self.recon_mse.keys():
                self.recon_mse[key] = self.recon_mse[key].to(self.device)
            self._recon_sigmas_host = self.recon_mse['sigmas'].detach().float().cpu().numpy()
            self._recon_mse_host = self.recon_mse['mse_list'].detach().float().cpu().numpy()

    def _timestep(self, sig):
        """(t_int, t_model) for a host sigma: condition.py:233 (.long() truncates), respace.py:123-128."""
        t_int = int(self.denoiser.sigma_to_t_host(float(sig)))
        tm = getattr(self.diffusion, "timestep_map", None)
        t_model = tm[t_int] if tm is not None else t_int
        if getattr(self.diffusion, "rescale_timesteps", False):
            orig = getattr(self.diffusion, "original_num_steps", self.diffusion.num_timesteps)
            t_model = float(t_model) * (1000.0 / orig)
        return t_int, t_model

    def _fused_eval(self, g, x, sig):
        """kdip_guided_eval (include/kdip.h): the whole evaluation as one library call / one CUDA-graph replay whenever the mat
        solver is a built-in closed form: guidance uncond / pgdm / dps / diffpir, and type I with a scalar x0 variance (Convert
        above mle_sigma_thres, Analytic, pgdm, dps, diffpir covariance types).  KDIP_FUSED_EVAL=0 selects the composed path."""
        import os
        if os.environ.get("KDIP_FUSED_EVAL", "1") == "0" or self.ortho_tf_type is not None:
            return None
        handle = getattr(self.operator, "handle", None)
        if handle is None or self.mat_solver is not _BUILTIN_SOLVERS.get(getattr(self.operator, "name", None)):
            return None
        sig = np.float32(sig)
        r2 = float(sig ** 2 / (1 + sig ** 2))
        mle = bool(sig < self.mle_sigma_thres)
        theta, zeta = 0.0, 0.0
        if g == "pgdm":
            theta = r2
        elif g == "diffpir":
            if self.lambda_ is None:
                return None
            theta = float(sig ** 2 / np.float32(self.lambda_))
        elif g == "dps":
            if self.zeta is None:
                return None
            zeta = float(np.float32(self.zeta))
        elif g == "I":
            ct = self.x0_cov_type
            if ct == "convert" and not mle:
                theta = r2
            elif ct == "analytic" and self.recon_mse is not None:
                theta = float(self._recon_mse_host.reshape(-1)[int(np.abs(self._recon_sigmas_host - sig).argmin())]) if mle else r2
            elif ct == "pgdm":
                theta = r2
            elif ct == "dps":
                theta = 0.0
            elif ct == "diffpir" and self.lambda_ is not None:
                theta = float(sig ** 2 / np.float32(self.lambda_))
            else:
                return None                       # per-pixel covariance (Convert below the threshold, TMPD): CG path
        elif g != "uncond":
            return None
        eng = self.inner_model.engine()
        # one FusedGuidedEval (its captured CUDA graphs and static buffers) per (operator handle, engine), shared by every denoiser
        # built on them: the sample scripts construct a new ConditionOpenAIDenoiser per measurement, and re-capturing the graph
        # (two eager evaluations + a settling run) for each of them cost ~1 % of a 199-evaluation trajectory
        cache = handle.__dict__.setdefault("_fused_cache", {})
        fe = cache.get(id(eng))
        if fe is None or fe.engine is not eng:
            fe = cache[id(eng)] = ops.FusedGuidedEval(eng, handle)
        self._fused = fe
        t_int, t_model = self._timestep(sig)
        c_in = float(np.float32(1) / np.sqrt(sig * sig + np.float32(1)))
        sc = ops.pmv_scalars_one(self.diffusion, t_int, c_in)
        return fe(g, float(sig), t_model, theta, zeta, sc, x, self._y_for(x.shape[0]))

    def uncond_pred(self, x, sigma):
        """-> (x0_mean [B,3,H,W], x0_var, theta0_var); x0_var is [B] (per-image scalar) or a [B,3,H,W] map."""
        B = x.shape[0]
        sig_list = _host_sigma(sigma, B)
        sig = np.float32(_uniform(sig_list))
        c_in = [float(np.float32(1) / np.sqrt(sig * sig + np.float32(1)))] * B           # external.py:97-100
        t1, tm1 = self._timestep(sig)
        t_int, t_model = [t1] * B, [tm1] * B
        c_in_dev = _dev(c_in, x.device)
        eng = self.inner_model.engine()
        out = eng.forward(x, _dev([float(t) for t in t_model], x.device), x_scale=c_in_dev)
        ct = self.x0_cov_type
        mle = bool(sig < self.mle_sigma_thres)
        sc = ops.pmv_scalars(self.diffusion, t_int, c_in, x.device)
        x0_mean, var = ops.pmv_epilogue(out, x, sc, ops.VAR_CONVERT if (ct == 'convert' and mle) else 0)
        self._ctx = dict(sc=sc, c_in_dev=c_in_dev, eng=eng, token=eng.forward_token)
        r2 = _dev([float(sig ** 2 / (1 + sig ** 2))] * B, x.device)
        if ct == 'convert':
            x0_var = var if mle else r2                                                   # Eq. (22), fused in the epilogue
        elif ct == 'analytic':
            assert self.recon_mse is not None
            if mle:
                idx = int(np.abs(self._recon_sigmas_host - sig).argmin())
                x0_var = self.recon_mse['mse_list'][idx].reshape(1).float().expand(B).contiguous()
            else:
                x0_var = r2
        elif ct == 'pgdm':
            x0_var = r2
        elif ct == 'dps':
            x0_var = torch.zeros(B, device=x.device)
        elif ct == 'diffpir':
            assert self.lambda_ is not None
            x0_var = _dev([float(sig ** 2 / np.float32(self.lambda_))] * B, x.device)
        elif ct == 'tmpd':
            # sigma^2 * grad_x sum(x0_mean): one extra input-VJP seeded with ones (condition.py:268-269)
            g, direct = self._score(x0_mean, torch.ones_like(x0_mean))
            x0_var = ops.lincomb(g, direct, _dev([float(sig ** 2) * c for c in c_in], x.device), _dev([float(sig ** 2)] * B, x.device))
        else:
            raise ValueError('Invalid posterior covariance type.')
        return x0_mean, x0_var, x0_var

    def _score(self, x0_mean, v):
        c = self._ctx
        if c["eng"].forward_token != c["token"]:
            raise RuntimeError("the UNet engine ran another forward since uncond_pred; its saved activations are gone")
        seed, direct = ops.pmv_vjp_seed(x0_mean, v, c["sc"])
        return c["eng"].vjp(seed), direct


class ConditionOpenAIDenoiserV2(ConditionDenoiser):
    """condition.py:277-300 (v2: DWT-Var head, continuous t, no clamp)."""

    def __init__(self, denoiser: OpenAIDenoiserV2, **kwargs):
        super().__init__(**kwargs)
        self.denoiser = denoiser
        ortho_tf_type = kwargs.get('ortho_tf_type', None)
        if ortho_tf_type is not None:
            assert ortho_tf_type == denoiser.ortho_tf_type, "ortho_tf_type must match the one used in the denoiser"

    def uncond_pred(self, x, sigma):
        B = x.shape[0]
        sig_list = _host_sigma(sigma, B)
        sig = np.float32(_uniform(sig_list))
        out, cov = self.denoiser.raw(x, sig_list)
        mle = bool(sig < self.mle_sigma_thres)
        x0_mean, x0_var, theta0_var = ops.v2_epilogue(out, cov, x, _dev(sig_list, x.device), want_var=mle)
        if not mle:
            x0_var = theta0_var = _dev([float(sig ** 2 / (1 + sig ** 2))] * B, x.device)
        eng = self.denoiser.inner_model.engine()
        c_in = [float(np.float32(1) / np.sqrt(sig * sig + np.float32(1)))] * B
        self._ctx = dict(eng=eng, token=eng.forward_token, c_in_dev=_dev(c_in, x.device),
                         sc=ops.v2_vjp_scalars(sig_list, x.device))
        return x0_mean, x0_var, theta0_var

    def _score(self, x0_mean, v):
        """x0 = x - sigma * eps(c_in x, t) is unclamped (condition.py:287-291): seed (-sigma v, 0) on the UNet's 6 output channels
        (the out_cov head does not enter x0_mean), direct term v."""
        c = self._ctx
        if c["eng"].forward_token != c["token"]:
            raise RuntimeError("the UNet engine ran another forward since uncond_pred; its saved activations are gone")
        seed, direct = ops.pmv_vjp_seed(None, v, c["sc"], like=x0_mean)
        return c["eng"].vjp(seed), direct


# ---------------------------------------------
# Implementation of mat solver (computing v)
# ---------------------------------------------

__MAT_SOLVER__ = {}


def register_mat_solver(name):
    def wrapper(func):
        __MAT_SOLVER__[name] = func
        return func
    return wrapper


def _solve(operator, y, x0_mean, theta0_var, ortho_tf):
    """Scalar variance ([B] / 0-dim / [1]) -> closed form; per-element map -> batched device CG (tol 1e-4, <= 1000 its)."""
    B = x0_mean.shape[0]
    if theta0_var.dim() <= 1:
        theta = theta0_var.reshape(-1).float()
        if theta.numel() == 1 and B > 1:
            theta = theta.expand(B)
        return operator.handle.mat_closed(y, x0_mean, theta.contiguous())
    theta = theta0_var
    if theta.shape[0] != B:
        theta = theta.expand(B, *theta.shape[1:])
    return operator.handle.mat_cg(y, x0_mean, theta, ot=ortho_tf.ortho_tf_type, tol=1e-4, maxiter=1000)


@register_mat_solver('inpainting')
@torch.no_grad()
def inpainting_mat(operator, y, x0_mean, theta0_var, ortho_tf=OrthoTransform()):
    """condition.py:317-348."""
    return _solve(operator, y, x0_mean, theta0_var, ortho_tf)


@torch.no_grad()
def _deblur_mat(operator, y, x0_mean, theta0_var, ortho_tf=OrthoTransform()):
    """condition.py:351-386."""
    return _solve(operator, y, x0_mean, theta0_var, ortho_tf)


@register_mat_solver('gaussian_blur')
@torch.no_grad()
def gaussian_blur_mat(operator, y, x0_mean, theta0_var, ortho_tf=OrthoTransform()):
    return _deblur_mat(operator, y, x0_mean, theta0_var, ortho_tf)


@register_mat_solver('motion_blur')
@torch.no_grad()
def motion_blur_mat(operator, y, x0_mean, theta0_var, ortho_tf=OrthoTransform()):
    return _deblur_mat(operator, y, x0_mean, theta0_var, ortho_tf)


@register_mat_solver('super_resolution')
@torch.no_grad()
def super_resolution_mat(operator, y, x0_mean, theta0_var, ortho_tf=OrthoTransform()):
    """condition.py:401-439."""
    return _solve(operator, y, x0_mean, theta0_var, ortho_tf)


# the solvers shipped here: ConditionDenoiser._fused_eval takes the one-call path only while the registry still holds them
_BUILTIN_SOLVERS = dict(__MAT_SOLVER__)
