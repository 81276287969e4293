"""Oracle: mat solvers, x0-covariance rules and guidance combine (test infrastructure only).

Follows condition/condition.py:83-208 (ConditionDenoiser.forward and the guidance impls), :231-274
(ConditionOpenAIDenoiser.uncond_pred), :287-300 (V2 uncond_pred) and :317-439 (mat solvers).
CG is scipy.sparse.linalg.cg with the legacy ``tol`` semantics of the reference's pinned environment
(python 3.8 => scipy <= 1.10: stop when ||r|| <= tol*||b||), spelled rtol=tol, atol=0 on current scipy.
The oracle is B=1 like the reference (condition.py:84); batches are looped per image.
"""
import numpy as np
import torch
from scipy.sparse.linalg import LinearOperator, cg
from torch.fft import fft2, ifft2

from . import operators_ref as ops
from .diffusion_ref import Schedule, convert_variance, get_scalings, pmv_epilogue
from .transforms_ref import OrthoTransform
from .unet_ref import unet_forward


def _cg(matvec, b, shape, tol=1e-4, maxiter=1000):
    n = int(np.prod(shape))
    iters = [0]

    def mv(u):
        iters[0] += 1
        return matvec(torch.Tensor(u).reshape(shape)).flatten().numpy()

    A = LinearOperator((n, n), matvec=mv, dtype=np.float32)
    u, info = cg(A, b.flatten().numpy(), rtol=tol, atol=0.0, maxiter=maxiter)
    return torch.Tensor(u).reshape(shape), info, iters[0]


def inpainting_mat(operator, y, x0_mean, theta, ot=OrthoTransform(), stats=None):
    """condition.py:317-348."""
    mask = operator.mask
    sigma_s = operator.sigma_s.clip(min=0.001)
    if theta.numel() == 1:
        return (mask * y - mask * x0_mean) / (sigma_s.pow(2) + theta)
    mv = lambda m: sigma_s ** 2 * m + mask * ot.inv(theta * ot(m))
    mat, info, it = _cg(mv, mask * y - mask * x0_mean, x0_mean.shape)
    if stats is not None:
        stats.update(info=info, iters=it)
    return mat


def deblur_mat(operator, y, x0_mean, theta, ot=OrthoTransform(), stats=None):
    """condition.py:351-386."""
    sigma_s = operator.sigma_s.clip(min=0.001)
    FB, FBC, F2B, _ = operator.pre_calculated
    if theta.numel() == 1:
        return ifft2(fft2(y - ifft2(FB * fft2(x0_mean))) / (sigma_s.pow(2) + theta * F2B) * FBC).real
    mv = lambda u: sigma_s ** 2 * u + ifft2(FB * fft2(ot.inv(theta * ot(ifft2(FBC * fft2(u)).real)))).real
    b = y - ifft2(FB * fft2(x0_mean)).real
    u, info, it = _cg(mv, b, y.shape)
    if stats is not None:
        stats.update(info=info, iters=it)
    return ifft2(FBC * fft2(u)).real


def super_resolution_mat(operator, y, x0_mean, theta, ot=OrthoTransform(), stats=None):
    """condition.py:401-439."""
    sigma_s = operator.sigma_s.clip(min=0.001).clip(min=1e-2)
    sf = operator.scale_factor
    FB, FBC, F2B, _ = operator.pre_calculated
    if theta.numel() == 1:
        invW = torch.mean(ops.splits(F2B, sf), dim=-1, keepdim=False)
        return ifft2(FBC * (fft2(y - ops.downsample(ifft2(FB * fft2(x0_mean)), sf)) /
                            (sigma_s.pow(2) + theta * invW)).repeat(1, 1, sf, sf)).real
    mv = lambda u: (sigma_s ** 2 * u + ops.downsample(
        ifft2(FB * fft2(ot.inv(theta * ot(ifft2(FBC * fft2(ops.upsample(u, sf))).real)))), sf)).real
    b = (y - ops.downsample(ifft2(FB * fft2(x0_mean)), sf)).real
    u, info, it = _cg(mv, b, y.shape)
    if stats is not None:
        stats.update(info=info, iters=it)
    return ifft2(FBC * fft2(ops.upsample(u, sf))).real


MAT_SOLVER = {"inpainting": inpainting_mat, "gaussian_blur": deblur_mat, "motion_blur": deblur_mat,
              "super_resolution": super_resolution_mat}


class ConditionDenoiserRef:
    """ConditionOpenAIDenoiser (condition.py:211-274) on top of the functional UNet oracle."""

    def __init__(self, sd, cfg, operator, measurement, guidance, x0_cov_type="pgdm", recon_mse=None,
                 zeta=None, lambda_=None, mle_sigma_thres=0.2, ortho_tf_type=None, eta=None, num_hutchinson_samples=None):
        self.sd, self.cfg = sd, cfg
        self.sched = Schedule()
        self.operator = operator
        self.y = measurement[0] if isinstance(measurement, tuple) else measurement
        self.guidance, self.x0_cov_type = guidance, x0_cov_type
        self.recon_mse, self.zeta, self.lambda_ = recon_mse, zeta, lambda_
        self.eta, self.num_hutchinson_samples = eta, num_hutchinson_samples
        self.thres = mle_sigma_thres
        self.ortho_tf_type = ortho_tf_type
        self.ot = OrthoTransform(ortho_tf_type)
        self.mat_solver = MAT_SOLVER[operator.name]
        self.last = {}

    def uncond_pred(self, x, sigma):
        """condition.py:231-274."""
        c_out, c_in = get_scalings(sigma)
        t = self.sched.sigma_to_t(sigma).long()
        out = unet_forward(self.sd, self.cfg, x * c_in, t)
        x0_mean, variance = pmv_epilogue(self.sched, out, x * c_in, t)
        r2 = sigma.pow(2) / (1 + sigma.pow(2))
        ct = self.x0_cov_type
        if ct == "convert":
            x0_var = convert_variance(self.sched, variance, t) if sigma < self.thres else r2
        elif ct == "analytic":
            if sigma < self.thres:
                idx = (self.recon_mse["sigmas"] - sigma[0]).abs().argmin()
                x0_var = self.recon_mse["mse_list"][idx]
            else:
                x0_var = r2
        elif ct == "pgdm":
            x0_var = r2
        elif ct == "dps":
            x0_var = torch.zeros(1)
        elif ct == "diffpir":
            x0_var = sigma.pow(2) / self.lambda_
        elif ct == "tmpd":
            x0_var = torch.autograd.grad(x0_mean.sum(), x, retain_graph=True)[0] * sigma.pow(2)
        else:
            raise ValueError("Invalid posterior covariance type.")
        return x0_mean, x0_var, x0_var

    def _mat(self, x0_mean, theta):
        with torch.no_grad():
            st = {}
            m = self.mat_solver(self.operator, self.y, x0_mean.detach(), theta.detach(), self.ot, stats=st)
            self.last.update(st)
            return m

    def __call__(self, x, sigma):
        """condition.py:83-131."""
        assert x.shape[0] == 1
        g = self.guidance
        if g in ("dps+mle", "pgdm+mle", "stsl+mle"):
            g = "I" if sigma < self.thres else g.split("+")[0]
        if g == "uncond":
            with torch.no_grad():
                hat = self.uncond_pred(x, sigma)[0]
        elif g == "I":
            x = x.detach().requires_grad_()
            x0_mean, x0_var, th0 = self.uncond_pred(x, sigma)
            mat = self._mat(x0_mean, x0_var if self.ortho_tf_type is None else th0)
            score = torch.autograd.grad((mat.detach() * x0_mean).sum(), x)[0]
            hat = x0_mean + sigma.pow(2) * score
        elif g == "pgdm":
            x = x.detach().requires_grad_()
            x0_mean = self.uncond_pred(x, sigma)[0]
            x0_var = sigma.pow(2) / (1 + sigma.pow(2))
            mat = self._mat(x0_mean, x0_var)
            score = torch.autograd.grad((mat.detach() * x0_mean).sum(), x)[0] * x0_var
            hat = x0_mean + sigma.pow(2) * score
        elif g == "dps":
            x = x.detach().requires_grad_()
            x0_mean = self.uncond_pred(x, sigma)[0]
            diff = self.y - self.operator.forward(x0_mean, noiseless=True)
            norm = torch.linalg.norm(diff)
            score = -torch.autograd.grad(norm, x)[0] * self.zeta
            hat = x0_mean + sigma.pow(2) * score
        elif g == "stsl":
            # condition.py:185-208: first-order data term + Hutchinson estimate of the trace of the x0 Jacobian;
            # eps is drawn from torch's global generator, one draw per sample, after the first uncond_pred
            x = x.detach().requires_grad_()
            x0_mean = self.uncond_pred(x, sigma)[0]
            first = -torch.linalg.norm(self.y - self.operator.forward(x0_mean, noiseless=True))
            second = 0
            for _ in range(self.num_hutchinson_samples):
                eps = torch.randn_like(x)
                second = second + -((self.uncond_pred(x + eps, sigma)[0] - x0_mean) * eps).sum() * sigma.pow(2)
            second = second / self.num_hutchinson_samples
            loss = self.zeta * first + (self.eta / x.numel()) * second
            hat = x0_mean + sigma.pow(2) * torch.autograd.grad(loss, x)[0]
        elif g == "diffpir":
            with torch.no_grad():
                x0_mean = self.uncond_pred(x, sigma)[0]
                x0_var = sigma.pow(2) / self.lambda_
                hat = x0_mean + self._mat(x0_mean, x0_var) * x0_var
        elif g == "II":
            with torch.no_grad():
                x0_mean, x0_var, th0 = self.uncond_pred(x, sigma)
                th = x0_var if self.ortho_tf_type is None else th0
                mat = self._mat(x0_mean, th)
                hat = x0_mean + self.ot.inv(self.ot(mat) * th)
        else:
            raise ValueError(f"Invalid guidance type: '{self.guidance}'.")
        return hat.clip(-1, 1).detach()


class ConditionDenoiserV2Ref(ConditionDenoiserRef):
    """ConditionOpenAIDenoiserV2 (condition/condition.py:277-300) on OpenAIDenoiserV2.forward(return_variance=True)
    (k_diffusion/external.py:161-169): continuous timestep, no clamp, x0 = x + c_out * eps, per-pixel variances from the
    ``out_cov`` 1x1 conv on the pre-head feature (pixel domain ``logvar`` and transform domain ``logvar_ot``)."""

    def __init__(self, sd, cfg, cov_w, cov_b, operator, measurement, guidance, mle_sigma_thres=1.0, ortho_tf_type=None, **kw):
        super().__init__(sd, cfg, operator, measurement, guidance, mle_sigma_thres=mle_sigma_thres, ortho_tf_type=ortho_tf_type, **kw)
        self.cov_w, self.cov_b = cov_w, cov_b

    def denoiser(self, x, sigma):
        """external.py:161-169 with return_variance=True."""
        c_out, c_in = get_scalings(sigma)
        out, feat = unet_forward(self.sd, self.cfg, x * c_in, self.sched.sigma_to_t(sigma), return_feature=True)
        logvar, logvar_ot = torch.nn.functional.conv2d(feat, self.cov_w, self.cov_b).chunk(2, dim=1)
        return out.chunk(2, dim=1)[0], logvar, logvar_ot

    def uncond_pred(self, x, sigma):
        """condition.py:287-300."""
        c_out, _ = get_scalings(sigma)
        model_output, logvar, logvar_ot = self.denoiser(x, sigma)
        x0_mean = model_output * c_out + x
        if sigma < self.thres:
            x0_var = logvar.exp() * c_out.pow(2)
            theta0_var = logvar_ot.exp() * c_out.pow(2)
        else:
            x0_var = theta0_var = sigma.pow(2) / (1 + sigma.pow(2))
        return x0_mean, x0_var, theta0_var
