"""sigma-space wrappers around the DDPM eps-model (k_diffusion/external.py:42-169): discrete schedule, sigma <-> t,
c_in / c_out, ``OpenAIDenoiser`` and ``OpenAIDenoiserV2`` (DWT-Var covariance head)."""
import numpy as np
import torch
from torch import nn

from kdip import ops

from . import sampling, utils


class DiscreteSchedule(nn.Module):
    """external.py:42-85."""

    def __init__(self, sigmas, quantize):
        super().__init__()
        self.register_buffer('sigmas', sigmas)
        self.register_buffer('log_sigmas', sigmas.log())
        self.quantize = quantize
        self._log_sigmas_host = self.log_sigmas.detach().cpu().numpy().astype(np.float32)

    @property
    def sigma_min(self):
        return self.sigmas[0]

    @property
    def sigma_max(self):
        return self.sigmas[-1]

    def get_sigmas(self, n=None):
        if n is None:
            return sampling.append_zero(self.sigmas.flip(0))
        t_max = len(self.sigmas) - 1
        t = torch.linspace(t_max, 0, n, device=self.sigmas.device)
        return sampling.append_zero(self.t_to_sigma(t))

    def sigma_to_t(self, sigma, quantize=None):
        quantize = self.quantize if quantize is None else quantize
        log_sigma = sigma.log()
        dists = log_sigma - self.log_sigmas[:, None]
        if quantize:
            return dists.abs().argmin(dim=0).view(sigma.shape)
        low_idx = dists.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=self.log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = self.log_sigmas[low_idx], self.log_sigmas[high_idx]
        w = ((low - log_sigma) / (low - high)).clamp(0, 1)
        t = (1 - w) * low_idx + w * high_idx
        return t.view(sigma.shape)

    def sigma_to_t_host(self, sigma):
        """Same piecewise-linear inverse evaluated on the host in fp32 (no device work, no sync): sigma float -> t float."""
        ls = self._log_sigmas_host
        with np.errstate(divide='ignore'):          # sigma = 0 (analytic_variance.py evaluates the trailing 0): log -> -inf, t -> 0
            log_sigma = np.log(np.float32(sigma))
        low_idx = min(int(np.count_nonzero(log_sigma - ls >= 0)) - 1, len(ls) - 2)
        if np.count_nonzero(log_sigma - ls >= 0) == 0:
            low_idx = 0   # cumsum().argmax() of an all-False column is 0
        high_idx = low_idx + 1
        low, high = ls[low_idx], ls[high_idx]
        w = np.float32(np.clip((low - log_sigma) / (low - high), 0, 1))
        return float(np.float32((np.float32(1) - w) * np.float32(low_idx) + w * np.float32(high_idx)))

    def t_to_sigma(self, t):
        t = t.float()
        low_idx, high_idx, w = t.floor().long(), t.ceil().long(), t.frac()
        log_sigma = (1 - w) * self.log_sigmas[low_idx] + w * self.log_sigmas[high_idx]
        return log_sigma.exp()


class DiscreteEpsDDPMDenoiser(DiscreteSchedule):
    """external.py:88-115 (forward only; the training loss is out of scope)."""

    def __init__(self, model, alphas_cumprod, quantize):
        super().__init__(((1 - alphas_cumprod) / alphas_cumprod) ** 0.5, quantize)
        self.inner_model = model
        self.sigma_data = 1.

    def get_scalings(self, sigma):
        c_out = -sigma
        c_in = 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        return c_out, c_in

    def get_eps(self, *args, **kwargs):
        return self.inner_model(*args, **kwargs)

    def forward(self, input, sigma, **kwargs):
        c_out, c_in = [utils.append_dims(x, input.ndim) for x in self.get_scalings(sigma)]
        eps = self.get_eps(input * c_in, self.sigma_to_t(sigma), **kwargs)
        return input + eps * c_out


class OpenAIDenoiser(DiscreteEpsDDPMDenoiser):
    """external.py:117-132."""

    def __init__(self, model, diffusion, quantize=False, has_learned_sigmas=True, device='cpu'):
        alphas_cumprod = torch.tensor(diffusion.alphas_cumprod, device=device, dtype=torch.float32)
        super().__init__(model, alphas_cumprod, quantize=quantize)
        self.has_learned_sigmas = has_learned_sigmas

    def get_eps(self, *args, **kwargs):
        model_output = self.inner_model(*args, **kwargs)
        if self.has_learned_sigmas:
            if kwargs.get('return_variance', False):
                return model_output
            return model_output.chunk(2, dim=1)[0]
        return model_output


class OpenAIDenoiserV2(DiscreteEpsDDPMDenoiser):
    """external.py:135-169: adds the ``out_cov`` 1x1 conv head -> (logvar, logvar_ot).  The head is fused into the
    UNet engine (it reads the pre-head feature in place instead of exporting [B,128,256,256] fp32)."""

    def __init__(self, model, diffusion, quantize=False, device='cpu', ortho_tf_type=None):
        from condition.utils import OrthoTransform
        alphas_cumprod = torch.tensor(diffusion.alphas_cumprod, device=device, dtype=torch.float32)
        super().__init__(model, alphas_cumprod, quantize=quantize)
        self.out_cov = nn.Conv2d(model.model_channels * int(model.channel_mult[0]), 2 * 3, 1)
        self.ortho_tf_type = ortho_tf_type
        self.ortho_tf = OrthoTransform(ortho_tf_type)

    def raw(self, input, sigma_host):
        """(unet_out [B,6,H,W], cov_out [B,6,H,W]) for host sigmas (list of floats), continuous t."""
        self.inner_model.out_cov = (self.out_cov.weight, self.out_cov.bias)
        eng = self.inner_model.engine()
        dev = input.device
        c_in = torch.tensor([1.0 / (s * s + 1.0) ** 0.5 for s in sigma_host], device=dev, dtype=torch.float32)
        t = torch.tensor([self.sigma_to_t_host(s) for s in sigma_host], device=dev, dtype=torch.float32)
        return eng.forward(input, t, x_scale=c_in, want_cov=True)

    def forward(self, input, sigma, return_variance=False):
        host = getattr(sigma, '_kdip_host', None)
        sig = [host] * input.shape[0] if host is not None else [float(v) for v in sigma.tolist()]
        out, cov = self.raw(input.detach().contiguous().float(), sig)
        model_output = out[:, :3]
        logvar, logvar_ot = cov.chunk(2, dim=1)
        if return_variance:
            return model_output, logvar, logvar_ot
        c_out = utils.append_dims(self.get_scalings(sigma)[0], input.ndim)
        return input + model_output * c_out
