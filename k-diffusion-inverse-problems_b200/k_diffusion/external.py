"""sigma-space wrappers around the DDPM eps-model (k_diffusion/external.py:42-169): discrete schedule, sigma <-> t,
c_in / c_out, ``OpenAIDenoiser`` and ``OpenAIDenoiserV2`` (DWT-Var covariance head)."""
import numpy as np
import torch
from torch import nn

from kdip import ops

from . import sampling, utils


class DiscreteSchedule(nn.Module):
    """The DDPM noise levels as a table: sigma <-> (fractional) timestep by piecewise-linear interpolation in log sigma
    (external.py:42-85).  Buffers ``sigmas`` / ``log_sigmas`` keep the reference's names (they are in its checkpoints)."""

    def __init__(self, sigmas, quantize):
        super().__init__()
        self.quantize = quantize
        self.register_buffer('sigmas', sigmas)
        self.register_buffer('log_sigmas', sigmas.log())
        self._log_sigmas_host = self.log_sigmas.detach().cpu().numpy().astype(np.float32)

    sigma_min = property(lambda self: self.sigmas[0])
    sigma_max = property(lambda self: self.sigmas[-1])

    def get_sigmas(self, n=None):
        """All table entries from the noisiest down (n is None) or n levels evenly spaced in t, zero-terminated (:57-62)."""
        if n is not None:
            steps = torch.linspace(len(self.sigmas) - 1, 0, n, device=self.sigmas.device)
            return sampling.append_zero(self.t_to_sigma(steps))
        return sampling.append_zero(self.sigmas.flip(0))

    def _bracket(self, log_sigma):
        """Index of the last table entry <= log_sigma (0 when there is none), capped so that index + 1 exists: what the
        reference's ``dists.ge(0).cumsum(0).argmax(0).clamp(max=n-2)`` selects on the ascending table (:71)."""
        below = (log_sigma.reshape(1, -1) >= self.log_sigmas[:, None]).sum(dim=0)
        return (below - 1).clamp(min=0, max=self.log_sigmas.shape[0] - 2)

    def sigma_to_t(self, sigma, quantize=None):
        """Nearest table index (quantize) or the interpolated fractional index (:67-79)."""
        log_sigma = sigma.log()
        if self.quantize if quantize is None else quantize:
            return (log_sigma.reshape(1, -1) - self.log_sigmas[:, None]).abs().argmin(dim=0).view(sigma.shape)
        lo = self._bracket(log_sigma)
        ls_lo, ls_hi = self.log_sigmas[lo], self.log_sigmas[lo + 1]
        w = ((ls_lo - log_sigma.reshape(-1)) / (ls_lo - ls_hi)).clamp(0, 1)
        return ((1 - w) * lo + w * (lo + 1)).view(sigma.shape)

    def sigma_to_t_host(self, sigma):
        """The same inverse evaluated on the host in fp32 (no device work, no sync): sigma float -> t float."""
        table = self._log_sigmas_host
        with np.errstate(divide='ignore'):          # sigma = 0 (analytic_variance.py evaluates the trailing 0): log -> -inf, t -> 0
            log_sigma = np.log(np.float32(sigma))
        lo = min(max(int(np.count_nonzero(log_sigma >= table)) - 1, 0), len(table) - 2)
        w = np.float32(np.clip((table[lo] - log_sigma) / (table[lo] - table[lo + 1]), 0, 1))
        return float(np.float32((np.float32(1) - w) * np.float32(lo) + w * np.float32(lo + 1)))

    def t_to_sigma(self, t):
        """Interpolate log sigma between the two neighbouring integer timesteps (:81-85)."""
        t = t.float()
        frac = t.frac()
        return ((1 - frac) * self.log_sigmas[t.floor().long()] + frac * self.log_sigmas[t.ceil().long()]).exp()


class DiscreteEpsDDPMDenoiser(DiscreteSchedule):
    """An eps-prediction DDPM model seen as a Karras denoiser: D(x, sigma) = x - sigma * eps(x / sqrt(sigma^2 + 1), t(sigma))
    (external.py:88-115; forward only, the training loss is out of scope)."""

    sigma_data = 1.

    def __init__(self, model, alphas_cumprod, quantize):
        super().__init__(((1 - alphas_cumprod) / alphas_cumprod) ** 0.5, quantize)
        self.inner_model = model

    def get_scalings(self, sigma):
        """(c_out, c_in) of :97-100."""
        return -sigma, 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5

    def get_eps(self, *args, **kwargs):
        return self.inner_model(*args, **kwargs)

    def forward(self, input, sigma, **kwargs):
        c_out, c_in = (utils.append_dims(c, input.ndim) for c in self.get_scalings(sigma))
        return input + c_out * self.get_eps(input * c_in, self.sigma_to_t(sigma), **kwargs)


def _alphas_cumprod(diffusion, device):
    return torch.tensor(diffusion.alphas_cumprod, device=device, dtype=torch.float32)


class OpenAIDenoiser(DiscreteEpsDDPMDenoiser):
    """Wrapper for the guided-diffusion UNets, which emit (eps, variance interpolation) on 6 channels (external.py:117-132)."""

    def __init__(self, model, diffusion, quantize=False, has_learned_sigmas=True, device='cpu'):
        super().__init__(model, _alphas_cumprod(diffusion, device), quantize=quantize)
        self.has_learned_sigmas = has_learned_sigmas

    def get_eps(self, *args, **kwargs):
        out = self.inner_model(*args, **kwargs)
        if not self.has_learned_sigmas or kwargs.get('return_variance', False):
            return out
        return out.chunk(2, dim=1)[0]


class OpenAIDenoiserV2(DiscreteEpsDDPMDenoiser):
    """external.py:135-169: adds the ``out_cov`` 1x1 conv head -> (logvar, logvar_ot).  The head is fused into the
    UNet engine (it reads the pre-head feature in place instead of exporting [B,128,256,256] fp32)."""

    def __init__(self, model, diffusion, quantize=False, device='cpu', ortho_tf_type=None):
        from condition.utils import OrthoTransform
        super().__init__(model, _alphas_cumprod(diffusion, device), quantize=quantize)
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
