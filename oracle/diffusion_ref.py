"""Oracle: DDPM schedule constants, sigma<->t and the p_mean_variance epilogue (test infrastructure only).

Follows guided_diffusion/gaussian_diffusion.py:27-35 (linear betas), :133-169 (constants, float64),
:262-276 + :305-311 + :328-333 (LEARNED_RANGE variance, eps -> x0, clamp) and
k_diffusion/external.py:67-79 (sigma_to_t), :97-100 (get_scalings), :120-123 (sigma table).
"""
import numpy as np
import torch


class Schedule:
    """gaussian_diffusion.py:133-169 with get_named_beta_schedule('linear', 1000) (:27-35), as wrapped by
    SpacedDiffusion (respace.py:63-85, timestep_respacing="" -> all steps, identity timestep_map)."""

    def __init__(self, T=1000):
        scale = 1000 / T
        betas = np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)
        # SpacedDiffusion with all 1000 steps retained re-derives betas from the cumulative products
        # (guided_diffusion/respace.py:68-80): beta_i = 1 - abar_i/abar_{i-1} — differs from linspace in the last ulp.
        ac = np.cumprod(1.0 - betas, axis=0)
        last, nb = 1.0, []
        for a in ac:
            nb.append(1 - a / last)
            last = a
        betas = np.array(nb, dtype=np.float64)
        self.betas = betas
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.log_betas = np.log(betas)
        # external.py:120-123,91: sigma table in float32
        ac = torch.tensor(self.alphas_cumprod, dtype=torch.float32)
        self.sigmas = ((1 - ac) / ac) ** 0.5
        self.log_sigmas = self.sigmas.log()

    def sigma_to_t(self, sigma):
        """external.py:67-79 (quantize=False): fractional t, fp32."""
        log_sigma = sigma.log()
        dists = log_sigma - self.log_sigmas[:, None]
        low_idx = dists.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=self.log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = self.log_sigmas[low_idx], self.log_sigmas[high_idx]
        w = ((low - log_sigma) / (low - high)).clamp(0, 1)
        t = (1 - w) * low_idx + w * high_idx
        return t.view(sigma.shape)

    def extract(self, arr, t):
        """gaussian_diffusion.py:895-908: float64 table -> fp32 scalar per batch element."""
        return torch.from_numpy(arr)[t].float()


def get_scalings(sigma):
    """external.py:97-100: c_out = -sigma, c_in = 1/sqrt(sigma^2+1)."""
    return -sigma, 1 / (sigma ** 2 + 1.0) ** 0.5


def pmv_epilogue(sched: Schedule, model_output, x_in, t):
    """p_mean_variance after the model call, gaussian_diffusion.py:262-276,296-297,305-311.

    model_output [B,6,H,W], x_in = x*c_in [B,3,H,W], t int64 [B] -> (pred_xstart, variance)."""
    C = x_in.shape[1]
    eps, v = torch.split(model_output, C, dim=1)
    e = lambda a: sched.extract(a, t)[:, None, None, None]
    min_log = e(sched.posterior_log_variance_clipped)
    max_log = e(sched.log_betas)
    frac = (v + 1) / 2
    variance = torch.exp(frac * max_log + (1 - frac) * min_log)
    x0 = e(sched.sqrt_recip_alphas_cumprod) * x_in - e(sched.sqrt_recipm1_alphas_cumprod) * eps
    return x0.clamp(-1, 1), variance


def convert_variance(sched: Schedule, variance, t):
    """condition/condition.py:243-246 (Eq. 22): ((var - beta~_t)/coef1_t^2).clip(1e-6)."""
    e = lambda a: sched.extract(a, t)[:, None, None, None]
    return ((variance - e(sched.posterior_variance)) / e(sched.posterior_mean_coef1).pow(2)).clip(min=1e-6)
