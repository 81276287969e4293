"""Karras schedule and Euler / Heun samplers (k_diffusion/sampling.py:17-23,46-48,118-135,159-184).

Same signatures as the reference.  The sigma schedule is pulled to the host ONCE per call, so the per-step branch
decisions (churn window, last Euler step) cost no device->host sync; the elementwise updates run in the fused libkdip
kernels (kdip_churn / kdip_euler_step / kdip_heun_step) instead of 4-7 ATen kernels per step.  ``randn_like`` is drawn
every step exactly like the reference (even when gamma == 0), so a given torch seed consumes the Philox stream identically.
"""
import math

import numpy as np
import torch

try:
    from tqdm.auto import trange
except Exception:  # pragma: no cover
    def trange(n, disable=None):
        return range(n)

from kdip import ops

from . import utils


def append_zero(x):
    return torch.cat([x, x.new_zeros([1])])


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7., device='cpu'):
    """Noise schedule of Karras et al. (2022): sampling.py:17-23."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return append_zero(sigmas).to(device)


def to_d(x, sigma, denoised):
    """Karras ODE derivative (sampling.py:46-48)."""
    return (x - denoised) / utils.append_dims(sigma, x.ndim)


class HostSigma:
    """fp32 host copies of the schedule with the reference's rounding (0-dim fp32 tensor arithmetic)."""

    def __init__(self, sigmas):
        self.s = np.asarray(sigmas.detach().float().cpu().numpy(), dtype=np.float32)
        self.n = len(self.s) - 1

    def gamma(self, i, s_churn, s_tmin, s_tmax):
        return min(s_churn / self.n, 2 ** 0.5 - 1) if s_tmin <= self.s[i] <= s_tmax else 0.

    def sigma_hat(self, i, gamma):
        return np.float32(self.s[i] * np.float32(gamma + 1))


def _sigma_arg(x, value):
    """sigma_hat * s_in: a [B] device tensor carrying its host value so kdip denoisers need no device->host read."""
    t = torch.full((x.shape[0],), float(value), device=x.device, dtype=x.dtype)
    t._kdip_host = float(value)
    return t


def _prep(x):
    if not x.is_cuda:
        raise RuntimeError("kdip samplers run on CUDA tensors only (B200, sm_100a); there is no CPU fallback")
    return x.detach().contiguous().float()


def sample_euler(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0., s_tmax=float('inf'),
                 s_noise=1., noise_sampler=None):
    """Algorithm 2 (Euler steps) of Karras et al. (2022): sampling.py:118-135.
    ``noise_sampler(i, x)`` (extension, default ``torch.randn_like``) lets tests inject the per-step noise."""
    extra_args = {} if extra_args is None else extra_args
    hs = HostSigma(sigmas)
    x = _prep(x)
    for i in trange(hs.n, disable=disable):
        gamma = hs.gamma(i, s_churn, s_tmin, s_tmax)
        eps = torch.randn_like(x) if noise_sampler is None else noise_sampler(i, x)
        sigma_hat = hs.sigma_hat(i, gamma)
        if gamma > 0:
            x = ops.churn_(x.clone(), eps, s_noise, hs.s[i], sigma_hat)
        denoised = model(x, _sigma_arg(x, sigma_hat), **extra_args)
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': torch.tensor(sigma_hat), 'denoised': denoised})
        dt = np.float32(hs.s[i + 1] - sigma_hat)
        x = ops.euler_step(x, denoised, sigma_hat, dt)
    return x


def sample_heun(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0., s_tmax=float('inf'),
                s_noise=1., noise_sampler=None):
    """Algorithm 2 (Heun steps) of Karras et al. (2022): sampling.py:159-184."""
    extra_args = {} if extra_args is None else extra_args
    hs = HostSigma(sigmas)
    x = _prep(x)
    for i in trange(hs.n, disable=disable):
        gamma = hs.gamma(i, s_churn, s_tmin, s_tmax)
        eps = torch.randn_like(x) if noise_sampler is None else noise_sampler(i, x)
        sigma_hat = hs.sigma_hat(i, gamma)
        if gamma > 0:
            x = ops.churn_(x.clone(), eps, s_noise, hs.s[i], sigma_hat)
        denoised = model(x, _sigma_arg(x, sigma_hat), **extra_args)
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': torch.tensor(sigma_hat), 'denoised': denoised})
        dt = np.float32(hs.s[i + 1] - sigma_hat)
        if hs.s[i + 1] == 0:
            x = ops.euler_step(x, denoised, sigma_hat, dt)                       # Euler method
        else:
            x_2, d = ops.euler_step(x, denoised, sigma_hat, dt, want_d=True)     # Heun's method
            denoised_2 = model(x_2, _sigma_arg(x, hs.s[i + 1]), **extra_args)
            x = ops.heun_step(x, d, x_2, denoised_2, hs.s[i + 1], dt)
    return x
