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
    """Every sampler of this module is CUDA-only (the updates are libkdip kernels); a CPU tensor raises."""
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
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': torch.tensor(sigma_hat, device=x.device), 'denoised': denoised})
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
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': torch.tensor(sigma_hat, device=x.device), 'denoised': denoised})
        dt = np.float32(hs.s[i + 1] - sigma_hat)
        if hs.s[i + 1] == 0:
            x = ops.euler_step(x, denoised, sigma_hat, dt)                       # Euler method
        else:
            x_2, d = ops.euler_step(x, denoised, sigma_hat, dt, want_d=True)     # Heun's method
            denoised_2 = model(x_2, _sigma_arg(x, hs.s[i + 1]), **extra_args)
            x = ops.heun_step(x, d, x_2, denoised_2, hs.s[i + 1], dt)
    return x


# ---- the other schedules and samplers of k_diffusion/sampling.py ----------------------------------------
# Same signatures and callback dictionaries as the reference.  Every step coefficient is computed on the host in fp32 from the
# host copy of the schedule (the reference does the same arithmetic on 0-dim fp32 tensors), and every state update is ONE fused
# kernel: kdip_euler_step for "x + (x - denoised)/sigma * dt", kdip_lincomb3 for the exponential-integrator forms.

def get_sigmas_exponential(n, sigma_min, sigma_max, device='cpu'):
    """Log-linear schedule (sampling.py:26-29)."""
    return append_zero(torch.linspace(math.log(sigma_max), math.log(sigma_min), n).exp()).to(device)


def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1., device='cpu'):
    """Polynomial-in-log-sigma schedule (sampling.py:32-36)."""
    lo, hi = math.log(sigma_min), math.log(sigma_max)
    ramp = torch.linspace(1, 0, n) ** rho
    return append_zero(torch.exp(ramp * (hi - lo) + lo)).to(device)


def get_sigmas_vp(n, beta_d=19.9, beta_min=0.1, eps_s=1e-3, device='cpu'):
    """Continuous VP schedule (sampling.py:39-43)."""
    t = torch.linspace(1, eps_s, n)
    return append_zero(torch.sqrt(torch.exp(beta_d * t ** 2 / 2 + beta_min * t) - 1)).to(device)


_F = np.float32


def get_ancestral_step(sigma_from, sigma_to, eta=1.):
    """(sigma_down, sigma_up) of an ancestral step (sampling.py:51-58), fp32 host arithmetic."""
    if not eta:
        return sigma_to, 0.
    sf, st = _F(sigma_from), _F(sigma_to)
    with np.errstate(invalid='ignore', divide='ignore'):
        up = min(st, _F(eta) * np.sqrt(st ** 2 * (sf ** 2 - st ** 2) / sf ** 2))
        down = np.sqrt(st ** 2 - _F(up) ** 2)
    return _F(down), _F(up)


def _log_midpoint(a, b):
    """exp(lerp(log a, log b, 0.5)) as torch evaluates it for weight 0.5: end - (end - start) * (1 - weight)."""
    la, lb = np.log(_F(a)), np.log(_F(b))
    return _F(np.exp(_F(lb - _F(lb - la) * _F(0.5))))


def _noise(noise_sampler, sigmas, i, x):
    return torch.randn_like(x) if noise_sampler is None else noise_sampler(sigmas[i], sigmas[i + 1]).to(x.device, x.dtype)


def _report(callback, x, i, sigmas, sigma_hat, denoised):
    if callback is not None:
        callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigma_hat, 'denoised': denoised})


def sample_euler_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1., noise_sampler=None):
    """Ancestral Euler steps (sampling.py:139-156)."""
    extra_args = {} if extra_args is None else extra_args
    hs = HostSigma(sigmas)
    x = _prep(x)
    for i in trange(hs.n, disable=disable):
        denoised = model(x, _sigma_arg(x, hs.s[i]), **extra_args)
        down, up = get_ancestral_step(hs.s[i], hs.s[i + 1], eta=eta)
        _report(callback, x, i, sigmas, sigmas[i], denoised)
        x = ops.euler_step(x, denoised, hs.s[i], _F(down - hs.s[i]))
        if hs.s[i + 1] > 0:
            x = ops.lincomb3(x, 1.0, _noise(noise_sampler, sigmas, i, x), _F(s_noise) * _F(up))
    return x


def _second_order_dpm2(model, x, denoised, sigma, sigma_to, extra_args):
    """One DPM-Solver-2 step from sigma to sigma_to > 0 through the log-midpoint (sampling.py:204-213,235-244)."""
    mid = _log_midpoint(sigma, sigma_to)
    x_2 = ops.euler_step(x, denoised, sigma, _F(mid - sigma))
    denoised_2 = model(x_2, _sigma_arg(x, mid), **extra_args)
    k = _F(_F(sigma_to - sigma) / mid)                         # x + (x_2 - denoised_2) / mid * dt_2
    return ops.lincomb3(x, 1.0, x_2, k, denoised_2, -k)


def sample_dpm_2(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0., s_tmax=float('inf'),
                 s_noise=1., noise_sampler=None):
    """DPM-Solver-2-style steps with the churn of Karras et al. (sampling.py:187-215).  ``noise_sampler(i, x)`` as in sample_heun."""
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
        _report(callback, x, i, sigmas, torch.tensor(sigma_hat, device=x.device), denoised)
        if hs.s[i + 1] == 0:
            x = ops.euler_step(x, denoised, sigma_hat, _F(hs.s[i + 1] - sigma_hat))
        else:
            x = _second_order_dpm2(model, x, denoised, sigma_hat, hs.s[i + 1], extra_args)
    return x


def sample_dpm_2_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1., noise_sampler=None):
    """Ancestral DPM-Solver-2 steps (sampling.py:218-248)."""
    extra_args = {} if extra_args is None else extra_args
    hs = HostSigma(sigmas)
    x = _prep(x)
    for i in trange(hs.n, disable=disable):
        denoised = model(x, _sigma_arg(x, hs.s[i]), **extra_args)
        down, up = get_ancestral_step(hs.s[i], hs.s[i + 1], eta=eta)
        _report(callback, x, i, sigmas, sigmas[i], denoised)
        if down == 0:
            x = ops.euler_step(x, denoised, hs.s[i], _F(down - hs.s[i]))
        else:
            x = _second_order_dpm2(model, x, denoised, hs.s[i], _F(down), extra_args)
            x = ops.lincomb3(x, 1.0, _noise(noise_sampler, sigmas, i, x), _F(s_noise) * _F(up))
    return x


def linear_multistep_coeff(order, t, i, j):
    """Integral over [t_i, t_{i+1}] of the j-th Lagrange basis polynomial through t_i .. t_{i-order+1} (sampling.py:251-261)."""
    from scipy import integrate
    if order - 1 > i:
        raise ValueError(f'Order {order} too high for step {i}')
    nodes = [t[i - k] for k in range(order)]

    def basis(tau):
        p = 1.
        for k, node in enumerate(nodes):
            if k != j:
                p *= (tau - node) / (nodes[j] - node)
        return p
    return integrate.quad(basis, t[i], t[i + 1], epsrel=1e-4)[0]


def sample_lms(model, x, sigmas, extra_args=None, callback=None, disable=None, order=4):
    """Linear multistep (Adams-Bashforth in sigma) sampler (sampling.py:259-275)."""
    extra_args = {} if extra_args is None else extra_args
    hs = HostSigma(sigmas)
    x = _prep(x)
    ds = []                                                     # newest first
    for i in trange(hs.n, disable=disable):
        denoised = model(x, _sigma_arg(x, hs.s[i]), **extra_args)
        inv = _F(1) / hs.s[i]
        ds.insert(0, ops.lincomb3(x, inv, denoised, -inv))      # to_d
        del ds[order:]
        _report(callback, x, i, sigmas, sigmas[i], denoised)
        cur = min(i + 1, order)
        c = [linear_multistep_coeff(cur, hs.s, i, j) for j in range(cur)]
        x = ops.lincomb3(x, 1.0, ds[0], c[0], ds[1] if cur > 1 else None, c[1] if cur > 1 else 0.)
        if cur > 2:
            x = ops.lincomb3(x, 1.0, ds[2], c[2], ds[3] if cur > 3 else None, c[3] if cur > 3 else 0., out=x)
        for j in range(4, cur):                                 # order > 4: one more term per kernel
            x = ops.lincomb3(x, 1.0, ds[j], c[j], out=x)
    return x


def _t_of(sigma):
    return -np.log(_F(sigma))


def sample_dpmpp_2s_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1., noise_sampler=None):
    """Ancestral DPM-Solver++(2S) steps (sampling.py:507-538)."""
    extra_args = {} if extra_args is None else extra_args
    hs = HostSigma(sigmas)
    x = _prep(x)
    for i in trange(hs.n, disable=disable):
        denoised = model(x, _sigma_arg(x, hs.s[i]), **extra_args)
        down, up = get_ancestral_step(hs.s[i], hs.s[i + 1], eta=eta)
        _report(callback, x, i, sigmas, sigmas[i], denoised)
        if down == 0:
            x = ops.euler_step(x, denoised, hs.s[i], _F(down - hs.s[i]))
        else:
            t, t_next = _t_of(hs.s[i]), _t_of(down)
            h = _F(t_next - t)
            s = _F(t + _F(0.5) * h)
            sig_s, sig_t = _F(np.exp(-s)), _F(np.exp(-t))
            x_2 = ops.lincomb3(x, _F(sig_s / sig_t), denoised, -_F(np.expm1(_F(-h * _F(0.5)))))
            denoised_2 = model(x_2, _sigma_arg(x, sig_s), **extra_args)
            x = ops.lincomb3(x, _F(_F(np.exp(-t_next)) / sig_t), denoised_2, -_F(np.expm1(-h)))
        if hs.s[i + 1] > 0:
            x = ops.lincomb3(x, 1.0, _noise(noise_sampler, sigmas, i, x), _F(s_noise) * _F(up))
    return x


def sample_dpmpp_2m(model, x, sigmas, extra_args=None, callback=None, disable=None):
    """DPM-Solver++(2M) (sampling.py:583-606): the previous denoised output extrapolates the current one."""
    extra_args = {} if extra_args is None else extra_args
    hs = HostSigma(sigmas)
    x = _prep(x)
    old = None
    for i in trange(hs.n, disable=disable):
        denoised = model(x, _sigma_arg(x, hs.s[i]), **extra_args)
        _report(callback, x, i, sigmas, sigmas[i], denoised)
        t = _t_of(hs.s[i])
        if hs.s[i + 1] == 0:
            ratio, e = _F(0), _F(-1)                            # sigma_fn(inf) = 0, expm1(-inf) = -1: x = denoised
        else:
            t_next = _t_of(hs.s[i + 1])
            h = _F(t_next - t)
            ratio, e = _F(_F(np.exp(-t_next)) / _F(np.exp(-t))), _F(np.expm1(-h))
        if old is None or hs.s[i + 1] == 0:
            x = ops.lincomb3(x, ratio, denoised, -e)
        else:
            r = _F(_F(t - _t_of(hs.s[i - 1])) / h)
            w = _F(_F(1) / _F(_F(2) * r))
            x = ops.lincomb3(x, ratio, denoised, -_F(e * _F(_F(1) + w)), old, _F(e * w))
        old = denoised
    return x
