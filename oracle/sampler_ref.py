"""Oracle: Karras schedule and Euler / Heun sampler loops (test infrastructure only).

Follows k_diffusion/sampling.py:17-23 (get_sigmas_karras), :46-48 (to_d), :118-135 (sample_euler),
:159-184 (sample_heun).  The per-step ``randn_like`` draw is replaced by an injectable ``noise_fn(i, x)`` so
CPU (mt19937) and CUDA (Philox) paths can be compared on identical noise (SURVEY.md §7 hard parts).
"""
import torch


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0):
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def _default_noise(i, x):
    return torch.randn_like(x)


def sample_euler(model, x, sigmas, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, noise_fn=_default_noise):
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        gamma = min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
        eps = noise_fn(i, x) * s_noise
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            x = x + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = model(x, sigma_hat * s_in)
        d = (x - denoised) / sigma_hat
        x = x + d * (sigmas[i + 1] - sigma_hat)
    return x


def sample_heun(model, x, sigmas, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, noise_fn=_default_noise):
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        gamma = min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
        eps = noise_fn(i, x) * s_noise
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            x = x + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = model(x, sigma_hat * s_in)
        d = (x - denoised) / sigma_hat
        dt = sigmas[i + 1] - sigma_hat
        if sigmas[i + 1] == 0:
            x = x + d * dt
        else:
            x_2 = x + d * dt
            denoised_2 = model(x_2, sigmas[i + 1] * s_in)
            d_2 = (x_2 - denoised_2) / sigmas[i + 1]
            x = x + (d + d_2) / 2 * dt
    return x


# ---- the other schedules / samplers of k_diffusion/sampling.py --------------------------------------
# Restated with the reference's 0-dim fp32 tensor arithmetic; pinned by tests/golden/golden_samplers.npz (outputs of the
# reference's own functions, tests/golden/make_golden_samplers.py).  ``noise_fn(i, x)`` replaces every random draw of step i.
import math

from scipy import integrate


def _zero_terminated(s):
    return torch.cat([s, s.new_zeros([1])])


def get_sigmas_exponential(n, sigma_min, sigma_max):                      # sampling.py:26-29
    return _zero_terminated(torch.linspace(math.log(sigma_max), math.log(sigma_min), n).exp())


def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1.0):         # sampling.py:32-36
    ramp = torch.linspace(1, 0, n) ** rho
    return _zero_terminated(torch.exp(ramp * (math.log(sigma_max) - math.log(sigma_min)) + math.log(sigma_min)))


def get_sigmas_vp(n, beta_d=19.9, beta_min=0.1, eps_s=1e-3):              # sampling.py:39-43
    t = torch.linspace(1, eps_s, n)
    return _zero_terminated(torch.sqrt(torch.exp(beta_d * t ** 2 / 2 + beta_min * t) - 1))


def ancestral_step(sigma_from, sigma_to, eta=1.0):                        # sampling.py:51-58
    if not eta:
        return sigma_to, 0.0
    var_ratio = sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2
    up = min(sigma_to, eta * var_ratio ** 0.5)
    return (sigma_to ** 2 - up ** 2) ** 0.5, up


def _ones(x):
    return x.new_ones([x.shape[0]])


def _dpm2_step(model, x, denoised, sigma, sigma_to):
    """x -> sigma_to via the geometric midpoint of (sigma, sigma_to): sampling.py:204-213 / :235-244."""
    d = (x - denoised) / sigma
    mid = sigma.log().lerp(sigma_to.log(), 0.5).exp()
    x_mid = x + d * (mid - sigma)
    d_mid = (x_mid - model(x_mid, mid * _ones(x))) / mid
    return x + d_mid * (sigma_to - sigma)


def sample_euler_ancestral(model, x, sigmas, eta=1.0, s_noise=1.0, noise_fn=_default_noise):      # sampling.py:139-156
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * _ones(x))
        down, up = ancestral_step(sigmas[i], sigmas[i + 1], eta)
        x = x + (x - denoised) / sigmas[i] * (down - sigmas[i])
        if sigmas[i + 1] > 0:
            x = x + noise_fn(i, x) * s_noise * up
    return x


def sample_dpm_2(model, x, sigmas, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, noise_fn=_default_noise):  # :187-215
    n = len(sigmas) - 1
    for i in range(n):
        gamma = min(s_churn / n, 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
        eps = noise_fn(i, x) * s_noise
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            x = x + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = model(x, sigma_hat * _ones(x))
        if sigmas[i + 1] == 0:
            x = x + (x - denoised) / sigma_hat * (sigmas[i + 1] - sigma_hat)
        else:
            x = _dpm2_step(model, x, denoised, sigma_hat, sigmas[i + 1])
    return x


def sample_dpm_2_ancestral(model, x, sigmas, eta=1.0, s_noise=1.0, noise_fn=_default_noise):      # sampling.py:218-248
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * _ones(x))
        down, up = ancestral_step(sigmas[i], sigmas[i + 1], eta)
        if down == 0:
            x = x + (x - denoised) / sigmas[i] * (down - sigmas[i])
        else:
            x = _dpm2_step(model, x, denoised, sigmas[i], down)
            x = x + noise_fn(i, x) * s_noise * up
    return x


def lms_coefficient(order, t, i, j):                                      # sampling.py:251-261
    if order - 1 > i:
        raise ValueError(f"Order {order} too high for step {i}")

    def lagrange(tau):
        out = 1.0
        for k in range(order):
            if k != j:
                out *= (tau - t[i - k]) / (t[i - j] - t[i - k])
        return out
    return integrate.quad(lagrange, t[i], t[i + 1], epsrel=1e-4)[0]


def sample_lms(model, x, sigmas, order=4):                                # sampling.py:259-275
    t = sigmas.detach().cpu().numpy()
    history = []                                                          # oldest first
    for i in range(len(sigmas) - 1):
        history.append((x - model(x, sigmas[i] * _ones(x))) / sigmas[i])
        history = history[-order:]
        cur = min(i + 1, order)
        x = x + sum(lms_coefficient(cur, t, i, j) * d for j, d in enumerate(reversed(history)))
    return x


def sample_dpmpp_2s_ancestral(model, x, sigmas, eta=1.0, s_noise=1.0, noise_fn=_default_noise):   # sampling.py:507-538
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * _ones(x))
        down, up = ancestral_step(sigmas[i], sigmas[i + 1], eta)
        if down == 0:
            x = x + (x - denoised) / sigmas[i] * (down - sigmas[i])
        else:
            t, t_next = -sigmas[i].log(), -down.log()
            h = t_next - t
            s = t + 0.5 * h
            x_2 = ((-s).exp() / (-t).exp()) * x - (-h * 0.5).expm1() * denoised
            denoised_2 = model(x_2, (-s).exp() * _ones(x))
            x = ((-t_next).exp() / (-t).exp()) * x - (-h).expm1() * denoised_2
        if sigmas[i + 1] > 0:
            x = x + noise_fn(i, x) * s_noise * up
    return x


def sample_dpmpp_2m(model, x, sigmas):                                    # sampling.py:583-606
    previous = None
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * _ones(x))
        t, t_next = -sigmas[i].log(), -sigmas[i + 1].log()
        h = t_next - t
        target = denoised
        if previous is not None and sigmas[i + 1] != 0:
            r = (t - (-sigmas[i - 1].log())) / h
            target = (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * previous
        x = ((-t_next).exp() / (-t).exp()) * x - (-h).expm1() * target
        previous = denoised
    return x
