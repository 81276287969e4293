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
