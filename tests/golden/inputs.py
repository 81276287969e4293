"""Seeded synthetic inputs shared by the golden generator and the tests (CPU torch generator, fixed seeds)."""
import torch

SIGMA_PROBE = torch.tensor([80.0, 33.3, 10.0, 1.5, 0.3, 0.19, 0.05, 0.01, 0.002, 200.0])

# (operator, guidance, x0_cov_type, sigma, extra ConditionDenoiser kwargs)
GUIDANCE_COMBOS = [
    ("gaussian_blur", "pgdm", "pgdm", 10.0, {}),
    ("gaussian_blur", "pgdm", "pgdm", 0.1, {}),
    ("gaussian_blur", "I", "convert", 10.0, {}),
    ("gaussian_blur", "I", "convert", 0.1, {}),
    ("gaussian_blur", "diffpir", "diffpir", 1.0, {"lambda_": 7.0}),
    ("gaussian_blur", "II", "convert", 0.1, {}),
    ("gaussian_blur", "I", "tmpd", 0.1, {}),
    ("inpainting", "pgdm", "pgdm", 10.0, {}),
    ("inpainting", "I", "convert", 0.1, {}),
    ("super_resolution", "I", "analytic", 10.0, {}),
    ("super_resolution", "I", "analytic", 0.1, {}),
    ("super_resolution", "I", "convert", 0.1, {}),
    ("motion_blur", "dps", "dps", 10.0, {"zeta": 1.0}),
    ("motion_blur", "dps", "dps", 0.1, {"zeta": 1.0}),
    ("super_resolution", "dps", "dps", 1.0, {"zeta": 1.0}),
]

# (tag, operator, guidance, cov, sampler, n_steps, churn)
SAMPLER_RUNS = [
    ("inpaint_pgdm_euler6", "inpainting", "pgdm", "pgdm", "euler", 6, False),
    ("gauss_pgdm_heun4", "gaussian_blur", "pgdm", "pgdm", "heun", 4, False),
    ("gauss_pgdm_heun4_churn", "gaussian_blur", "pgdm", "pgdm", "heun", 4, True),
]


def _g(seed):
    return torch.Generator().manual_seed(seed)


def image(size, batch=1, seed=1):
    """Ground-truth image in [-1, 1] (x0 = rand*2-1)."""
    return torch.rand(batch, 3, size, size, generator=_g(seed)) * 2 - 1


def unet_input(size, batch, seed):
    return torch.randn(batch, 3, size, size, generator=_g(seed))


def unet_seed(shape, seed):
    return torch.randn(*shape, generator=_g(seed))


def xt(size, sigma, seed=21, batch=1):
    """A noisy iterate x_t = x0' + sigma*n."""
    g = _g(seed)
    return (torch.rand(batch, 3, size, size, generator=g) * 2 - 1) * 0.7 + sigma * torch.randn(batch, 3, size, size, generator=g)


def xT(size, seed=3, batch=1, sigma_max=80.0):
    return torch.randn(batch, 3, size, size, generator=_g(seed)) * sigma_max


def theta_map(size, seed=5):
    """A positive per-pixel variance map, like Convert's Eq. 22 output at small sigma."""
    return torch.rand(1, 3, size, size, generator=_g(seed)) * 0.05 + 1e-4


def sub(a):
    """4x4 spatial subsample used for the 256x256 golden outputs (arrays whose last dim is 256 only)."""
    return a[..., ::4, ::4] if a.shape[-1] == 256 else a


# ---- v2 (DWT-Var) denoiser path: 64x64 UNet of width 128 (the reference hard-codes out_cov = Conv2d(128, 6, 1)) ----
V2_SIGMAS = [0.5, 2.0]     # below / above mle_sigma_thres = 1.0 (sample_condition_openai_v2.py default)


def v2_config():
    from oracle import unet_ref
    return unet_ref.UNetConfig(64, 128, 1, "16,8")


def v2_out_cov(seed=9):
    """Random-init out_cov head (k_diffusion/external.py:141): weight [6,128,1,1], bias [6]."""
    g = _g(seed)
    return torch.randn(6, 128, 1, 1, generator=g) * 0.05, torch.randn(6, generator=g) * 0.5 - 1.0


# ---- STSL guidance (condition.py:185-208): (operator, sigma, zeta, eta, num_hutchinson_samples, torch seed of the eps draws) ----
STSL_CASES = [
    ("gaussian_blur", 1.0, 1.0, 2000.0, 2, 77),
    ("inpainting", 0.3, 0.5, 1.0e5, 1, 78),
    ("super_resolution", 3.0, 1.0, 300.0, 2, 79),
]


# ---- the other k_diffusion samplers (make_golden_samplers.py) ----------------------------------------
SAMPLER_N, SAMPLER_SIGMA_MIN, SAMPLER_SIGMA_MAX = 10, 0.05, 20.0
SAMPLER_SHAPE = (2, 3, 16, 16)
SAMPLER_NOISE_SEED = 17


def SAMPLER_MODEL(x, sigma, **kw):
    """Analytic denoiser: posterior mean of N(0.1-shifted, I) data under the Karras forward process, up to the shift."""
    return x / (1 + sigma.view(-1, 1, 1, 1) ** 2) + 0.1


def sampler_start():
    return torch.randn(*SAMPLER_SHAPE, generator=_g(23)) * SAMPLER_SIGMA_MAX


def sampler_noises():
    """What torch.randn_like(x) returns step by step after torch.manual_seed(SAMPLER_NOISE_SEED) (CPU generator)."""
    g = torch.Generator().manual_seed(SAMPLER_NOISE_SEED)
    return [torch.randn(*SAMPLER_SHAPE, generator=g) for _ in range(SAMPLER_N)]


EXTRA_SAMPLER_CASES = [
    ("euler_ancestral", dict(eta=1.0, s_noise=1.0)),
    ("euler_ancestral/eta0.5", dict(eta=0.5, s_noise=1.003)),
    ("dpm_2", dict()),
    ("dpm_2/churn", dict(s_churn=40., s_tmin=0.05, s_tmax=50., s_noise=1.003)),
    ("dpm_2_ancestral", dict(eta=1.0, s_noise=1.0)),
    ("lms", dict(order=4)),
    ("lms/order2", dict(order=2)),
    ("dpmpp_2s_ancestral", dict(eta=1.0, s_noise=1.0)),
    ("dpmpp_2s_ancestral/eta0", dict(eta=0.0)),
    ("dpmpp_2m", dict()),
]
