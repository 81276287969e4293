"""k_diffusion — B200-native mirror of the reference's sampler / denoiser-wrapper interface (hot path only):
``sampling.get_sigmas_karras / sample_euler / sample_heun``, ``external.OpenAIDenoiser[V2]``,
``evaluation.compute_features``, ``utils.append_dims``, ``config.load_config``."""
from . import config, evaluation, external, sampling, utils  # noqa: F401
