"""Checkpoint front end: the reference's two on-disk weight formats -> the modules of this package.

1. OpenAI guided-diffusion ``.pt`` (``diffusion_ffhq_10m.pt``, ``256x256_diffusion_uncond.pt``): a flat ``state_dict`` of
   ``UNetModel`` loaded by ``inner_model.load_state_dict(dist_util.load_state_dict(path, map_location="cpu"))``
   (sample_condition_openai.py:128-132).  ``UNetModel`` here declares the same keys / shapes, so that call works as is;
   ``load_openai_unet`` wraps it together with the factory call.
2. PyTorch-Lightning ``.ckpt`` of the DWT-Var model (``ffhq_dwt.ckpt``) written by train_openai.py:77-135: ``state_dict`` holds
   ``model.*`` (training copy) and ``model_ema.*`` (EMA copy), each an ``OpenAIDenoiserV2``: ``inner_model.<UNet keys>``,
   ``out_cov.weight [6,128,1,1]``, ``out_cov.bias [6]`` and the schedule buffers ``sigmas`` / ``log_sigmas``
   (k_diffusion/external.py:45-46); ``hyper_parameters`` holds ``model_config`` / ``train_config`` (``save_hyperparameters()``,
   train_openai.py:81).  sample_condition_openai_v2.py:117 uses ``OpenAIDenoiser.load_from_checkpoint(path).model_ema``.

Weights are packed into the kernel layouts (bf16 K-major slabs, csrc/pack.cu) lazily by ``UNetModel.engine()`` on first use.
"""
import torch

from guided_diffusion import dist_util


def unwrap_state_dict(obj):
    """A checkpoint object -> its flat name->tensor mapping (Lightning nests it under 'state_dict')."""
    if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        return obj["state_dict"]
    return obj


def split_denoiser_state_dict(sd, copy="model_ema"):
    """Lightning DWT-Var state_dict -> (unet_state_dict, out_cov_weight, out_cov_bias, buffers) of one copy
    (``model_ema`` = what the sampler uses, sample_condition_openai_v2.py:117; ``model`` = the training copy)."""
    pre = copy + "."
    own = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    if not own:
        raise KeyError(f"checkpoint has no '{pre}*' entries (found prefixes: {sorted({k.split('.')[0] for k in sd})})")
    unet = {k[len("inner_model."):]: v for k, v in own.items() if k.startswith("inner_model.")}
    buffers = {k: v for k, v in own.items() if k in ("sigmas", "log_sigmas")}
    extra = set(own) - {"inner_model." + k for k in unet} - set(buffers) - {"out_cov.weight", "out_cov.bias"}
    if extra:
        raise KeyError(f"unexpected entries under '{pre}': {sorted(extra)[:5]}")
    if "out_cov.weight" not in own or "out_cov.bias" not in own:
        raise KeyError(f"'{pre}out_cov.weight/bias' missing: not a DWT-Var (OpenAIDenoiserV2) checkpoint")
    return unet, own["out_cov.weight"], own["out_cov.bias"], buffers


def create_unet(openai_overrides):
    """create_model_and_diffusion with the overrides of ``config['model']['openai']`` (sample_condition_openai.py:115-129)."""
    from condition.diffpir_utils.utils_model import create_argparser
    from guided_diffusion.script_util import args_to_dict, create_model_and_diffusion, model_and_diffusion_defaults
    args = create_argparser(dict(openai_overrides)).parse_args([])
    return create_model_and_diffusion(**args_to_dict(args, model_and_diffusion_defaults().keys()))


def load_openai_unet(path, openai_overrides, device=None):
    """(UNetModel.eval() on ``device``, diffusion) from an OpenAI ``.pt`` (format 1)."""
    model, diffusion = create_unet(openai_overrides)
    model.load_state_dict(unwrap_state_dict(dist_util.load_state_dict(path, map_location="cpu")))
    return model.eval().to(device if device is not None else dist_util.dev()), diffusion


def build_denoiser_v2(model_config, state_dict, copy="model_ema", device=None):
    """An ``OpenAIDenoiserV2`` carrying the ``copy`` weights of a Lightning DWT-Var state_dict (format 2)."""
    from k_diffusion.external import OpenAIDenoiserV2
    device = device if device is not None else dist_util.dev()
    unet_sd, cov_w, cov_b, buffers = split_denoiser_state_dict(state_dict, copy)
    inner, diffusion = create_unet(model_config["openai"])
    inner.load_state_dict(unet_sd)
    inner = inner.eval().to(device)
    den = OpenAIDenoiserV2(inner, diffusion, device=device, ortho_tf_type=model_config.get("ortho_tf_type"))
    with torch.no_grad():
        den.out_cov.weight.copy_(cov_w)
        den.out_cov.bias.copy_(cov_b)
    den = den.to(device)
    if "log_sigmas" in buffers:       # the schedule is a function of the DDPM constants; a mismatch means another diffusion
        ref = buffers["log_sigmas"].to(den.log_sigmas.device, den.log_sigmas.dtype)
        if ref.shape != den.log_sigmas.shape or not torch.allclose(ref, den.log_sigmas, rtol=1e-5, atol=1e-6):
            raise ValueError("checkpoint log_sigmas differ from the schedule rebuilt from model_config['openai']")
    return den.eval()
