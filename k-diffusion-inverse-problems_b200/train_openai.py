"""Inference-side stand-in for the reference's ``train_openai.OpenAIDenoiser`` LightningModule (train_openai.py:77-135).

sample_condition_openai_v2.py:117 does ``OpenAIDenoiser.load_from_checkpoint(args.checkpoint).model_ema.eval()``; this class
provides exactly that: ``load_from_checkpoint`` reads the Lightning ``.ckpt`` (``hyper_parameters`` -> model / train config,
``state_dict`` -> ``model.*`` and ``model_ema.*``) and exposes ``.model`` and ``.model_ema`` as ``k_diffusion.external
.OpenAIDenoiserV2`` instances running on libkdip.  Training (``training_step``, EMA update, optimiser) is out of scope
 and raises.
"""
from guided_diffusion import dist_util
from kdip.checkpoint import build_denoiser_v2, unwrap_state_dict


class OpenAIDenoiser:
    def __init__(self, model_config, train_config, state_dict=None, device=None):
        self.model_config, self.train_config = model_config, train_config
        self.hparams = {"model_config": model_config, "train_config": train_config}
        if state_dict is None:
            raise NotImplementedError("kdip train_openai.OpenAIDenoiser is inference-only: build it with load_from_checkpoint")
        self.model = build_denoiser_v2(model_config, state_dict, "model", device)
        self.model_ema = build_denoiser_v2(model_config, state_dict, "model_ema", device)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **overrides):
        ckpt = dist_util.load_state_dict(checkpoint_path, map_location="cpu")
        if "hyper_parameters" not in ckpt:
            raise KeyError("not a Lightning checkpoint of train_openai.OpenAIDenoiser: 'hyper_parameters' missing")
        hp = dict(ckpt["hyper_parameters"])
        hp.update(overrides)
        return cls(hp["model_config"], hp["train_config"], state_dict=unwrap_state_dict(ckpt), device=map_location)

    def eval(self):
        return self

    def training_step(self, *a, **k):
        raise NotImplementedError("training is outside the kdip hot path (SURVEY.md §2)")
