"""JSON model config loader (k_diffusion/config.py:11-47): the keys the sample scripts read are
``model.input_size / input_channels / sigma_min / sigma_max / openai / recon_mse / ortho_tf_type`` and
``dataset.location`` (sample_condition_openai.py:112-121).  Training-side factories are out of scope."""
import json

_DEFAULTS = {
    "model": {"sigma_data": 1.0, "patch_size": 1, "dropout_rate": 0.0, "augment_wrapper": True, "augment_prob": 0.0,
              "mapping_cond_dim": 0, "unet_cond_dim": 0, "cross_cond_dim": 0, "cross_attn_depths": None, "skip_stages": 0,
              "has_variance": False, "loss_config": "karras"},
    "dataset": {"type": "imagefolder"},
    "optimizer": {"type": "adamw", "lr": 1e-4, "betas": [0.95, 0.999], "eps": 1e-6, "weight_decay": 1e-3},
    "lr_sched": {"type": "constant"},
    "ema_sched": {"type": "inverse", "power": 0.6667, "max_value": 0.9999},
}


def _merge(base, head):
    out = dict(base)
    for k, v in head.items():
        out[k] = _merge(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else v
    return out


def load_config(file):
    """``file``: an open file object (as in the reference) or a path."""
    if isinstance(file, (str, bytes)):
        with open(file) as f:
            return _merge(_DEFAULTS, json.load(f))
    return _merge(_DEFAULTS, json.load(file))
