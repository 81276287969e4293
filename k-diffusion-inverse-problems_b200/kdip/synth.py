"""Synthetic weights for benchmarks and smoke runs (no checkpoints are available offline).

Every tensor is non-zero — in particular the reference's ``zero_module`` tensors (unet.py:210,295,617), which would
make every ResBlock / attention / head output exactly zero — so no branch of the network is vacuous."""
import math

import torch


def synthetic_state_dict(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        shp = tuple(v.shape)
        is_norm = (".in_layers.0." in k or ".out_layers.0." in k or ".norm." in k or k.startswith("out.0."))
        if is_norm:
            sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g)) if k.endswith("weight") else 0.1 * torch.randn(shp, generator=g)
        elif len(shp) > 1:
            fan_in = math.prod(shp[1:])
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * (math.sqrt(3.0) * 0.8 / math.sqrt(fan_in))
        else:
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.1
    return sd
