"""Orthogonal transforms of the lazy posterior covariance (condition/utils.py:50-139), on the device.

The reference round-trips through the host for every call (scipy.fft.dctn / pywt on numpy); here ``OrthoTransform``
launches the libkdip kernels (kdip_ortho): Haar level-3 DWT packed like ``pywt.coeffs_to_array`` and the orthonormal
DCT-II over (C, H, W).  Quirk kept: the reference's DCT has no ``axes`` argument, so it transforms the 3-channel axis
too (and the batch axis, which is the identity at the reference's B = 1); images stay independent here for B > 1.
"""
from __future__ import annotations

import copy

import torch

from kdip import ops

__OT__ = dict()


def register_ot(name: str):
    def wrapper(cls):
        __OT__[name] = cls
        return cls
    return wrapper


class OrthoTransform:
    """condition/utils.py:50-67: ``ot(x)`` = W^T x, ``ot.inv(x)`` = W x; ``ortho_tf_type`` in {None, 'dct', 'dwt'}."""

    def __init__(self, ortho_tf_type=None):
        self.ortho_tf_type = ortho_tf_type
        if ortho_tf_type is not None:
            if ortho_tf_type not in __OT__:
                raise ValueError(f"Invalid orthogonal transform type: '{ortho_tf_type}'.")
            self.ot = __OT__[ortho_tf_type]()
            self.iot = self.ot.inv()

    def __call__(self, x: torch.Tensor):
        return x if self.ortho_tf_type is None else self.ot(x)

    def inv(self, x: torch.Tensor):
        return x if self.ortho_tf_type is None else self.iot(x)


class OrthoLinearFunction:
    """condition/utils.py:13-47,80-86: a linear map with ``forward`` / ``transpose``; ``inv()`` swaps them."""

    def __call__(self, x):
        return self.forward(x)

    def inv(self) -> OrthoLinearFunction:
        out = copy.copy(self)
        out.forward, out.transpose = self.transpose, self.forward
        return out


@register_ot('dct')
class DiscreteCosineTransform(OrthoLinearFunction):
    """condition/utils.py:89-103."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.ortho('dct', x, inverse=False)

    def transpose(self, x: torch.Tensor) -> torch.Tensor:
        return ops.ortho('dct', x, inverse=True)


@register_ot('dwt')
class DiscreteWaveletTransform(OrthoLinearFunction):
    """condition/utils.py:107-139 (haar, level 3 — the only configuration the path uses)."""

    def __init__(self, level=3, wavelet='haar') -> None:
        if level != 3 or wavelet != 'haar':
            raise NotImplementedError("kdip DWT implements the path's configuration only: haar, level 3")
        self.level, self.wavelet = level, wavelet

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.ortho('dwt', x, inverse=False)

    def transpose(self, x: torch.Tensor) -> torch.Tensor:
        return ops.ortho('dwt', x, inverse=True)
