"""Antialiased bicubic ``Resizer`` (condition/dps_utils/resizer.py:8-198) = the super-resolution forward operator.

The per-output weights / field-of-view tables are computed on the host in numpy exactly as the reference does at
construction (resizer.py:104-168); the resize itself (separable weighted gather, H pass then W pass) and its adjoint
run in the libkdip kernels of the super-resolution operator handle."""
import numpy as np
import torch
from torch import nn


def cubic(x):
    absx = np.abs(x)
    absx2, absx3 = absx ** 2, absx ** 3
    return ((1.5 * absx3 - 2.5 * absx2 + 1) * (absx <= 1) +
            (-0.5 * absx3 + 2.5 * absx2 - 4 * absx + 2) * ((1 < absx) & (absx <= 2)))


def contributions(in_length, out_length, scale, kernel=cubic, kernel_width=4.0, antialiasing=True):
    """resizer.py:104-168 -> (weights [out, taps] float64, field_of_view [out, taps] int)."""
    fixed_kernel = (lambda arg: scale * kernel(scale * arg)) if antialiasing else kernel
    kernel_width *= 1.0 / scale if antialiasing else 1.0
    out_coordinates = np.arange(1, out_length + 1)
    shifted_out_coordinates = out_coordinates - (out_length - in_length * scale) / 2
    match_coordinates = shifted_out_coordinates / scale + 0.5 * (1 - 1 / scale)
    left_boundary = np.floor(match_coordinates - kernel_width / 2)
    expanded_kernel_width = np.ceil(kernel_width) + 2
    field_of_view = np.squeeze(np.int16(np.expand_dims(left_boundary, axis=1) + np.arange(expanded_kernel_width) - 1))
    weights = fixed_kernel(1.0 * np.expand_dims(match_coordinates, axis=1) - field_of_view - 1)
    sum_weights = np.sum(weights, axis=1)
    sum_weights[sum_weights == 0] = 1.0
    weights = 1.0 * weights / np.expand_dims(sum_weights, axis=1)
    mirror = np.uint(np.concatenate((np.arange(in_length), np.arange(in_length - 1, -1, step=-1))))
    field_of_view = mirror[np.mod(field_of_view, mirror.shape[0])]
    non_zero_out_pixels = np.nonzero(np.any(weights, axis=0))
    return np.squeeze(weights[:, non_zero_out_pixels]), np.squeeze(field_of_view[:, non_zero_out_pixels])


class Resizer(nn.Module):
    """Square NCHW images, equal scale on H and W, cubic kernel (what the path constructs, measurements.py:91)."""

    def __init__(self, in_shape, scale_factor=None, output_shape=None, kernel=None, antialiasing=True):
        super().__init__()
        if kernel not in (None, "cubic"):
            raise NotImplementedError("kdip Resizer implements the cubic kernel only")
        if scale_factor is None or not np.isscalar(scale_factor) or in_shape[-1] != in_shape[-2]:
            raise NotImplementedError("kdip Resizer needs a scalar scale_factor and square images")
        S = int(in_shape[-1])
        out = int(np.ceil(S * scale_factor))
        antialiasing = bool(antialiasing) and scale_factor < 1
        w, fov = contributions(S, out, scale_factor, cubic, 4.0, antialiasing)
        self.in_size, self.out_size, self.scale_factor = S, out, scale_factor
        self.tables = (np.asarray(w, dtype=np.float32), np.asarray(fov, dtype=np.int32))   # fp32 like torch.tensor(weights.T, float32)
        self._handle = None

    def handle(self, device):
        if self._handle is None or self._handle.device != torch.device(device):
            from kdip.ops import OperatorHandle
            delta = np.ones((1, 1), dtype=np.float32)
            self._handle = OperatorHandle("super_resolution", self.in_size, 0.0, device, psf=delta,
                                          sf=self.in_size // self.out_size, resizer=self.tables)
        return self._handle

    def forward(self, in_tensor):
        return self.handle(in_tensor.device).forward(in_tensor, None)
