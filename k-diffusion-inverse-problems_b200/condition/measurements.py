"""Measurement operators A, A^T of the guided-sampling path (condition/measurements.py:24-319), device-resident.

Same registry (``register_operator`` / ``get_operator``), constructor keyword arguments (= the keys of
configs/*_config.yaml + ``device``) and attributes (``name``, ``device``, ``sigma_s``, ``in_shape``, ``mask``,
``scale_factor``, ``out_shape``, ``pre_calculated``) as the reference.  The arithmetic runs in libkdip through an
``OperatorHandle`` (batched real FFTs for the circular blur model, separable gather for the Resizer, masked
loads for inpainting).  Operators are batch-capable (B >= 1); the reference's B = 1 is a special case.
Kernel data fixtures (motion PSF, bicubic kernels) are the reference's, stored in kernels/fixed_kernels.npz; the
Gaussian PSF is regenerated exactly as condition/dps_utils/img_utils.py:276-281 does.
Out of scope (no mat solver in the reference, condition/condition.py:317-401): 'noise', 'colorization',
'phase_retrieval', 'nonlinear_blur' operators and the __NOISE__ classes.
"""
import os
from abc import ABC, abstractmethod

import numpy as np
import scipy.ndimage
import torch

from kdip import ops
from kdip.ops import OperatorHandle

from .dps_utils.resizer import Resizer

_KERNELS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernels", "fixed_kernels.npz")

__OPERATOR__ = {}


def register_operator(name: str):
    def wrapper(cls):
        if __OPERATOR__.get(name, None):
            raise NameError(f"Name {name} is already registered!")
        cls.name = name
        __OPERATOR__[name] = cls
        return cls
    return wrapper


def get_operator(name: str, **kwargs):
    if __OPERATOR__.get(name, None) is None:
        raise NameError(f"Name {name} is not defined.")
    return __OPERATOR__[name](**kwargs)


class _ForwardFn(torch.autograd.Function):
    """operator.forward as an autograd node: the backward is the adjoint of the SAME forward (kdip_op_forward_adjoint), which is
    what torch.autograd derives in the reference (auto_transpose, measurements.py:48-52; DPS, condition.py:144-146)."""

    @staticmethod
    def forward(ctx, x, handle, noise):
        ctx.handle = handle
        return handle.forward(x, noise)

    @staticmethod
    def backward(ctx, g):
        return ctx.handle.forward_adjoint(g.contiguous()), None, None


def _apply_forward(handle, data, noise):
    if torch.is_grad_enabled() and data.requires_grad:
        return _ForwardFn.apply(data, handle, noise)
    return handle.forward(data, noise)


class LinearOperator(ABC):
    @abstractmethod
    def forward(self, data, flatten=False, noiseless=False):
        raise NotImplementedError("The class {} requires a forward function!".format(self.__class__.__name__))

    def auto_transpose(self, y, flatten=False):
        """measurements.py:48-52: the VJP of ``forward`` (noiseless) at a random point - for a linear operator its exact adjoint.
        Works for the built-in operators (their forward is an autograd node over the CUDA adjoint kernels) and for any
        user-registered operator whose forward is written in torch."""
        with torch.enable_grad():
            input = torch.randn(y.shape[0], *self.in_shape[-3:]).to(self.device).requires_grad_()
            out = self.forward(input, flatten=flatten, noiseless=True)
            if flatten:
                out = out[1]
            res = torch.autograd.grad((y * out).sum(), input, retain_graph=True)[0]
        return res

    def _noise(self, shape_like, noiseless):
        return None if noiseless else torch.randn_like(shape_like)


class _SpectralOperator(LinearOperator):
    """Shared by blur / SR: ``pre_calculated`` = (FB, FBC, F2B, FBFy) of utils_sisr.pre_calculate, built on demand."""

    handle: OperatorHandle

    def _set_measurement(self, y):
        self._last_y = y

    @property
    def pre_calculated(self):
        FB = self.handle.otf()[None, None]
        FBC = torch.conj(FB)
        F2B = FB.real ** 2 + FB.imag ** 2
        y = getattr(self, "_last_y", None)
        FBFy = None
        if y is not None:
            sf = getattr(self, "scale_factor", 1)
            from .diffpir_utils.utils_sisr import upsample
            FBFy = FBC * self.handle.fft2(upsample(y, sf) if sf > 1 else y)     # utils_sisr.py:91-95, on libkdip's FFT kernels
        return FB, FBC, F2B, FBFy


def gaussian_psf(kernel_size, std):
    n = np.zeros((kernel_size, kernel_size))
    n[kernel_size // 2, kernel_size // 2] = 1
    return scipy.ndimage.gaussian_filter(n, sigma=std)


class _BlurOperator(_SpectralOperator):
    def _init(self, in_shape, kernel_size, kernel, sigma_s, device):
        self.device = device
        self.kernel_size = kernel_size
        self.kernel = torch.Tensor(kernel)                                  # f64 -> f32, measurements.py:135,173
        self.sigma_s = torch.Tensor([sigma_s]).to(device)
        self.in_shape = in_shape
        self.handle = OperatorHandle(self.name, in_shape[-1], sigma_s, device, psf=self.kernel.numpy())

    def forward(self, data, flatten=False, noiseless=False):
        # y = A x + sigma_s * randn_like(y): the noise add is fused into the inverse-FFT epilogue
        y = _apply_forward(self.handle, data, None if noiseless else torch.randn_like(data, dtype=torch.float32))
        self._set_measurement(y.detach())
        if flatten:
            return y, y.reshape(y.shape[0], -1)
        return y

    def transpose(self, y, flatten=False):
        if flatten:
            y = y.reshape(y.shape[0], *self.in_shape[-3:])
        return self.handle.transpose(y)

    def get_kernel(self):
        return self.kernel.view(1, 1, self.kernel_size, self.kernel_size)


@register_operator(name='motion_blur')
class MotionBlurOperator(_BlurOperator):
    """measurements.py:125-160: the fixed motion PSF fixture (motion_ks61_std0.5.npy), circular blur via the OTF."""

    def __init__(self, in_shape, kernel_size, intensity, sigma_s, device):
        kernel = np.load(_KERNELS)["motion_ks61_std0p5"]
        self._init(in_shape, kernel_size, kernel, sigma_s, device)


@register_operator(name='gaussian_blur')
class GaussialBlurOperator(_BlurOperator):
    """measurements.py:163-199."""

    def __init__(self, in_shape, kernel_size, intensity, sigma_s, device):
        self._init(in_shape, kernel_size, gaussian_psf(kernel_size, intensity), sigma_s, device)


@register_operator(name='super_resolution')
class SuperResolutionOperator(_SpectralOperator):
    """measurements.py:86-122: y = Resizer(x) + noise, while ``transpose`` / ``pre_calculated`` / the mat solver use the
    bicubic-kernel FFT model — the reference's deliberate model mismatch is preserved."""

    def __init__(self, in_shape, scale_factor, sigma_s, device):
        self.device = device
        self.down_sample = Resizer(in_shape, 1 / scale_factor)
        self.scale_factor = scale_factor
        self.sigma_s = torch.Tensor([sigma_s]).to(device)
        k_index = scale_factor - 2 if scale_factor < 5 else 2
        self.kernel = torch.Tensor(np.load(_KERNELS)[f"bicubic_x{k_index + 2}"].astype(np.float64))
        self.in_shape = in_shape
        out_shape = tuple(int(s / scale_factor) for s in in_shape[-2:])
        self.out_shape = (1, 3, *out_shape)
        self.handle = OperatorHandle("super_resolution", in_shape[-1], sigma_s, device, psf=self.kernel.numpy(),
                                     sf=scale_factor, resizer=self.down_sample.tables)

    def forward(self, data, flatten=False, noiseless=False):
        if noiseless:
            y = _apply_forward(self.handle, data, None)
        else:
            B = data.shape[0]
            noise = torch.randn(B, *self.out_shape[-3:], device=data.device, dtype=torch.float32)
            y = _apply_forward(self.handle, data, noise)
        self._set_measurement(y.detach())
        if flatten:
            return y, y.reshape(y.shape[0], -1)
        return y

    def transpose(self, y, flatten=False):
        if flatten:
            y = y.reshape(y.shape[0], *self.out_shape[-3:])
        return self.handle.transpose(y)

    def get_kernel(self):
        return self.kernel.view(1, 1, *self.kernel.shape)


@register_operator(name='inpainting')
class InpaintingOperator(LinearOperator):
    """measurements.py:202-244: y = mask * (x + sigma_s n); flatten gathers the kept pixels (bit-exact index ops)."""

    def __init__(self, device, sigma_s, mask_opt):
        self.device = device
        self.sigma_s = torch.Tensor([sigma_s]).to(device)
        self.in_shape = (1, 3, mask_opt['image_size'], mask_opt['image_size'])
        self.mask = self.generate_mask(mask_opt)
        self.handle = OperatorHandle("inpainting", self.in_shape[-1], sigma_s, device, mask=self.mask[0].cpu().numpy())
        self._idx = torch.nonzero(self.mask[0].flatten() > 0).flatten().to(torch.int32).contiguous()

    def forward(self, data: torch.Tensor, flatten=False, noiseless=False):
        y = _apply_forward(self.handle, data, None if noiseless else torch.randn_like(data))
        if flatten:
            # the gather is an index op: torch's own indexing keeps it differentiable when a graph is being recorded
            yf = y.flatten(1)[:, self._idx.long()] if y.requires_grad else ops.gather(y, self._idx)
            return y, yf
        return y

    def transpose(self, data, flatten=False):
        if flatten:
            return ops.scatter(data, self._idx, self.in_shape[-3:])
        return data.clone()

    def generate_mask(self, mask_opt):
        mask_generator = MaskGenerator(**mask_opt)
        img = torch.randn(*self.in_shape).to(self.device)
        return mask_generator(img)


class MaskGenerator:
    """measurements.py:247-319.  Consumes numpy's global RNG exactly like the reference ('box': two randint draws;
    'random': one uniform + one choice)."""

    def __init__(self, mask_type, mask_len_range=None, mask_prob_range=None, image_size=256, margin=(16, 16)):
        assert mask_type in ['box', 'random', 'both', 'extreme']
        self.mask_type = mask_type
        self.mask_len_range = mask_len_range
        self.mask_prob_range = mask_prob_range
        self.image_size = image_size
        self.margin = margin

    def __call__(self, img):
        if self.mask_type == 'random':
            return self._retrieve_random(img)
        if self.mask_type == 'box':
            return self._retrieve_box(img)[0]
        if self.mask_type == 'extreme':
            return 1. - self._retrieve_box(img)[0]

    def _retrieve_box(self, img):
        lo, hi = int(self.mask_len_range[0]), int(self.mask_len_range[1])
        mask_h = np.random.randint(lo, hi)
        mask_w = np.random.randint(lo, hi)
        return self._random_sq_bbox(img, (mask_h, mask_w), self.image_size, self.margin)

    def _retrieve_random(self, img):
        total = self.image_size ** 2
        prob = np.random.uniform(*self.mask_prob_range)
        mask_vec = torch.ones([1, total])
        samples = np.random.choice(total, int(total * prob), replace=False)
        mask_vec[:, samples] = 0
        mask_b = mask_vec.view(1, self.image_size, self.image_size).repeat(3, 1, 1)
        mask = torch.ones_like(img, device=img.device)
        mask[:, ...] = mask_b
        return mask

    def _random_sq_bbox(self, img, mask_shape, image_size=256, margin=(16, 16)):
        """Centred box (the reference computes the box position deterministically, measurements.py:310-313)."""
        B, C, H, W = img.shape
        h, w = mask_shape
        t = (margin[0] + (image_size - margin[0] - h)) // 2
        l = (margin[1] + (image_size - margin[1] - w)) // 2
        mask = torch.ones([B, C, H, W], device=img.device)
        mask[..., t:t + h, l:l + w] = 0
        return mask, t, t + h, l, l + w
