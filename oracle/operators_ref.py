"""Oracle: measurement operators A, A^T and OTF helpers (test infrastructure only).

Follows condition/measurements.py:86-244 (operators), :247-319 (MaskGenerator),
condition/diffpir_utils/utils_sisr.py:9-61,79-96 (splits, p2o, upsample, downsample, pre_calculate) and
condition/dps_utils/resizer.py:8-198 (antialiased bicubic Resizer).
Kernel data: the Gaussian PSF is regenerated exactly as condition/dps_utils/img_utils.py:276-281 does
(scipy.ndimage.gaussian_filter of a delta; bit-identical to condition/kernels/gaussian_ks61_std3.0.npy, checked
in tests/golden/make_golden.py); the motion PSF and bicubic kernels are the reference's data fixtures.
"""
import os

import numpy as np
import scipy.ndimage
import torch
from torch.fft import fft2, ifft2

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "k-diffusion-inverse-problems_b200", "condition", "kernels", "fixed_kernels.npz")


def gaussian_psf(ks=61, std=3.0):
    n = np.zeros((ks, ks))
    n[ks // 2, ks // 2] = 1
    return torch.Tensor(scipy.ndimage.gaussian_filter(n, sigma=std))      # f64 -> f32, measurements.py:173


def motion_psf():
    return torch.Tensor(np.load(_DATA)["motion_ks61_std0p5"])             # measurements.py:134,159


def bicubic_psf(sf=4):
    return torch.Tensor(np.load(_DATA)[f"bicubic_x{sf if sf < 5 else 4}"].astype(np.float64))  # measurements.py:95-97


def p2o(psf, shape):
    """utils_sisr.py:22-41: zero-pad PSF to shape, roll by -floor(k/2), fft2."""
    otf = torch.zeros(psf.shape[:-2] + tuple(shape)).type_as(psf)
    otf[..., :psf.shape[2], :psf.shape[3]].copy_(psf)
    for axis, axis_size in enumerate(psf.shape[2:]):
        otf = torch.roll(otf, -int(axis_size / 2), dims=axis + 2)
    return torch.fft.fftn(otf, dim=(-2, -1))


def splits(a, sf):
    """utils_sisr.py:9-19."""
    b = torch.stack(torch.chunk(a, sf, dim=2), dim=4)
    return torch.cat(torch.chunk(b, sf, dim=3), dim=4)


def upsample(x, sf):
    """utils_sisr.py:44-52: zero-fill."""
    z = torch.zeros((x.shape[0], x.shape[1], x.shape[2] * sf, x.shape[3] * sf)).type_as(x)
    z[..., 0::sf, 0::sf].copy_(x)
    return z


def downsample(x, sf):
    """utils_sisr.py:55-61."""
    return x[..., 0::sf, 0::sf]


def pre_calculate(x, k, sf):
    """utils_sisr.py:79-96."""
    w, h = x.shape[-2:]
    FB = p2o(k, (w * sf, h * sf))
    FBC = torch.conj(FB)
    F2B = torch.pow(torch.abs(FB), 2)
    FBFy = FBC * torch.fft.fftn(upsample(x, sf), dim=(-2, -1))
    return FB, FBC, F2B, FBFy


# ---- Resizer (resizer.py) ---------------------------------------------------------------------

def _cubic(x):
    a = np.abs(x)
    return ((1.5 * a ** 3 - 2.5 * a ** 2 + 1) * (a <= 1) +
            (-0.5 * a ** 3 + 2.5 * a ** 2 - 4 * a + 2) * ((1 < a) & (a <= 2)))


def resizer_contributions(in_length, out_length, scale, kernel_width=4.0, antialiasing=True):
    """resizer.py:104-168 for the cubic kernel.  Returns (weights [out, taps] f64, field_of_view [out, taps] int)."""
    fixed_kernel = (lambda arg: scale * _cubic(scale * arg)) if antialiasing else _cubic
    kernel_width *= 1.0 / scale if antialiasing else 1.0
    out_coordinates = np.arange(1, out_length + 1)
    shifted = out_coordinates - (out_length - in_length * scale) / 2
    match = shifted / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(match - kernel_width / 2)
    expanded = np.ceil(kernel_width) + 2
    fov = np.squeeze(np.int16(np.expand_dims(left, axis=1) + np.arange(expanded) - 1))
    weights = fixed_kernel(1.0 * np.expand_dims(match, axis=1) - fov - 1)
    s = np.sum(weights, axis=1)
    s[s == 0] = 1.0
    weights = 1.0 * weights / np.expand_dims(s, axis=1)
    mirror = np.uint(np.concatenate((np.arange(in_length), np.arange(in_length - 1, -1, step=-1))))
    fov = mirror[np.mod(fov, mirror.shape[0])]
    nz = np.nonzero(np.any(weights, axis=0))
    return np.squeeze(weights[:, nz]), np.squeeze(fov[:, nz])


class Resizer:
    """resizer.py:8-74 with scale_factor = 1/sf on the last two dims of in_shape [1,3,H,W]."""

    def __init__(self, in_shape, scale):
        self.tabs = []
        out_shape = np.uint(np.ceil(np.array(in_shape) * np.array([1, 1, scale, scale])))
        # sorted_dims = argsort(scale_factor) restricted to != 1 -> [2, 3] for equal scales (stable argsort)
        sf = [1, 1, scale, scale]
        for dim in [int(d) for d in np.argsort(np.array(sf)) if sf[d] != 1]:
            w, fov = resizer_contributions(in_shape[dim], int(out_shape[dim]), scale)
            self.tabs.append((dim, torch.tensor(w.T, dtype=torch.float32),
                              torch.tensor(fov.T.astype(np.int32), dtype=torch.long)))

    def __call__(self, x):
        for dim, w, fov in self.tabs:
            x = torch.transpose(x, dim, 0)
            x = torch.sum(x[fov] * w.reshape(list(w.shape) + [1] * 3), dim=0)
            x = torch.transpose(x, dim, 0)
        return x


# ---- operators (measurements.py) ---------------------------------------------------------------

class BlurOperator:
    """measurements.py:125-199 (motion_blur / gaussian_blur): circular convolution via the OTF."""

    def __init__(self, name, sigma_s, kernel_size=61, intensity=3.0, in_shape=(1, 3, 256, 256)):
        self.name = name
        self.kernel = gaussian_psf(kernel_size, intensity) if name == "gaussian_blur" else motion_psf()
        self.kernel_size = kernel_size
        self.sigma_s = torch.Tensor([sigma_s])
        self.in_shape = in_shape

    def get_kernel(self):
        return self.kernel.view(1, 1, self.kernel_size, self.kernel_size)

    def forward(self, data, flatten=False, noiseless=False, noise=None):
        FB, FBC, F2B, _ = pre_calculate(data, self.get_kernel(), 1)
        y = ifft2(FB * fft2(data)).real
        if not noiseless:
            y = y + self.sigma_s * (torch.randn_like(y) if noise is None else noise)
        self.pre_calculated = (FB, FBC, F2B, FBC * fft2(y))
        return (y, y.reshape(y.shape[0], -1)) if flatten else y

    def transpose(self, y, flatten=False):
        if flatten:
            y = y.reshape(y.shape[0], *self.in_shape[-3:])
        FB, FBC, F2B, _ = pre_calculate(y, self.get_kernel(), 1)
        return ifft2(FBC * fft2(y)).real


class SuperResolutionOperator:
    """measurements.py:86-122: y = Resizer(x) (+noise); pre_calculated from the bicubic FFT model."""
    name = "super_resolution"

    def __init__(self, sigma_s, scale_factor=4, in_shape=(1, 3, 256, 256)):
        self.down_sample = Resizer(in_shape, 1 / scale_factor)
        self.scale_factor = scale_factor
        self.sigma_s = torch.Tensor([sigma_s])
        self.kernel = bicubic_psf(scale_factor)
        self.in_shape = in_shape
        self.out_shape = (1, 3, int(in_shape[-2] / scale_factor), int(in_shape[-1] / scale_factor))

    def get_kernel(self):
        return self.kernel.view(1, 1, *self.kernel.shape)

    def forward(self, data, flatten=False, noiseless=False, noise=None):
        y = self.down_sample(data)
        if not noiseless:
            y = y + self.sigma_s * (torch.randn_like(y) if noise is None else noise)
        self.pre_calculated = pre_calculate(y, self.get_kernel(), self.scale_factor)
        return (y, y.reshape(y.shape[0], -1)) if flatten else y

    def transpose(self, y, flatten=False):
        if flatten:
            y = y.reshape(y.shape[0], *self.out_shape[-3:])
        FB, FBC, F2B, FBFy = pre_calculate(y, self.get_kernel(), self.scale_factor)
        return ifft2(FBFy).real


def box_mask(image_size=256, mask_len=128, margin=(16, 16), batch=1):
    """measurements.py:275-284,300-319 with mask_len_range=(l, l+1): centred box of zeros."""
    h = w = mask_len
    maxt = image_size - margin[0] - h
    maxl = image_size - margin[1] - w
    t = (margin[0] + maxt) // 2
    l = (margin[1] + maxl) // 2
    mask = torch.ones([batch, 3, image_size, image_size])
    mask[..., t:t + h, l:l + w] = 0
    return mask


def random_mask(image_size, prob, rng: np.random.RandomState):
    """measurements.py:286-298: same per-pixel mask on the 3 channels (uses numpy RNG)."""
    total = image_size ** 2
    mask_vec = torch.ones([1, total])
    samples = rng.choice(total, int(total * prob), replace=False)
    mask_vec[:, samples] = 0
    return mask_vec.view(1, image_size, image_size).repeat(3, 1, 1)[None]


class InpaintingOperator:
    """measurements.py:202-244."""
    name = "inpainting"

    def __init__(self, sigma_s, mask):
        self.sigma_s = torch.Tensor([sigma_s])
        self.mask = mask
        self.in_shape = (1, 3, mask.shape[-2], mask.shape[-1])

    def forward(self, data, flatten=False, noiseless=False, noise=None):
        y = data.clone()
        if not noiseless:
            y = y + self.sigma_s * (torch.randn_like(y) if noise is None else noise)
        y = y * self.mask
        if flatten:
            idx = torch.where(self.mask > 0)
            return y, y[..., idx[-3], idx[-2], idx[-1]]
        return y

    def transpose(self, data, flatten=False):
        y = data.clone()
        if flatten:
            idx = torch.where(self.mask > 0)
            x = torch.zeros(y.shape[0], *self.in_shape[-3:])
            x[..., idx[-3], idx[-2], idx[-1]] = y
            return x
        return y
