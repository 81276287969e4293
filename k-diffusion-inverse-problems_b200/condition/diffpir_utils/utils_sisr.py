"""FFT helpers of condition/diffpir_utils/utils_sisr.py:9-61,79-96 used by the reference's operators for their
``pre_calculated`` attribute.  On the guided-sampling path itself these quantities live inside the libkdip operator
handle (OTF, |OTF|^2 alias mean); the functions here serve user code that reads ``operator.pre_calculated`` or calls
the helpers directly.  ``splits`` / ``upsample`` / ``downsample`` are pure index ops."""
import torch


def splits(a, sf):
    """[..., W, H] -> [..., W/sf, H/sf, sf^2]: the sf x sf aliases of each low-res bin (utils_sisr.py:9-19)."""
    b = torch.stack(torch.chunk(a, sf, dim=2), dim=4)
    return torch.cat(torch.chunk(b, sf, dim=3), dim=4)


def p2o(psf, shape):
    """PSF [..., k, k] -> OTF [..., *shape] (utils_sisr.py:22-41): zero-pad, roll by -floor(k/2), fft2."""
    otf = torch.zeros(psf.shape[:-2] + tuple(shape)).type_as(psf)
    otf[..., :psf.shape[2], :psf.shape[3]].copy_(psf)
    for axis, axis_size in enumerate(psf.shape[2:]):
        otf = torch.roll(otf, -int(axis_size / 2), dims=axis + 2)
    return torch.fft.fftn(otf, dim=(-2, -1))


def upsample(x, sf=3):
    """zero-filling upsample (utils_sisr.py:44-52)."""
    z = torch.zeros((x.shape[0], x.shape[1], x.shape[2] * sf, x.shape[3] * sf)).type_as(x)
    z[..., 0::sf, 0::sf].copy_(x)
    return z


def downsample(x, sf=3):
    """stride-sf subsample (utils_sisr.py:55-61)."""
    return x[..., 0::sf, 0::sf]


def pre_calculate(x, k, sf):
    """(FB, FBC, F2B, FBFy) of utils_sisr.py:79-96."""
    w, h = x.shape[-2:]
    FB = p2o(k, (w * sf, h * sf))
    FBC = torch.conj(FB)
    F2B = torch.pow(torch.abs(FB), 2)
    FBFy = FBC * torch.fft.fftn(upsample(x, sf=sf), dim=(-2, -1))
    return FB, FBC, F2B, FBFy
