"""Oracle: orthogonal transforms W^T (forward) / W (inverse) of condition/utils.py:50-139 (test infrastructure only).

DCT: scipy.fft.dctn/idctn(norm='ortho') with NO axes argument -> all four axes of [B,3,H,W] are transformed
(condition/utils.py:91-103).  For B=1 the batch axis is the identity; the 3-channel axis IS transformed.
The oracle applies it per image (B=1 semantics of the reference, condition/condition.py:84).

DWT: pywt.wavedec2(x,'haar',level=3,axes=(-2,-1)) + pywt.coeffs_to_array (condition/utils.py:116-132).
PyWavelets is neither vendored under /root/reference nor installed here (and unpinned in environment.yml), so this
is a restatement of its published conventions — PARITY UNPINNED for the packed slot layout:
  cA[k] = (x[2k] + x[2k+1])/sqrt2, cD[k] = (x[2k] - x[2k+1])/sqrt2 per axis;
  level block 'da' (detail on axis -2 / rows, approx on axis -1) = pywt's cH, 'ad' = cV, 'dd' = cD;
  coeffs_to_array puts cA_3 at [0:s,0:s]; for each level the key letters address the axes in order:
  'd' on an axis -> slice [s:2s] on that axis, 'a' -> [0:s]   (so cH='da' -> rows [s:2s], cols [0:s]).
The layout is one table (`DWT_SLOT`) so it can be flipped once checked against a pywt install.
"""
import os

import numpy as np
import scipy.fft
import torch

LEVEL = 3
# block name -> (row_is_detail, col_is_detail).  "code" = the slicing rule of pywt.coeffs_to_array as restated above;
# "diagram" = the placement drawn in the PyWavelets docstring ('da' top-right).  Same switch as csrc/transforms.cu
# (KDIP_DWT_LAYOUT); tests/golden/make_golden_pywt.py settles it on a machine that has PyWavelets.
DWT_LAYOUTS = {"code": {"da": (1, 0), "ad": (0, 1), "dd": (1, 1)},
               "diagram": {"da": (0, 1), "ad": (1, 0), "dd": (1, 1)}}
DWT_SLOT = DWT_LAYOUTS[os.environ.get("KDIP_DWT_LAYOUT", "code")]


def dct_forward(x):
    """condition/utils.py:91-96, applied per image."""
    out = [scipy.fft.dctn(xi[None].detach().numpy(), norm="ortho")[0] for xi in x]
    return torch.Tensor(np.stack(out))


def dct_inverse(x):
    """condition/utils.py:98-103, applied per image."""
    out = [scipy.fft.idctn(xi[None].detach().numpy(), norm="ortho")[0] for xi in x]
    return torch.Tensor(np.stack(out))


def _haar_split(a, axis):
    a = np.moveaxis(a, axis, -1)
    lo = (a[..., 0::2] + a[..., 1::2]) / np.sqrt(2.0)
    hi = (a[..., 0::2] - a[..., 1::2]) / np.sqrt(2.0)
    return np.moveaxis(lo, -1, axis), np.moveaxis(hi, -1, axis)


def _haar_merge(lo, hi, axis):
    lo = np.moveaxis(lo, axis, -1)
    hi = np.moveaxis(hi, axis, -1)
    out = np.empty(lo.shape[:-1] + (lo.shape[-1] * 2,), dtype=lo.dtype)
    out[..., 0::2] = (lo + hi) / np.sqrt(2.0)
    out[..., 1::2] = (lo - hi) / np.sqrt(2.0)
    return np.moveaxis(out, -1, axis)


def dwt_forward(x, slot=None):
    """condition/utils.py:116-123: level-3 Haar, packed like pywt.coeffs_to_array. float32 in/out
    (pywt computes float32 input in float32)."""
    DWT_SLOT = slot or globals()["DWT_SLOT"]
    a = x.detach().numpy().astype(np.float32)
    out = np.empty_like(a)
    cur = a
    for _ in range(LEVEL):
        lo_r, hi_r = _haar_split(cur, -2)                  # rows (axis -2) first letter
        aa, ad = _haar_split(lo_r, -1)
        da, dd = _haar_split(hi_r, -1)
        s = aa.shape[-1]
        for name, blk in (("da", da), ("ad", ad), ("dd", dd)):
            r, c = DWT_SLOT[name]
            out[..., r * s:(r + 1) * s, c * s:(c + 1) * s] = blk
        cur = aa
    out[..., :cur.shape[-2], :cur.shape[-1]] = cur
    return torch.tensor(out)


def dwt_inverse(x, slot=None):
    """condition/utils.py:125-132: array_to_coeffs + waverec2."""
    DWT_SLOT = slot or globals()["DWT_SLOT"]
    a = x.detach().numpy().astype(np.float32)
    s = a.shape[-1] >> LEVEL
    cur = a[..., :s, :s]
    for _ in range(LEVEL):
        blk = {}
        for name in ("da", "ad", "dd"):
            r, c = DWT_SLOT[name]
            blk[name] = a[..., r * s:(r + 1) * s, c * s:(c + 1) * s]
        lo_r = _haar_merge(cur, blk["ad"], -1)
        hi_r = _haar_merge(blk["da"], blk["dd"], -1)
        cur = _haar_merge(lo_r, hi_r, -2)
        s *= 2
    return torch.tensor(np.ascontiguousarray(cur))


class OrthoTransform:
    """condition/utils.py:50-67."""

    def __init__(self, ortho_tf_type=None):
        self.ortho_tf_type = ortho_tf_type
        self._f, self._i = {None: (lambda x: x, lambda x: x), "dct": (dct_forward, dct_inverse),
                            "dwt": (dwt_forward, dwt_inverse)}[ortho_tf_type]

    def __call__(self, x):
        return self._f(x)

    def inv(self, x):
        return self._i(x)
