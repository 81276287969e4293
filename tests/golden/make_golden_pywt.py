"""Settle the packed layout of the level-3 Haar DWT against a real PyWavelets install (NOT available in the build container):

    pip install PyWavelets && python tests/golden/make_golden_pywt.py     # writes tests/golden/golden_pywt.npz and prints the verdict

The reference calls pywt.wavedec2(x, 'haar', level=3, axes=(-2,-1)) + pywt.coeffs_to_array (condition/utils.py:116-123).
kdip restates that as one slot table with two candidates (oracle/transforms_ref.py::DWT_LAYOUTS, csrc/transforms.cu::band_offset,
switch KDIP_DWT_LAYOUT=code|diagram).  Once golden_pywt.npz exists, tests/test_oracle_golden.py::test_dwt_layout_pinned checks the
default layout against it and the 'parity unpinned' caveat goes away."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    import pywt
    import inputs as I
    from oracle import transforms_ref as T
    x = I.image(64, batch=1, seed=6).numpy()
    arr, _ = pywt.coeffs_to_array(pywt.wavedec2(x, "haar", level=3, axes=(-2, -1)), axes=(-2, -1))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_pywt.npz"), x=x, dwt=arr.astype(np.float32),
                        pywt_version=np.array(pywt.__version__))
    for name, slot in T.DWT_LAYOUTS.items():
        err = np.abs(T.dwt_forward(torch.tensor(x), slot=slot).numpy() - arr).max()
        print(f"layout '{name}': max abs difference to PyWavelets {pywt.__version__} = {err:.3e}")


if __name__ == "__main__":
    main()
