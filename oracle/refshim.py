"""Import the UNMODIFIED reference from /root/reference on CPU (build container only; test infrastructure).

The reference's arithmetic imports cleanly; only non-arithmetic third-party modules are missing here
(SURVEY.md §8(c), Appendix C).  This shim registers empty stand-ins for them, points ``hdf5storage.loadmat`` at
scipy, adapts ``scipy.sparse.linalg.cg(tol=...)`` to the current keyword (legacy semantics rtol=tol, atol=0) and
provides the Haar restatement as ``pywt`` (PyWavelets is absent: the DWT layout stays *parity unpinned*).
Nothing on the GPU box may import this module: /root/reference does not exist there.
"""
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "condition"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Returns a dict of the reference modules.  Changes cwd to the reference root (kernel files are opened by
    relative path, condition/measurements.py:95,134,173)."""
    if not available():
        raise RuntimeError("reference tree not present (only the build container has /root/reference)")
    import numpy as np
    import scipy.io
    import scipy.sparse.linalg as ssl

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    _mod("gpytorch", LinearOperator=object)
    _mod("gpytorch.distributions", MultivariateNormal=_Dummy)
    for n in ("skimage", "skimage.transform", "skimage.metrics", "matplotlib", "matplotlib.pyplot", "accelerate",
              "lpips", "clip", "cleanfid", "blobfile"):
        _mod(n)
    _mod("cleanfid.inception_torchscript", InceptionV3W=_Dummy)
    _mod("resize_right", resize=None)
    _mod("hdf5storage", loadmat=scipy.io.loadmat)
    _mod("torchdiffeq", odeint=None)
    _mod("torchsde", BrownianTree=_Dummy)
    _mod("mpi4py", MPI=_Dummy())

    def _merge(a, b):
        out = dict(a)
        for k, v in b.items():
            out[k] = _merge(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else v
        return out
    _mod("jsonmerge", merge=_merge)

    # pywt stand-in built from the oracle's Haar restatement (layout caveat: transforms_ref.py)
    from . import transforms_ref as T

    def wavedec2(x, wavelet="haar", level=3, axes=(-2, -1)):
        import torch
        arr = T.dwt_forward(torch.tensor(np.asarray(x, dtype=np.float32))).numpy()
        return ("packed", arr)

    def coeffs_to_array(c, axes=(-2, -1)):
        return c[1], "slices"

    def array_to_coeffs(a, slices, output_format="wavedec2"):
        return ("packed", np.asarray(a))

    def waverec2(c, wavelet="haar", axes=(-2, -1)):
        import torch
        return T.dwt_inverse(torch.tensor(np.asarray(c[1], dtype=np.float32))).numpy()
    _mod("pywt", wavedec2=wavedec2, coeffs_to_array=coeffs_to_array, array_to_coeffs=array_to_coeffs,
         waverec2=waverec2)

    os.chdir(REF)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import condition.condition as CC
    import condition.measurements as CM
    import condition.utils as CU
    import k_diffusion as K
    from condition.diffpir_utils import utils_model
    from guided_diffusion import script_util

    def cg_legacy(A, b, tol=1e-5, maxiter=None):
        return ssl.cg(A, b, rtol=tol, atol=0.0, maxiter=maxiter)
    CC.cg = cg_legacy
    return dict(CC=CC, CM=CM, CU=CU, K=K, utils_model=utils_model, script_util=script_util)


def build_reference_unet(mods, overrides):
    """create_model_and_diffusion with utils_model.create_argparser defaults (sample_condition_openai.py:115-129)."""
    su = mods["script_util"]
    args = mods["utils_model"].create_argparser(overrides).parse_args([])
    kw = su.args_to_dict(args, su.model_and_diffusion_defaults().keys())
    model, diffusion = su.create_model_and_diffusion(**kw)
    return model.eval(), diffusion
