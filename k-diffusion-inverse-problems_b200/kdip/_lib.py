"""ctypes binding of libkdip.so (the C-ABI drop-in boundary, include/kdip.h).

The prototypes are read from include/kdip.h itself, so the header stays the single source of truth.  There is no
CPU fallback: every compute entry point needs the CUDA library and a B200; a missing library raises ImportError
with the build command.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
HEADER = os.path.join(_ROOT, "include", "kdip.h")
LIB_PATH = os.path.join(_HERE, "libkdip.so")

KDIP_OK, KDIP_EINVAL, KDIP_ESHAPE, KDIP_EALIGN, KDIP_ECUDA, KDIP_ENOTCONV, KDIP_ENOMEM = 0, -1, -2, -3, -4, -5, -6


class KdipError(RuntimeError):
    pass


_SCALARS = {"int": ctypes.c_int, "float": ctypes.c_float, "size_t": ctypes.c_size_t, "int64_t": ctypes.c_int64,
            "int32_t": ctypes.c_int32, "kdip_stream_t": ctypes.c_void_p}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:^|\n)\s*(const char\*|unsigned long long|int|void)\s+(kdip_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = {"int": ctypes.c_int, "void": None, "const char*": ctypes.c_char_p,
                   "unsigned long long": ctypes.c_ulonglong}[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = a.replace("const ", "").split()[0]
                    argtypes.append(_SCALARS[base])
        protos[name] = (restype, argtypes)
    return protos


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(or k-diffusion-inverse-problems_b200/csrc/build.sh). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def check(rc):
    """Map KDIP_E* to the exception types the reference's Python raises (SURVEY.md §8(b) error convention)."""
    if rc == KDIP_OK:
        return
    msg = (lib.kdip_last_error() or b"").decode()
    if rc in (KDIP_EINVAL, KDIP_ESHAPE, KDIP_EALIGN):
        raise ValueError(f"libkdip: {msg}")
    if rc == KDIP_ENOMEM:
        raise MemoryError(f"libkdip: {msg}")
    raise KdipError(f"libkdip error {rc}: {msg}")


class PmvScalars(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("c_in", "recip", "recipm1", "min_log", "max_log", "post_var", "coef1_sq")]


class GuidedCfg(ctypes.Structure):
    """kdip_guided_cfg (include/kdip.h)."""
    _fields_ = [("guidance", ctypes.c_int), ("sigma", ctypes.c_float), ("t_model", ctypes.c_float), ("theta", ctypes.c_float),
                ("zeta", ctypes.c_float), ("sc", PmvScalars)]


class UNetProfile(ctypes.Structure):
    _fields_ = [("conv_ms", ctypes.c_float), ("other_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("conv_flops", ctypes.c_double), ("conv_launches", ctypes.c_int), ("other_steps", ctypes.c_int)]


class OpDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("S", ctypes.c_int), ("sf", ctypes.c_int), ("sigma_s", ctypes.c_float),
                ("psf", ctypes.c_void_p), ("ksize", ctypes.c_int), ("mask", ctypes.c_void_p), ("rs_w", ctypes.c_void_p),
                ("rs_idx", ctypes.c_void_p), ("rs_taps", ctypes.c_int)]


class ConvSeg(ctypes.Structure):
    _fields_ = [("act", ctypes.c_void_p), ("C", ctypes.c_int), ("wgt", ctypes.c_void_p), ("taps", ctypes.c_int)]


class ConvDesc(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("Cout_pad", ctypes.c_int),
                ("Cout", ctypes.c_int), ("nseg", ctypes.c_int), ("seg", ConvSeg * 3), ("bias", ctypes.c_void_p),
                ("residual", ctypes.c_void_p), ("res_mode", ctypes.c_int), ("out", ctypes.c_void_p),
                ("out_mode", ctypes.c_int), ("out_scale", ctypes.c_float), ("chan_stats", ctypes.c_void_p),
                ("gn_x0", ctypes.c_void_p), ("gn_x1", ctypes.c_void_p), ("gn_C0", ctypes.c_int), ("gn_silu", ctypes.c_int),
                ("gn_ab", ctypes.c_void_p), ("gn_red", ctypes.c_void_p), ("in_ab", ctypes.c_void_p * 3), ("in_ab_C", ctypes.c_int),
                ("in_silu", ctypes.c_int)]


class UNetArch(ctypes.Structure):
    _fields_ = [("image_size", ctypes.c_int), ("in_channels", ctypes.c_int), ("model_channels", ctypes.c_int),
                ("out_channels", ctypes.c_int), ("num_res_blocks", ctypes.c_int), ("num_head_channels", ctypes.c_int),
                ("n_mult", ctypes.c_int), ("channel_mult", ctypes.c_float * 8), ("n_att", ctypes.c_int),
                ("attention_ds", ctypes.c_int * 8), ("precision", ctypes.c_int)]


def stream_ptr():
    """Current torch CUDA stream as a cudaStream_t (autograd propagates it to backward threads)."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libkdip needs contiguous CUDA tensors (no CPU fallback)"
    return ctypes.c_void_p(t.data_ptr())
