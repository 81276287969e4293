"""kdip — host-side mirror of the reference's Denoiser / Condition / operator interface for the guided-sampling
hot path, backed by libkdip.so (hand-written sm_100a CUDA behind a C ABI).  No CPU fallback."""
from ._lib import KdipError, check, lib  # noqa: F401
