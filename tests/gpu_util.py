"""Helpers for the -m gpu parity tests: everything goes through the C ABI (kdip._lib), never through the oracle."""
import ctypes

import torch

from kdip._lib import ConvDesc, check, lib, ptr, stream_ptr


def to_nhwc_bf16(x):
    """fp32 NCHW torch tensor -> bf16 NHWC via the library's converter."""
    N, C, H, W = x.shape
    out = torch.empty(N, H, W, C, dtype=torch.bfloat16, device=x.device)
    check(lib.kdip_nchw_f32_to_nhwc_bf16(ptr(x.contiguous().float()), N, C, H, W, ptr(out), stream_ptr()))
    return out


def to_nchw_f32(x):
    N, H, W, C = x.shape
    out = torch.empty(N, C, H, W, dtype=torch.float32, device=x.device)
    check(lib.kdip_nhwc_bf16_to_nchw_f32(ptr(x.contiguous()), N, C, H, W, ptr(out), stream_ptr()))
    return out


def pad_rows(c):
    return 16 if c <= 16 else (32 if c <= 32 else (c + 63) // 64 * 64)


def pack_weight(w, flip=False, ci_off=0, ci_sub=None):
    """w fp32 [O,I,kh,kw] (cuda) -> packed bf16 [taps*rows_pad, cols_pad]."""
    O, I = w.shape[0], w.shape[1]
    taps = w.shape[2] * w.shape[3] if w.dim() == 4 else 1
    if ci_sub is not None:
        w = w[:, ci_off:ci_off + ci_sub].contiguous()
        I = ci_sub
    rows, cols = (I, O) if flip else (O, I)
    rp, cp = pad_rows(rows), (cols + 63) // 64 * 64
    dst = torch.empty(taps * rp, cp, dtype=torch.bfloat16, device=w.device)
    check(lib.kdip_pack_conv_weight(ptr(w.contiguous().float()), O, I, taps, rp, cp, int(flip), ptr(dst), stream_ptr()))
    return dst, rp


def run_conv(segs, N, H, W, cout, bias=None, residual=None, res_mode=0, out_mode=0, out_scale=1.0, stats=None, gn=None, in_gn=None):
    """segs: list of (act bf16 NHWC, packed weight, taps).  Returns the output tensor.
    in_gn = (ab fp32 [N, C, 2], silu, [channel offset of each segment or None]): fused GroupNorm apply on the operand path."""
    d = ConvDesc()
    d.N, d.H, d.W = N, H, W
    cout_pad = pad_rows(cout)
    d.Cout_pad, d.Cout, d.nseg = cout_pad, cout, len(segs)
    keep = []
    for i, (act, wp, taps) in enumerate(segs):
        d.seg[i].act = act.data_ptr()
        d.seg[i].C = act.shape[-1]
        d.seg[i].wgt = wp.data_ptr()
        d.seg[i].taps = taps
        keep += [act, wp]
    if bias is not None:
        b = torch.zeros(cout_pad, device=bias.device)
        b[:cout] = bias
        d.bias = b.data_ptr()
        keep.append(b)
    if residual is not None:
        d.residual = residual.data_ptr()
    d.res_mode = res_mode
    dev = segs[0][0].device
    if out_mode == 0:
        out = torch.full((N, H, W, cout), float("nan"), dtype=torch.bfloat16, device=dev)
    else:
        out = torch.full((N, cout, H, W), float("nan"), dtype=torch.float32, device=dev)
    d.out, d.out_mode, d.out_scale = out.data_ptr(), out_mode, out_scale
    if stats is not None:
        d.chan_stats = stats.data_ptr()
    if gn is not None:   # fused GroupNorm-backward reduction: (x0 bf16 NHWC, x1 or None, ab [N,C,2], silu, red [N,C,2] zeroed)
        x0, x1, ab, silu, red = gn
        d.gn_x0, d.gn_C0, d.gn_silu = x0.data_ptr(), x0.shape[-1], int(silu)
        d.gn_x1 = x1.data_ptr() if x1 is not None else None
        d.gn_ab, d.gn_red = ab.data_ptr(), red.data_ptr()
    if in_gn is not None:
        ab, silu, offs = in_gn
        for i, off in enumerate(offs):
            if off is not None:
                d.in_ab[i] = ab.data_ptr() + off * 8
        d.in_ab_C, d.in_silu = ab.shape[1], int(silu)
    plan = ctypes.c_void_p()
    check(lib.kdip_conv_plan_create(ctypes.byref(d), ctypes.byref(plan)))
    try:
        check(lib.kdip_conv_plan_run(plan, stream_ptr()))
        torch.cuda.synchronize()
    finally:
        lib.kdip_conv_plan_destroy(plan)
    return out


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
