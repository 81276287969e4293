"""tcgen05 implicit-GEMM conv (csrc/conv_gemm.cu) vs torch conv2d on the same bf16-rounded operands (fp32 math).
Tolerance: the kernel accumulates in fp32 and stores bf16 -> relative error <= 2^-8 of the output scale."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 1.0 / 128  # bf16 store (2^-9 relative) + accumulation-order differences, relative to max |out|


def _mk(N, C, H, W, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(N, C, H, W, device="cuda", generator=g)


def _bf(x):
    return x.to(torch.bfloat16).float()


CASES = [
    # N, H, W, Cin, Cout, taps
    (2, 16, 16, 64, 64, 9),
    (1, 32, 32, 128, 128, 9),
    (2, 8, 8, 128, 256, 9),      # 8x8 level: two images per 128-pixel tile
    (1, 8, 8, 64, 64, 9),        # N=1 with TN=2: out-of-bounds image in the TMA box
    (3, 16, 16, 192, 64, 1),     # 1x1, K = 3 chunks
    (2, 16, 16, 64, 576, 1),     # qkv-like: 9 N-tiles of 64
    (1, 64, 64, 128, 128, 9),    # many tiles per CTA? (32 tiles) + persistent loop
    (4, 32, 32, 256, 512, 9),    # BN=256, 2 N-tiles, deep K
    (2, 32, 32, 320, 192, 9),    # non power-of-two channel counts (tiny config)
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_plain(case):
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    N, H, W, Ci, Co, taps = case
    k = 3 if taps == 9 else 1
    x = _mk(N, Ci, H, W, 1)
    w = _mk(Co, Ci, k, k, 2) / (Ci * taps) ** 0.5
    b = _mk(1, Co, 1, 1, 3).flatten()
    wp, _ = pack_weight(w)
    out = run_conv([(to_nhwc_bf16(x), wp, taps)], N, H, W, Co, bias=b)
    ref = F.conv2d(_bf(x), _bf(w), b, padding=k // 2)
    got = to_nchw_f32(out)
    assert torch.isfinite(got).all(), "NaN left in output: some pixels/channels were never written"
    e = relerr(got, ref)
    print(f"conv {case}: rel err {e:.3e}")
    if not e < TOL:   # failure forensics: where is it wrong?
        bad = ((got - ref).abs() > TOL * ref.abs().max())
        print("  bad fraction", bad.float().mean().item())
        print("  bad per image", bad.float().mean((1, 2, 3)).tolist())
        print("  bad per channel block of 16", bad.float().mean((0, 2, 3)).view(-1, 16).mean(1).tolist()[:16])
        print("  bad per row", bad.float().mean((0, 1, 3)).tolist()[:32])
        print("  bad per col", bad.float().mean((0, 1, 2)).tolist()[:32])
        print("  sample got/ref", got.flatten()[:8].tolist(), ref.flatten()[:8].tolist())
    assert e < TOL


def test_conv_dgrad_matches_autograd():
    """flip_transpose packing = input-gradient of the conv (condition/condition.py:172 autograd site)."""
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    N, H, W, Ci, Co = 2, 16, 16, 128, 64
    x = _mk(N, Ci, H, W, 1).requires_grad_()
    w = _mk(Co, Ci, 3, 3, 2) / (Ci * 9) ** 0.5
    g = _mk(N, Co, H, W, 4)
    y = F.conv2d(x, _bf(w), padding=1)
    (gx,) = torch.autograd.grad(y, x, _bf(g))
    wd, _ = pack_weight(w, flip=True)
    out = run_conv([(to_nhwc_bf16(g), wd, 9)], N, H, W, Ci)
    e = relerr(to_nchw_f32(out), gx)
    print(f"dgrad rel err {e:.3e}")
    assert e < TOL


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_conv_residual_modes(mode):
    """identity / avg-pool / nearest-up skip paths of ResBlock (unet.py:190-197,257) fused in the epilogue."""
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    N, H, W, C = 2, 16, 16, 64
    x = _mk(N, C, H, W, 1)
    w = _mk(C, C, 3, 3, 2) / (C * 9) ** 0.5
    rs = {1: (H, W), 2: (2 * H, 2 * W), 3: (H // 2, W // 2)}[mode]
    r = _mk(N, C, rs[0], rs[1], 5)
    wp, _ = pack_weight(w)
    out = run_conv([(to_nhwc_bf16(x), wp, 9)], N, H, W, C, residual=to_nhwc_bf16(r), res_mode=mode)
    rr = _bf(r)
    rr = {1: rr, 2: F.avg_pool2d(rr, 2), 3: F.interpolate(rr, scale_factor=2, mode="nearest")}[mode]
    ref = F.conv2d(_bf(x), _bf(w), padding=1) + rr
    e = relerr(to_nchw_f32(out), ref)
    print(f"residual mode {mode}: rel err {e:.3e}")
    assert e < TOL


@pytest.mark.parametrize("case", [(3, 8, 8, 128), (2, 32, 32, 256), (1, 64, 64, 64)], ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("variant", ["default", "mt2", "nopair"])
def test_conv_nearest_up_skip_tma(case, variant, monkeypatch):
    """Upsample ResBlock skip (unet.py:107,190-197): half-resolution source box TMA-loaded per output tile, in every tile mode."""
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    if variant == "mt2":
        monkeypatch.setenv("KDIP_CONV_MT", "2")
    if variant == "nopair":
        monkeypatch.setenv("KDIP_CONV_PAIR", "0")
        monkeypatch.setenv("KDIP_CONV_MT", "1")
    N, H, W, C = case
    x = _mk(N, C, H, W, 1)
    w = _mk(C, C, 3, 3, 2) / (C * 9) ** 0.5
    r = _mk(N, C, H // 2, W // 2, 5)
    b = _mk(1, C, 1, 1, 3).flatten()
    stats = torch.zeros(N, C, 2, device="cuda")
    out = run_conv([(to_nhwc_bf16(x), pack_weight(w)[0], 9)], N, H, W, C, bias=b, residual=to_nhwc_bf16(r), res_mode=3, stats=stats)
    ref = F.conv2d(_bf(x), _bf(w), b, padding=1) + F.interpolate(_bf(r), scale_factor=2, mode="nearest")
    got = to_nchw_f32(out)
    e = relerr(got, ref)
    print(f"nearest-up skip {case} {variant}: rel err {e:.3e}")
    assert e < TOL
    assert torch.allclose(stats[..., 0], got.sum((2, 3)), rtol=1e-3, atol=2e-2)


@pytest.mark.parametrize("case", [
    (2, 16, 16, 64, 64, 9, 0),      # 4 pixel tiles -> 2 work items of 2 tiles
    (4, 32, 32, 128, 128, 9, 1),    # identity residual, both tiles of a work item
    (3, 8, 8, 128, 128, 9, 0),      # TN=2, odd N: 2 pixel tiles, the second one half out of range
    (2, 32, 32, 192, 64, 1, 1),     # 1x1
    (2, 16, 16, 64, 256, 9, 0),     # Cout 256 -> N tile 128 or 64 with several N tiles per pixel-tile pair
], ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("pairmt", ["1", "0"])
def test_conv_two_tiles_per_work_item(case, pairmt, monkeypatch):
    """mt = 2: two M=128 accumulators share every weight stage (forced on small shapes through KDIP_CONV_MT=2), as single
    CTAs and as CTA pairs (four pixel tiles per work item)."""
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    monkeypatch.setenv("KDIP_CONV_MT", "2")
    monkeypatch.setenv("KDIP_CONV_PAIRMT", pairmt)
    monkeypatch.setenv("KDIP_CONV_WS", "1" if case[0] % 2 == 0 else "0")   # weight-stationary MMA pairs and plain pairs
    N, H, W, Ci, Co, taps, res = case
    k = 3 if taps == 9 else 1
    x = _mk(N, Ci, H, W, 1)
    w = _mk(Co, Ci, k, k, 2) / (Ci * taps) ** 0.5
    b = _mk(1, Co, 1, 1, 3).flatten()
    r = _mk(N, Co, H, W, 5) if res else None
    stats = torch.zeros(N, Co, 2, device="cuda")
    out = run_conv([(to_nhwc_bf16(x), pack_weight(w)[0], taps)], N, H, W, Co, bias=b,
                   residual=to_nhwc_bf16(r) if res else None, res_mode=res, stats=stats)
    ref = F.conv2d(_bf(x), _bf(w), b, padding=k // 2) + (_bf(r) if res else 0)
    got = to_nchw_f32(out)
    assert torch.isfinite(got).all()
    e = relerr(got, ref)
    print(f"mt=2 conv {case}: rel err {e:.3e}")
    assert e < TOL
    assert torch.allclose(stats[..., 0], got.sum((2, 3)), rtol=1e-3, atol=2e-2)
    assert torch.allclose(stats[..., 1], (got * got).sum((2, 3)), rtol=1e-3, atol=2e-2)


@pytest.mark.parametrize("case", [
    (2, 16, 16, 64, 128, 0, 1, 9),     # one source, SiLU
    (3, 8, 8, 128, 192, 128, 1, 9),    # two sources (128 + 64), TN=2 with an out-of-range image
    (2, 32, 32, 192, 128, 64, 1, 9),   # two sources (64 + 64), CTA pairs / two tiles per work item
    (2, 16, 16, 192, 256, 0, 0, 1),    # attention: identity activation, 1x1 (qkv input-gradient)
], ids=lambda c: "x".join(map(str, c)))
def test_conv_fused_groupnorm_backward_reduction(case):
    """kdip_conv_desc.gn_*: red[n][c] = (sum g_u, sum g_u x), g_u = g act'(A x + B), over the stored (bf16) conv output g."""
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    N, H, W, Ci, C, C0, silu, taps = case
    k = 3 if taps == 9 else 1
    gin = _mk(N, Ci, H, W, 1)
    w = _mk(C, Ci, k, k, 2) / (Ci * taps) ** 0.5
    x = _mk(N, C, H, W, 3) * 1.3 + 0.2
    ab = torch.stack([1 + 0.2 * _mk(N, C, 1, 1, 4).flatten(0), 0.3 * _mk(N, C, 1, 1, 5).flatten(0)], -1).reshape(N, C, 2).contiguous()
    red = torch.zeros(N, C, 2, device="cuda")
    if C0:
        x0, x1 = to_nhwc_bf16(x[:, :C0].contiguous()), to_nhwc_bf16(x[:, C0:].contiguous())
    else:
        x0, x1 = to_nhwc_bf16(x), None
    out = run_conv([(to_nhwc_bf16(gin), pack_weight(w)[0], taps)], N, H, W, C, gn=(x0, x1, ab, silu, red))
    g = to_nchw_f32(out)
    ref_out = F.conv2d(_bf(gin), _bf(w), padding=k // 2)
    assert relerr(g, ref_out) < TOL
    xb = _bf(x)
    u = ab[:, :, 0, None, None] * xb + ab[:, :, 1, None, None]
    if silu:
        sg = torch.sigmoid(u)
        gu = g * (sg * (1 + u * (1 - sg)))
    else:
        gu = g
    r1, r2 = gu.sum((2, 3)), (gu * xb).sum((2, 3))
    scale = max(r1.abs().max().item(), r2.abs().max().item())
    e1, e2 = (red[..., 0] - r1).abs().max().item() / scale, (red[..., 1] - r2).abs().max().item() / scale
    print(f"fused GN-bwd reduction {case}: rel err {e1:.2e} {e2:.2e}")
    assert e1 < 2e-3 and e2 < 2e-3     # tanh.approx SiLU' (2^-11) on bf16 operands, fp32 accumulation


HALO_CASES = [
    # N, H, W, Cin, Cout, residual mode (0 none, 1 identity, 3 nearest-up), extra 1x1 segments
    (1, 4, 128, 64, 64, 0, 0),
    (2, 8, 256, 128, 128, 1, 0),      # two column blocks per row, identity skip
    (1, 6, 128, 128, 192, 0, 0),      # N tile 64 x 3
    (2, 4, 128, 64, 256, 3, 0),       # two N tiles of 128, nearest-up skip from a [2, 64] source
    (1, 8, 128, 128, 128, 0, 2),      # conv3x3 + two 1x1 skip segments (channel-changing ResBlock over a concatenated input)
]


@pytest.mark.parametrize("case", HALO_CASES, ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("pair", ["1", "0"])
def test_conv_halo_pipeline(case, pair, monkeypatch):
    """Halo pipeline (images >= 128 pixels wide): rows y-1..y+2 loaded once per chunk, taps read them at row offsets dx+1;
    as CTA pairs (four rows per work item, weight rows split) and as single CTAs."""
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    monkeypatch.setenv("KDIP_CONV_HALO", "1")
    monkeypatch.setenv("KDIP_HALO_PAIR", pair)
    N, H, W, Ci, Co, res, nskip = case
    x = _mk(N, Ci, H, W, 1)
    w = _mk(Co, Ci, 3, 3, 2) / (Ci * 9) ** 0.5
    b = _mk(1, Co, 1, 1, 3).flatten()
    segs = [(to_nhwc_bf16(x), pack_weight(w)[0], 9)]
    ref = F.conv2d(_bf(x), _bf(w), b, padding=1)
    if nskip:
        C0, C1 = 128, 64
        s0, s1 = _mk(N, C0, H, W, 7), _mk(N, C1, H, W, 8)
        ws = _mk(Co, C0 + C1, 1, 1, 9) / (C0 + C1) ** 0.5
        segs += [(to_nhwc_bf16(s0), pack_weight(ws, ci_off=0, ci_sub=C0)[0], 1), (to_nhwc_bf16(s1), pack_weight(ws, ci_off=C0, ci_sub=C1)[0], 1)]
        ref = ref + F.conv2d(torch.cat([_bf(s0), _bf(s1)], 1), _bf(ws))
    r = None
    if res == 1:
        r = _mk(N, Co, H, W, 5)
        ref = ref + _bf(r)
    elif res == 3:
        r = _mk(N, Co, H // 2, W // 2, 5)
        ref = ref + F.interpolate(_bf(r), scale_factor=2, mode="nearest")
    stats = torch.zeros(N, Co, 2, device="cuda")
    out = run_conv(segs, N, H, W, Co, bias=b, residual=to_nhwc_bf16(r) if r is not None else None, res_mode=res, stats=stats)
    got = to_nchw_f32(out)
    assert torch.isfinite(got).all(), "NaN left in output: some pixels/channels were never written"
    e = relerr(got, ref)
    print(f"halo conv {case}: rel err {e:.3e}")
    if not e < TOL:
        bad = ((got - ref).abs() > TOL * ref.abs().max())
        print("  bad fraction", bad.float().mean().item(), "per row", bad.float().mean((0, 1, 3)).tolist(), "per col/16",
              bad.float().mean((0, 1, 2)).view(-1, 16).mean(1).tolist())
    assert e < TOL
    assert torch.allclose(stats[..., 0], got.sum((2, 3)), rtol=1e-3, atol=3e-2)
    assert torch.allclose(stats[..., 1], (got * got).sum((2, 3)), rtol=1e-3, atol=3e-2)


XF_CASES = [
    # N, H, W, C0, C1 (second concatenated source), Cout, silu, extra raw 1x1 skip segment, residual
    (1, 4, 128, 64, 0, 64, 1, 0, 0),
    (2, 8, 256, 128, 0, 128, 1, 0, 1),     # two column blocks per row (left / right image edges in different tiles), identity skip
    (3, 6, 128, 128, 64, 128, 1, 0, 0),    # concatenated input = two 3x3 K-segments, odd number of row pairs (single CTAs)
    (2, 4, 128, 64, 0, 192, 0, 0, 0),      # no activation (attention-style norm), N tile 64 x 3
    (2, 8, 128, 128, 0, 128, 1, 1, 0),     # conv2 of a channel-changing ResBlock: normalised 3x3 segment + raw 1x1 skip segment
]


@pytest.mark.parametrize("case", XF_CASES, ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("pair", ["1", "0"])
def test_conv_fused_groupnorm_apply(case, pair, monkeypatch):
    """GroupNorm affine (+SiLU) applied to the landed rows in shared memory (kdip_conv_desc.in_ab) vs the separate gn_apply pass
    followed by the same conv: the operand bits are the same, so the outputs must be identical; and vs torch on the fp32 formula."""
    from kdip._lib import check, lib, ptr, stream_ptr
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    monkeypatch.setenv("KDIP_HALO_PAIR", pair)
    N, H, W, C0, C1, Co, silu, skip, res = case
    C = C0 + C1
    x0 = _mk(N, C0, H, W, 1) * 1.5 + 0.3
    x1 = _mk(N, C1, H, W, 2) if C1 else None
    g = torch.Generator(device="cuda").manual_seed(3)
    ab = torch.stack([1 + 0.3 * torch.randn(N, C, device="cuda", generator=g), 0.5 * torch.randn(N, C, device="cuda", generator=g)], -1).contiguous()
    w = _mk(Co, C, 3, 3, 4) / (C * 9) ** 0.5
    b = _mk(1, Co, 1, 1, 5).flatten()
    s0, s1 = to_nhwc_bf16(x0), (to_nhwc_bf16(x1) if C1 else None)
    # unfused: gn_apply -> normalised bf16 tensor -> conv
    a = torch.empty(N, H, W, C, dtype=torch.bfloat16, device="cuda")
    check(lib.kdip_layer_gn_apply(ptr(s0), C0, ptr(s1), C1, N, H, W, ptr(ab), silu, 0, ptr(a), stream_ptr()))
    extra, extra_ref = [], 0
    if skip:
        xs = _mk(N, 64, H, W, 7)
        ws = _mk(Co, 64, 1, 1, 8) / 8
        extra = [(to_nhwc_bf16(xs), pack_weight(ws)[0], 1)]
        extra_ref = F.conv2d(_bf(xs), _bf(ws))
    r = to_nhwc_bf16(_mk(N, Co, H, W, 9)) if res else None
    kw = dict(bias=b, residual=r, res_mode=1 if res else 0)
    st_a, st_b = torch.zeros(N, Co, 2, device="cuda"), torch.zeros(N, Co, 2, device="cuda")
    ref = run_conv([(a, pack_weight(w)[0], 9)] + extra, N, H, W, Co, stats=st_a, **kw)
    if C1:
        segs = [(s0, pack_weight(w, ci_off=0, ci_sub=C0)[0], 9), (s1, pack_weight(w, ci_off=C0, ci_sub=C1)[0], 9)]
        offs = [0, C0]
    else:
        segs, offs = [(s0, pack_weight(w)[0], 9)], [0]
    got = run_conv(segs + extra, N, H, W, Co, stats=st_b, in_gn=(ab, silu, offs + [None] * len(extra)), **kw)
    assert torch.isfinite(got.float()).all()
    if C1 == 0:
        assert torch.equal(got, ref), f"fused != unfused: max diff {(got.float() - ref.float()).abs().max().item()}"
        assert torch.equal(st_a, st_b) or torch.allclose(st_a, st_b, rtol=1e-5, atol=1e-3)   # atomics: order only
    else:   # two K-segments accumulate in a different order than one concatenated segment
        assert relerr(got, ref) < 1.0 / 256
    # torch: fp32 affine (+SiLU), rounded to bf16 like the stored operand, zero padding applied AFTER the normalisation
    xin = torch.cat([_bf(x0)] + ([_bf(x1)] if C1 else []), 1)
    u = ab[:, :, 0, None, None] * xin + ab[:, :, 1, None, None]
    u = _bf(F.silu(u) if silu else u)
    t = F.conv2d(u, _bf(w), b, padding=1) + extra_ref
    if res:
        t = t + to_nchw_f32(r)
    e = relerr(to_nchw_f32(got), t)
    print(f"fused GN-apply conv {case} pair={pair}: rel err vs torch {e:.3e}")
    assert e < TOL


@pytest.mark.parametrize("pair", ["1", "0"])
def test_conv_fused_groupnorm_apply_repeatable(pair, monkeypatch):
    """Pipeline-protocol regression (ring of 7 row slots shared by two transform warpgroups): many work items per CTA, hundreds of
    launches with a cache-thrashing copy in between so the rows land out of order - every output must be bit-identical to the first
    (a row transformed twice or consumed early changes bits; a lost barrier phase traps)."""
    import ctypes
    from kdip._lib import ConvDesc, check, lib, ptr, stream_ptr
    from gpu_util import pack_weight, to_nhwc_bf16
    monkeypatch.setenv("KDIP_HALO_PAIR", pair)
    N, H, W, C, Co = 12, 256, 256, 128, 128
    x = to_nhwc_bf16(_mk(N, C, H, W, 1) * 1.5 + 0.3)
    g = torch.Generator(device="cuda").manual_seed(3)
    ab = torch.stack([1 + 0.3 * torch.randn(N, C, device="cuda", generator=g), 0.5 * torch.randn(N, C, device="cuda", generator=g)], -1).contiguous()
    wp = pack_weight(_mk(Co, C, 3, 3, 4) / (C * 9) ** 0.5)[0]
    outs = [torch.empty(N, H, W, Co, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    plans = []
    for o in outs:
        d = ConvDesc()
        d.N, d.H, d.W, d.Cout_pad, d.Cout, d.nseg = N, H, W, Co, Co, 1
        d.seg[0].act, d.seg[0].C, d.seg[0].wgt, d.seg[0].taps = x.data_ptr(), C, wp.data_ptr(), 9
        d.out, d.out_mode, d.out_scale = o.data_ptr(), 0, 1.0
        d.in_ab[0], d.in_ab_C, d.in_silu = ab.data_ptr(), C, 1
        pl = ctypes.c_void_p()
        check(lib.kdip_conv_plan_create(ctypes.byref(d), ctypes.byref(pl)))
        plans.append(pl)
    try:
        check(lib.kdip_conv_plan_run(plans[0], stream_ptr()))
        ref = outs[0].clone()
        thrash = torch.empty(96 << 20, dtype=torch.uint8, device="cuda")
        bad = torch.zeros((), dtype=torch.int64, device="cuda")
        for it in range(300):
            if it % 3 == 0:
                thrash.add_(1)                      # evicts x from L2: the next launch's rows come from HBM
            k = it & 1
            check(lib.kdip_conv_plan_run(plans[k], stream_ptr()))
            bad += (outs[k] != ref).sum()
        torch.cuda.synchronize()
        assert int(bad.item()) == 0, f"{int(bad.item())} output elements changed between identical launches"
    finally:
        for pl in plans:
            lib.kdip_conv_plan_destroy(pl)


def test_conv_fused_groupnorm_apply_needs_halo():
    """in_ab on a shape the row-tile pipeline does not cover is refused (never silently ignored)."""
    from gpu_util import pack_weight, run_conv, to_nhwc_bf16
    x, w = _mk(1, 64, 16, 16, 1), _mk(64, 64, 3, 3, 2)
    ab = torch.ones(1, 64, 2, device="cuda")
    with pytest.raises(ValueError):
        run_conv([(to_nhwc_bf16(x), pack_weight(w)[0], 9)], 1, 16, 16, 64, in_gn=(ab, 1, [0]))


def test_conv_halo_matches_tile_pipeline():
    """Same conv through the halo pipeline (default) and (KDIP_CONV_HALO=0) the 8x16-tile pipeline: identical up to fp32 summation order."""
    import os
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    N, H, W, Ci, Co = 2, 16, 256, 128, 128
    x, w = _mk(N, Ci, H, W, 1), _mk(Co, Ci, 3, 3, 2) / (Ci * 9) ** 0.5
    xa, wp = to_nhwc_bf16(x), pack_weight(w)[0]
    a = to_nchw_f32(run_conv([(xa, wp, 9)], N, H, W, Co))          # default: halo pipeline
    os.environ["KDIP_CONV_HALO"] = "0"
    try:
        b = to_nchw_f32(run_conv([(xa, wp, 9)], N, H, W, Co))      # 8x16-tile pipeline
    finally:
        del os.environ["KDIP_CONV_HALO"]
    assert relerr(a, b) < 1.0 / 256


def test_conv_three_segments():
    """conv3x3(a2) + 1x1 skip over two concatenated sources accumulated in one TMEM tile (unet.py:222,257,662)."""
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
    N, H, W, Co, C0, C1 = 2, 16, 16, 128, 192, 64
    a2, s0, s1 = _mk(N, Co, H, W, 1), _mk(N, C0, H, W, 2), _mk(N, C1, H, W, 3)
    w2 = _mk(Co, Co, 3, 3, 4) / (Co * 9) ** 0.5
    ws = _mk(Co, C0 + C1, 1, 1, 5) / (C0 + C1) ** 0.5
    b = _mk(1, Co, 1, 1, 6).flatten()
    segs = [(to_nhwc_bf16(a2), pack_weight(w2)[0], 9),
            (to_nhwc_bf16(s0), pack_weight(ws, ci_off=0, ci_sub=C0)[0], 1),
            (to_nhwc_bf16(s1), pack_weight(ws, ci_off=C0, ci_sub=C1)[0], 1)]
    out = run_conv(segs, N, H, W, Co, bias=b)
    ref = F.conv2d(_bf(a2), _bf(w2), b, padding=1) + F.conv2d(torch.cat([_bf(s0), _bf(s1)], 1), _bf(ws))
    e = relerr(to_nchw_f32(out), ref)
    print(f"3-segment rel err {e:.3e}")
    assert e < TOL


@pytest.mark.parametrize("cout", [6, 3])
def test_conv_small_cout_fp32_nchw(cout):
    """output head (128 -> 6, unet.py:617) and first-layer input-gradient (C -> 3): N padded to 16, fp32 NCHW store."""
    from gpu_util import pack_weight, run_conv, to_nhwc_bf16, relerr
    N, H, W, Ci = 2, 32, 32, 128
    x = _mk(N, Ci, H, W, 1)
    w = _mk(cout, Ci, 3, 3, 2) / (Ci * 9) ** 0.5
    b = _mk(1, cout, 1, 1, 3).flatten()
    wp, _ = pack_weight(w)
    out = run_conv([(to_nhwc_bf16(x), wp, 9)], N, H, W, cout, bias=b, out_mode=1, out_scale=0.5)
    ref = 0.5 * F.conv2d(_bf(x), _bf(w), b, padding=1)
    assert torch.isfinite(out).all()
    e = relerr(out, ref)
    print(f"small-cout {cout}: rel err {e:.3e}")
    assert e < 1e-4          # fp32 store: only accumulation order differs


@pytest.mark.parametrize("case", [
    (3, 8, 8, 64, 64, 0),       # TN=2: one tile spans two images, odd N -> out-of-range image in the last tile
    (2, 32, 32, 128, 128, 0),   # CTA pairs, two 64-channel slabs
    (4, 16, 16, 64, 256, 1),    # BN=256: two staging passes; identity residual
    (1, 64, 64, 64, 192, 0),    # 3 slabs: the second pass is half empty
    (2, 16, 16, 64, 128, 2),    # avg-pool skip -> legacy epilogue statistics
], ids=lambda c: "x".join(map(str, c)))
def test_conv_chan_stats(case):
    """GroupNorm statistics (nn.py:17-19) fused into the conv epilogue = sums over the stored bf16 output."""
    from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16
    N, H, W, Ci, C, res_mode = case
    x = _mk(N, Ci, H, W, 1)
    w = _mk(C, Ci, 3, 3, 2) / (Ci * 9) ** 0.5
    stats = torch.zeros(N, C, 2, device="cuda")
    res = None
    if res_mode == 1:
        res = to_nhwc_bf16(_mk(N, C, H, W, 5))
    elif res_mode == 2:
        res = to_nhwc_bf16(_mk(N, C, 2 * H, 2 * W, 5))
    out = run_conv([(to_nhwc_bf16(x), pack_weight(w)[0], 9)], N, H, W, C, residual=res, res_mode=res_mode, stats=stats)
    got = to_nchw_f32(out)
    assert torch.isfinite(got).all()
    s1, s2 = got.sum((2, 3)), (got * got).sum((2, 3))
    assert torch.allclose(stats[..., 0], s1, rtol=1e-3, atol=2e-2), (stats[..., 0] - s1).abs().max()
    assert torch.allclose(stats[..., 1], s2, rtol=1e-3, atol=2e-2), (stats[..., 1] - s2).abs().max()


def test_conv_rejects_bad_shapes():
    from gpu_util import pack_weight, run_conv, to_nhwc_bf16
    x = _mk(1, 48, 16, 16, 1)     # 48 channels: not a multiple of 64
    w = _mk(64, 48, 3, 3, 2)
    with pytest.raises(ValueError):
        run_conv([(to_nhwc_bf16(x), pack_weight(w)[0], 9)], 1, 16, 16, 64)
