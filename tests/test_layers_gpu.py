"""Per-layer parity of the non-GEMM UNet kernels (GroupNorm fwd/bwd, attention fwd/bwd, direct conv, embeddings)
against a plain PyTorch fp32 reference of the same op on the same bf16-rounded inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1.0 / 100


def _mk(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


def _bf(x):
    return x.to(torch.bfloat16).float()


def _resample(y, rs):
    return {0: y, 1: F.avg_pool2d(y, 2), 2: F.interpolate(y, scale_factor=2, mode="nearest")}[rs]


@pytest.mark.parametrize("C0,C1,rs,silu,film", [(64, 0, 0, 1, False), (128, 64, 0, 1, True), (192, 128, 1, 1, False),
                                                  (64, 0, 2, 1, True), (256, 0, 0, 0, False), (512, 256, 0, 1, True)])
def test_groupnorm_forward_backward(C0, C1, rs, silu, film):
    from kdip._lib import check, lib, ptr, stream_ptr
    from gpu_util import to_nchw_f32, to_nhwc_bf16, relerr
    N, H, W = 2, 16, 16
    C = C0 + C1
    x0 = _mk(N, C0, H, W, seed=1) * 1.5 + 0.3
    x1 = _mk(N, C1, H, W, seed=2) if C1 else None
    gamma, beta = 1 + 0.1 * _mk(C, seed=3), 0.1 * _mk(C, seed=4)
    filmt = _mk(N, 2 * C + 10, seed=5, scale=0.3) if film else None
    st = stream_ptr()
    s0, s1 = to_nhwc_bf16(x0), (to_nhwc_bf16(x1) if C1 else None)
    stats0 = torch.zeros(N, C0, 2, device="cuda")
    stats1 = torch.zeros(N, max(C1, 1), 2, device="cuda")
    check(lib.kdip_layer_chan_stats(ptr(s0), N, H * W, C0, ptr(stats0), st))
    if C1:
        check(lib.kdip_layer_chan_stats(ptr(s1), N, H * W, C1, ptr(stats1), st))
    ab = torch.empty(N, C, 2, device="cuda")
    mr = torch.empty(N, 32, 2, device="cuda")
    check(lib.kdip_layer_gn_finalize(ptr(stats0), C0, ptr(stats1) if C1 else None, C1, N, H * W, ptr(gamma), ptr(beta),
                                     ptr(filmt), (2 * C + 10) if film else 0, 10 if film else 0, ptr(ab), ptr(mr), st))
    Ho, Wo = {0: (H, W), 1: (H // 2, W // 2), 2: (2 * H, 2 * W)}[rs]
    out = torch.empty(N, Ho, Wo, C, dtype=torch.bfloat16, device="cuda")
    check(lib.kdip_layer_gn_apply(ptr(s0), C0, ptr(s1), C1, N, H, W, ptr(ab), silu, rs, ptr(out), st))

    # reference (unet.py:237-253 pattern) with autograd for the backward
    xin = torch.cat([_bf(x0)] + ([_bf(x1)] if C1 else []), 1).requires_grad_()
    u = F.group_norm(xin, 32, gamma, beta, eps=1e-5)
    if film:
        sc, sh = filmt[:, 10:10 + C, None, None], filmt[:, 10 + C:10 + 2 * C, None, None]
        u = u * (1 + sc) + sh
    y = _resample(F.silu(u) if silu else u, rs)
    e = relerr(to_nchw_f32(out), y.detach())
    print(f"gn fwd C={C0}+{C1} rs={rs} silu={silu} film={film}: rel err {e:.3e}")
    assert e < TOL

    gy = _mk(N, C, Ho, Wo, seed=6)
    extra = _mk(N, C, H, W, seed=7)
    (gx,) = torch.autograd.grad(y, xin, _bf(gy))
    gx = gx + _bf(extra)
    red = torch.zeros(N, C, 2, device="cuda")
    kk = torch.empty(N, C, 4, device="cuda")
    d0 = torch.empty(N, H, W, C0, dtype=torch.bfloat16, device="cuda")
    d1 = torch.empty(N, H, W, max(C1, 8), dtype=torch.bfloat16, device="cuda")
    gy_n, extra_n = to_nhwc_bf16(gy), to_nhwc_bf16(extra)   # keep references: ptr() does not own the tensor
    check(lib.kdip_layer_gn_bwd(ptr(s0), C0, ptr(s1), C1, N, H, W, ptr(ab), ptr(mr), silu, rs, ptr(gy_n),
                                ptr(extra_n), 1, ptr(red), ptr(kk), ptr(d0), ptr(d1) if C1 else None, st))
    got = to_nchw_f32(d0)
    if C1:
        got = torch.cat([got, to_nchw_f32(d1)[:, :C1]], 1)
    e = relerr(got, gx)
    print(f"gn bwd: rel err {e:.3e}")
    assert e < TOL

    # extra_mode 2: the extra gradient lives at g_y's resolution and goes through resample^T only (identity-skip path of an
    # up/down ResBlock, unet.py:190-197,257)
    extra2 = _mk(N, C, Ho, Wo, seed=8)
    z = xin.detach().clone().requires_grad_()
    (ge,) = torch.autograd.grad(_resample(z, rs), z, _bf(extra2))
    xin2 = xin.detach().clone().requires_grad_()
    u2 = F.group_norm(xin2, 32, gamma, beta, eps=1e-5)
    if film:
        u2 = u2 * (1 + sc) + sh
    y2 = _resample(F.silu(u2) if silu else u2, rs)
    (gx2,) = torch.autograd.grad(y2, xin2, _bf(gy))
    gx2 = gx2 + ge
    extra2_n = to_nhwc_bf16(extra2)
    red.zero_()
    check(lib.kdip_layer_gn_bwd(ptr(s0), C0, ptr(s1), C1, N, H, W, ptr(ab), ptr(mr), silu, rs, ptr(gy_n),
                                ptr(extra2_n), 2, ptr(red), ptr(kk), ptr(d0), ptr(d1) if C1 else None, st))
    got = to_nchw_f32(d0)
    if C1:
        got = torch.cat([got, to_nchw_f32(d1)[:, :C1]], 1)
    e = relerr(got, gx2)
    print(f"gn bwd (extra at g_y resolution): rel err {e:.3e}")
    assert e < TOL


@pytest.mark.parametrize("T,heads,N", [(64, 1, 2), (64, 8, 3), (128, 2, 3), (256, 3, 2), (256, 8, 1), (1024, 2, 1),
                                         (384, 2, 2), (512, 3, 1), (1024, 8, 2), (2048, 1, 1)])   # T >= 384: streamed-block tcgen05 kernels
def test_attention_forward_backward(T, heads, N):
    """QKVAttentionLegacy (unet.py:339-356): per-head interleaved [q,k,v], scale ch^-1/4 on q and k, fp32 softmax."""
    from kdip._lib import check, lib, ptr, stream_ptr
    from gpu_util import relerr
    ch = 64
    C = heads * ch
    qkv = _mk(N, 3 * C, T, seed=1)                                  # reference layout [N, H*3*C, T]
    qkv_b = _bf(qkv).requires_grad_()
    q, k, v = qkv_b.reshape(N * heads, ch * 3, T).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale).float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(N, -1, T)
    mine_in = qkv.permute(0, 2, 1).contiguous().to(torch.bfloat16)  # [N, T, 3C]
    out = torch.empty(N, T, C, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(N, heads, T, device="cuda")
    st = stream_ptr()
    check(lib.kdip_layer_attention_fwd(ptr(mine_in), N, T, heads, ptr(out), ptr(lse), st))
    e = relerr(out.float().permute(0, 2, 1), a.detach())
    print(f"attn fwd T={T}: rel err {e:.3e}")
    assert e < TOL
    ga = _mk(N, C, T, seed=2)
    (gq,) = torch.autograd.grad(a, qkv_b, _bf(ga))
    dqkv = torch.empty(N, T, 3 * C, dtype=torch.bfloat16, device="cuda")
    ga_n = ga.permute(0, 2, 1).contiguous().to(torch.bfloat16)
    check(lib.kdip_layer_attention_bwd(ptr(mine_in), ptr(out), ptr(ga_n), ptr(lse), N, T, heads, ptr(dqkv), st))
    e = relerr(dqkv.float().permute(0, 2, 1), gq)
    print(f"attn bwd T={T}: rel err {e:.3e}")
    assert e < 2 * TOL


@pytest.mark.parametrize("O,I,flip", [(128, 3, 0), (64, 3, 0), (6, 128, 1), (6, 64, 1)])
def test_conv_small_cin(O, I, flip):
    """first layer (3 -> C, with the c_in input scale fused) and the head's input-gradient (6 -> C)."""
    from kdip._lib import check, lib, ptr, stream_ptr
    from gpu_util import to_nchw_f32, relerr
    N, H, W = 2, 32, 32
    w = _mk(O, I, 3, 3, seed=1) / (I * 9) ** 0.5
    cin, cout = (O, I) if flip else (I, O)
    x = _mk(N, cin, H, W, seed=2)
    sc = torch.tensor([0.5, 2.0], device="cuda")
    b = _mk(cout, seed=3) if not flip else None
    wsc = torch.empty(9 * cin * cout, device="cuda")
    out = torch.empty(N, H, W, cout, dtype=torch.bfloat16, device="cuda")
    check(lib.kdip_layer_conv_small_cin(ptr(x), ptr(sc), ptr(w), ptr(b), N, O, I, flip, H, W, ptr(wsc), ptr(out), stream_ptr()))
    xs = x * sc[:, None, None, None]
    if flip:
        ref = F.conv_transpose2d(xs, w, padding=1)      # input-gradient of conv2d(., w, padding=1)
    else:
        ref = F.conv2d(xs, w, b, padding=1)
    e = relerr(to_nchw_f32(out), ref)
    print(f"small-cin O={O} I={I} flip={flip}: rel err {e:.3e}")
    assert e < TOL


def test_time_embedding():
    """nn.py:103-121 + unet.py:473-477 + emb_layers (unet.py:199-205)."""
    from kdip._lib import check, lib, ptr, stream_ptr
    from oracle.unet_ref import timestep_embedding   # oracle used as the checker only
    N, mc = 3, 128
    ted = 4 * mc
    t = torch.tensor([0.0, 37.0, 998.25], device="cuda")
    w1, b1 = _mk(ted, mc, seed=1, scale=mc ** -0.5), _mk(ted, seed=2, scale=0.1)
    w2, b2 = _mk(ted, ted, seed=3, scale=ted ** -0.5), _mk(ted, seed=4, scale=0.1)
    R = 640
    wall, ball = _mk(R, ted, seed=5, scale=ted ** -0.5), _mk(R, seed=6, scale=0.1)
    semb = torch.empty(N, ted, device="cuda")
    out = torch.empty(N, R, device="cuda")
    st = stream_ptr()
    check(lib.kdip_layer_time_embed(ptr(t), N, mc, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(semb), st))
    check(lib.kdip_layer_emb_proj(ptr(semb), N, ted, ptr(wall), ptr(ball), R, ptr(out), st))
    e0 = timestep_embedding(t.cpu(), mc).cuda()
    emb = F.linear(F.silu(F.linear(e0, w1, b1)), w2, b2)
    ref = F.linear(F.silu(emb), wall, ball)
    err = (out - ref).abs().max().item()
    print(f"time-embed abs err {err:.3e}")
    assert err < 2e-3     # fp32 path; sin/cos of arguments up to ~1000 rad
