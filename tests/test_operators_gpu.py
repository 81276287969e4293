"""Measurement operators, closed-form / CG mat solvers and orthogonal transforms through the product API (which calls
libkdip's C ABI) vs the reference's golden vectors (tests/golden) and the CPU oracle on the same seeded inputs.

Tolerances (fp32 kernels, different FFT factorisation / summation order than torch-CPU): 2e-5 relative to the output
scale for operators and closed forms; CG solutions 3e-4 of the solution scale (both CGs stop at 1e-4 relative residual).
Mask / gather / scatter outputs are compared bit-exactly."""
import numpy as np
import pytest
import torch

import inputs as I

pytestmark = pytest.mark.gpu

NAMES = ["gaussian_blur", "motion_blur", "super_resolution", "inpainting"]


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def make_op(name, size, dev="cuda"):
    from condition.measurements import get_operator
    if name == "inpainting":
        np.random.seed(0)
        return get_operator(name="inpainting", sigma_s=0.05, device=dev, mask_opt=dict(
            mask_type="box", mask_len_range=(size // 2, size // 2 + 1), image_size=size))
    if name == "super_resolution":
        return get_operator(name=name, in_shape=(1, 3, size, size), scale_factor=4, sigma_s=0.05, device=dev)
    return get_operator(name=name, in_shape=(1, 3, size, size), kernel_size=61, intensity=3.0 if name == "gaussian_blur" else 0.5,
                        sigma_s=0.05, device=dev)


def make_ref(name, size):
    from oracle import operators_ref as ops
    return {"gaussian_blur": lambda: ops.BlurOperator("gaussian_blur", 0.05, in_shape=(1, 3, size, size)),
            "motion_blur": lambda: ops.BlurOperator("motion_blur", 0.05, intensity=0.5, in_shape=(1, 3, size, size)),
            "super_resolution": lambda: ops.SuperResolutionOperator(0.05, 4, in_shape=(1, 3, size, size)),
            "inpainting": lambda: ops.InpaintingOperator(0.05, ops.box_mask(size, size // 2))}[name]()


def cpu_noise(shape, seed=2):
    torch.manual_seed(seed)
    return torch.randn(*shape)


def ref_noise(name, ref, x0, seed=2):
    """The measurement noise the reference drew after torch.manual_seed(seed): randn_like(y) fills in y's MEMORY order,
    and the Resizer's output is a transposed (non-contiguous) tensor, so the SR draw must reuse the oracle's strides."""
    torch.manual_seed(seed)
    if name == "super_resolution":
        return torch.randn_like(ref.down_sample(x0)).contiguous()
    return torch.randn(*x0.shape)


@pytest.mark.parametrize("size", [64, 256])
@pytest.mark.parametrize("name", NAMES)
def test_operator_forward_transpose_closed_form(name, size, golden_small):
    from condition.condition import __MAT_SOLVER__
    G = golden_small
    sub = I.sub if size == 256 else (lambda a: a)
    op, ref = make_op(name, size), make_ref(name, size)
    x0 = I.image(size, batch=1, seed=1)
    noise = ref_noise(name, ref, x0)                             # the reference's torch.manual_seed(2) draw
    y = op.handle.forward(x0.cuda(), noise.cuda())
    assert rel(sub(y.cpu()), G[f"op{size}.{name}.y"]) < 2e-5
    y0 = op.forward(x0.cuda(), noiseless=True)
    assert rel(sub(y0.cpu()), G[f"op{size}.{name}.y_noiseless"]) < 2e-5
    # full-resolution comparison against the oracle on the same noise
    y_ref = ref.forward(x0, noise=noise)
    assert rel(y, y_ref) < 2e-5
    At = op.transpose(y_ref.cuda())
    assert rel(sub(At.cpu()), G[f"op{size}.{name}.At_y"]) < 2e-5
    assert rel(At, ref.transpose(y_ref)) < 2e-5
    # flatten / transpose(flatten) are pure index ops
    _, yf = op.forward(x0.cuda(), flatten=True, noiseless=True)
    assert yf.shape[1] == int(G[f"op{size}.{name}.yflat_sum"][1])
    if name == "inpainting":
        y0r, yfr = ref.forward(x0, flatten=True, noiseless=True)
        assert torch.equal(y0.cpu(), y0r) and torch.equal(yf.cpu(), yfr), "mask / gather must be bit-exact"
        assert torch.equal(op.transpose(yf, flatten=True).cpu(), ref.transpose(yfr, flatten=True)), "scatter must be bit-exact"
    # closed-form mat (scalar variance)
    xm = I.image(size, batch=1, seed=4) * 0.8
    mat = __MAT_SOLVER__[name](op, y_ref.cuda(), xm.cuda(), torch.tensor([0.37]).cuda())
    assert rel(sub(mat.cpu()), G[f"mat{size}.{name}.scalar"]) < 5e-5


@pytest.mark.parametrize("ot", [None, "dct", "dwt"])
@pytest.mark.parametrize("name", NAMES)
def test_cg_mat(name, ot, golden_small):
    from condition.condition import __MAT_SOLVER__
    from condition.utils import OrthoTransform
    G = golden_small
    op, ref = make_op(name, 64), make_ref(name, 64)
    x0 = I.image(64, batch=1, seed=1)
    y_ref = ref.forward(x0, noise=ref_noise(name, ref, x0))
    xm = I.image(64, batch=1, seed=4) * 0.8
    th = I.theta_map(64, seed=5)
    mat = __MAT_SOLVER__[name](op, y_ref.cuda(), xm.cuda(), th.cuda(), OrthoTransform(ot))
    gold = G[f"mat64.{name}.cg.{ot}"]
    e = rel(mat, gold)
    print(f"cg {name} ot={ot}: rel err {e:.2e}, iters {op.handle.last_cg_iters}")
    assert e < 3e-4


def test_cg_batched_per_image_convergence():
    """B = 3 with per-image y / x0 / theta: each image must equal its own B = 1 solve (per-image alpha, beta, stop)."""
    from condition.condition import __MAT_SOLVER__
    op = make_op("gaussian_blur", 64)
    ys, xs, ths = [], [], []
    for b in range(3):
        x0 = I.image(64, batch=1, seed=10 + b)
        ys.append(op.handle.forward(x0.cuda(), cpu_noise((1, 3, 64, 64), seed=20 + b).cuda()))
        xs.append((I.image(64, batch=1, seed=30 + b) * 0.8).cuda())
        ths.append((I.theta_map(64, seed=40 + b) * (1 + 10 * b)).cuda())
    full = __MAT_SOLVER__["gaussian_blur"](op, torch.cat(ys), torch.cat(xs), torch.cat(ths))
    iters_full = list(op.handle.last_cg_iters)
    for b in range(3):
        one = __MAT_SOLVER__["gaussian_blur"](op, ys[b], xs[b], ths[b])
        assert op.handle.last_cg_iters[0] == iters_full[b]
        assert rel(full[b:b + 1], one) < 1e-5


def test_ortho_transforms(golden_small):
    from condition.utils import OrthoTransform
    xx = I.image(64, batch=1, seed=6).cuda()
    yy = I.image(64, batch=1, seed=8).cuda()
    for ot in ("dct", "dwt"):
        W = OrthoTransform(ot)
        assert rel(W(xx), golden_small[f"ot.{ot}.fwd"]) < 1e-5
        assert rel(W.inv(xx), golden_small[f"ot.{ot}.inv"]) < 1e-5
        assert rel(W.inv(W(xx)), xx) < 1e-5
        assert abs(((W(xx) * W(yy)).sum() - (xx * yy).sum()).item()) < 1e-2       # orthogonality
    # batch > 1: images stay independent (the reference is B = 1)
    xb = torch.cat([xx, yy])
    for ot in ("dct", "dwt"):
        W = OrthoTransform(ot)
        assert torch.equal(W(xb)[1:2], W(yy))


@pytest.mark.parametrize("name", ["gaussian_blur", "motion_blur", "super_resolution"])
def test_full_size_properties(name):
    """256x256, B = 8: adjointness <A x, y> = <x, A^T y> for the spectral model, and the CG solution's residual."""
    op = make_op(name, 256)
    B = 8
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 3, 256, 256, generator=g).cuda()
    if name == "super_resolution":
        # transpose() is the adjoint of the FFT model down(A x), not of the Resizer: check it against the mat solver's operator
        ysz = 64
        yv = torch.randn(B, 3, ysz, ysz, generator=g).cuda()
        Aty = op.transpose(yv)
        from condition.diffpir_utils.utils_sisr import downsample
        FB = op.handle.otf()[None, None]
        Ax = downsample(torch.fft.ifft2(FB * torch.fft.fft2(x)).real, 4)       # torch.fft only as the independent checker
        lhs, rhs = (Ax * yv).sum().item(), (x * Aty).sum().item()
    else:
        yv = torch.randn(B, 3, 256, 256, generator=g).cuda()
        lhs = (op.forward(x, noiseless=True) * yv).sum().item()
        rhs = (x * op.transpose(yv)).sum().item()
    assert abs(lhs - rhs) < 2e-3 * max(1.0, abs(lhs)), (lhs, rhs)
    # closed form at theta -> 0 solves sigma_s^2 v = A^T r / ... : check linearity in y instead (size independent)
    x0 = torch.rand(B, 3, 256, 256, generator=g).cuda() * 2 - 1
    th = torch.full((B,), 0.3).cuda()
    y1 = op.forward(x0, noiseless=True)
    m0 = op.handle.mat_closed(y1, x0, th)
    if name != "super_resolution":     # SR: y comes from the Resizer, the solver models down(A x): the residual is not zero
        assert m0.abs().max().item() < 1e-3, "y = A x0 must give mat = 0"
    m1 = op.handle.mat_closed(y1 + yv, x0, th)
    m2 = op.handle.mat_closed(y1 + 2 * yv, x0, th)
    assert rel(m2 - m0, 2 * (m1 - m0)) < 1e-4


@pytest.mark.parametrize("B", [2, 14])
@pytest.mark.parametrize("name", ["gaussian_blur", "motion_blur"])
def test_fused_spectral_filter_matches_three_passes(name, B, monkeypatch):
    """The single-launch cluster kernel (spec_filter_256_kernel: spectrum resident in distributed shared memory, Nyquist column
    packed into column 0) against the three-pass path and the oracle, for A, A^T and the closed-form mat with per-image theta;
    B = 14 (42 planes) runs more clusters than are co-resident."""
    op, ref = make_op(name, 256), make_ref(name, 256)
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1)
    yv = torch.randn(B, 3, 256, 256, generator=g)
    th = (torch.rand(B, generator=g) * 0.5 + 0.01)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("KDIP_FFT_FUSED", mode)
        h = op.handle
        outs[mode] = [h.forward(x.cuda(), None), h.transpose(yv.cuda()), h.mat_closed(yv.cuda(), x.cuda(), th.cuda())]
    for a, b in zip(outs["1"], outs["0"]):
        assert rel(a, b) < 2e-6, rel(a, b)
    assert rel(outs["1"][0][:2], ref.forward(x[:2], noiseless=True)) < 2e-5
    assert rel(outs["1"][1][:2], ref.transpose(yv[:2])) < 2e-5


def test_fused_dwt_covariance_matches_two_transforms(monkeypatch):
    """W diag(theta) W^T in one pass (dwt_cov_256_kernel) vs forward * theta -> inverse: the CG solutions at 256x256 with a
    DWT-domain theta map must be bit-identical (same arithmetic, same iteration counts)."""
    op = make_op("gaussian_blur", 256)
    B = 3
    g = torch.Generator().manual_seed(21)
    x0 = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).cuda()
    y = op.handle.forward(x0, torch.randn(B, 3, 256, 256, generator=g).cuda())
    tmap = (torch.rand(B, 3, 256, 256, generator=g) * 0.05 + 1e-4).cuda()
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("KDIP_DWT_FUSED", mode)
        res[mode] = (op.handle.mat_cg(y, x0, tmap, ot="dwt").clone(), list(op.handle.last_cg_iters))
    assert res["1"][1] == res["0"][1] and max(res["1"][1]) > 3, res["1"][1]
    assert torch.equal(res["1"][0], res["0"][0]), (res["1"][0] - res["0"][0]).abs().max().item()
