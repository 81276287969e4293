"""HBM-bound guidance / sampler kernels through the C ABI vs the oracle (fp32; tolerance 1e-6 relative unless noted;
mask / gather / scatter are bit-exact)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def test_sampler_updates():
    from kdip._lib import check, lib, ptr, stream_ptr
    B, n = 3, 3 * 64 * 64
    x, den, noise = _mk(B, n, seed=1) * 10, _mk(B, n, seed=2), _mk(B, n, seed=3)
    sigma, sigma_hat, sigma_next, s_noise = 2.0, 2.6, 1.4, 1.003
    st = stream_ptr()
    # NOTE: every device tensor handed to the C ABI must stay referenced until the kernel ran (no inline .cuda())
    xc, noise_c, den_c = x.cuda(), noise.cuda(), den.cuda()
    check(lib.kdip_churn(ptr(xc), ptr(noise_c), s_noise, sigma, sigma_hat, B * n, st))
    ref = x + noise * s_noise * (sigma_hat ** 2 - sigma ** 2) ** 0.5
    assert torch.allclose(xc.cpu(), ref, rtol=1e-6, atol=1e-5)
    dt = sigma_next - sigma_hat
    x2, d = torch.empty_like(xc), torch.empty_like(xc)
    check(lib.kdip_euler_step(ptr(xc), ptr(den_c), sigma_hat, dt, ptr(x2), ptr(d), B * n, st))
    d_ref = (ref - den) / sigma_hat
    x2_ref = ref + d_ref * dt
    assert torch.allclose(d.cpu(), d_ref, rtol=1e-6, atol=1e-5)
    assert torch.allclose(x2.cpu(), x2_ref, rtol=1e-6, atol=1e-5)
    den2 = _mk(B, n, seed=4)
    den2_c = den2.cuda()
    xo = torch.empty_like(xc)
    check(lib.kdip_heun_step(ptr(xc), ptr(d), ptr(x2), ptr(den2_c), sigma_next, dt, ptr(xo), B * n, st))
    d2 = (x2_ref - den2) / sigma_next
    assert torch.allclose(xo.cpu(), ref + (d_ref + d2) / 2 * dt, rtol=1e-6, atol=1e-5)


def _scalars(sched, sigma, t, B):
    from kdip._lib import PmvScalars
    arr = (PmvScalars * B)()
    for b in range(B):
        arr[b].c_in = float(1 / (sigma ** 2 + 1) ** 0.5)
        arr[b].recip = float(np.float32(sched.sqrt_recip_alphas_cumprod[t]))
        arr[b].recipm1 = float(np.float32(sched.sqrt_recipm1_alphas_cumprod[t]))
        arr[b].min_log = float(np.float32(sched.posterior_log_variance_clipped[t]))
        arr[b].max_log = float(np.float32(sched.log_betas[t]))
        arr[b].post_var = float(np.float32(sched.posterior_variance[t]))
        arr[b].coef1_sq = float(np.float32(sched.posterior_mean_coef1[t]) ** 2)
    buf = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    return buf


@pytest.mark.parametrize("sigma", [10.0, 0.1])
def test_pmv_epilogue_and_vjp_seed(sigma):
    from kdip._lib import check, lib, ptr, stream_ptr
    from oracle import diffusion_ref as D
    sched = D.Schedule()
    B, HW = 2, 64 * 64
    sig = torch.tensor([sigma])
    t = int(sched.sigma_to_t(sig).long())
    out, x = _mk(B, 6, HW, seed=1), _mk(B, 3, HW, seed=2) * (1 + sigma)
    c_in = 1 / (sigma ** 2 + 1) ** 0.5
    x0_ref, var_ref = D.pmv_epilogue(sched, out.view(B, 6, 64, 64), (x * c_in).view(B, 3, 64, 64), torch.tensor([t] * B))
    conv_ref = D.convert_variance(sched, var_ref, torch.tensor([t] * B))
    sc = _scalars(sched, sigma, t, B)
    x0, var = torch.empty(B, 3, HW, device="cuda"), torch.empty(B, 3, HW, device="cuda")
    st = stream_ptr()
    out_c, x_c = out.cuda(), x.cuda()
    check(lib.kdip_pmv_epilogue(ptr(out_c), ptr(x_c), ptr(sc), ptr(x0), ptr(var), 2, B, HW, st))
    assert torch.allclose(x0.cpu().view_as(x0_ref), x0_ref, rtol=1e-5, atol=2e-6)
    # Eq. 22 subtracts two nearly equal variances and divides by coef1^2 (tiny at large t): an ulp of exp() is amplified
    # by 1/coef1^2, so the absolute tolerance is 8 ulp(variance) * 1/coef1^2 (conditioning of the formula, not the kernel).
    amp = float(1.0 / np.float32(sched.posterior_mean_coef1[t]) ** 2)
    atol = 8 * 1.2e-7 * float(var_ref.max()) * amp
    assert torch.allclose(var.cpu().view_as(conv_ref), conv_ref, rtol=2e-4, atol=atol), (atol, (var.cpu().view_as(conv_ref) - conv_ref).abs().max())
    raw = torch.empty(B, 3, HW, device="cuda")
    check(lib.kdip_pmv_epilogue(ptr(out_c), ptr(x_c), ptr(sc), ptr(x0), ptr(raw), 1, B, HW, st))
    assert torch.allclose(raw.cpu().view_as(var_ref), var_ref, rtol=2e-6, atol=0)      # model variance: 1-2 ulp of exp
    # VJP seed vs autograd through the oracle epilogue
    v = _mk(B, 3, HW, seed=3)
    xo = x.clone().requires_grad_()
    oo = out.clone().requires_grad_()
    x0a, _ = D.pmv_epilogue(sched, oo.view(B, 6, 64, 64), (xo * c_in).view(B, 3, 64, 64), torch.tensor([t] * B))
    g_out, g_x = torch.autograd.grad((x0a * v.view_as(x0a)).sum(), [oo, xo])
    seed, direct = torch.empty(B, 6, HW, device="cuda"), torch.empty(B, 3, HW, device="cuda")
    v_c = v.cuda()
    check(lib.kdip_pmv_vjp_seed(ptr(x0), ptr(v_c), ptr(sc), ptr(seed), ptr(direct), B, HW, st))
    assert torch.allclose(seed.cpu(), g_out, rtol=1e-5, atol=1e-6)
    assert torch.allclose(direct.cpu(), g_x, rtol=1e-5, atol=1e-6)


def test_combine_and_inpaint_bit_exact():
    from kdip._lib import check, lib, ptr, stream_ptr
    from oracle import operators_ref as ops
    B, S = 2, 64
    CHW = 3 * S * S
    st = stream_ptr()
    mask = ops.box_mask(S, S // 2)
    op = ops.InpaintingOperator(0.05, mask)
    x, noise = _mk(B, 3, S, S, seed=1), _mk(B, 3, S, S, seed=2)
    y_ref, yflat_ref = op.forward(x, flatten=True, noise=noise)
    y = torch.empty(B, 3, S, S, device="cuda")
    x_c, noise_c, mask_c = x.cuda(), noise.cuda(), mask[0].cuda().contiguous()
    check(lib.kdip_inpaint_forward(ptr(x_c), ptr(noise_c), ptr(mask_c), 0.05, ptr(y), B, CHW, st))
    assert torch.allclose(y.cpu(), y_ref, rtol=0, atol=1e-7)
    # flatten / transpose(flatten): index ops are bit-exact
    idx = torch.nonzero(mask[0].flatten() > 0).flatten().to(torch.int32).cuda()
    M = idx.numel()
    yf = torch.empty(B, M, device="cuda")
    check(lib.kdip_gather(ptr(y), ptr(idx), ptr(yf), B, CHW, M, st))
    assert torch.equal(yf.cpu(), y.cpu()[..., torch.where(mask > 0)[-3], torch.where(mask > 0)[-2], torch.where(mask > 0)[-1]])
    back = torch.empty(B, CHW, device="cuda")
    check(lib.kdip_scatter(ptr(yf), ptr(idx), ptr(back), B, CHW, M, st))
    assert torch.equal(back.cpu().view(B, 3, S, S), op.transpose(yf.cpu(), flatten=True))
    # closed-form mat (condition.py:323)
    x0 = _mk(B, 3, S, S, seed=3)
    theta = torch.tensor([0.37, 0.9])
    mat = torch.empty(B, 3, S, S, device="cuda")
    x0_c, theta_c = x0.cuda(), theta.cuda()
    check(lib.kdip_inpaint_mat_scalar(ptr(y), ptr(x0_c), ptr(mask_c), ptr(theta_c), 0.05, ptr(mat), B, CHW, st))
    ref = (mask * y.cpu() - mask * x0) / (torch.tensor(0.05).pow(2) + theta[:, None, None, None])
    assert torch.allclose(mat.cpu(), ref, rtol=1e-6, atol=1e-7)
    assert torch.equal(mat.cpu()[mask.expand(B, -1, -1, -1) == 0], torch.zeros(int((mask == 0).sum()) * B))
    # combine (condition.py:131,156): clip(x0 + coef*(c_in*g + direct), -1, 1)
    g, d = _mk(B, 3, S, S, seed=4), _mk(B, 3, S, S, seed=5)
    coef, cin = torch.tensor([0.3, 1.7]), torch.tensor([0.5, 0.25])
    hat = torch.empty(B, 3, S, S, device="cuda")
    g_c, d_c, coef_c, cin_c = g.cuda(), d.cuda(), coef.cuda(), cin.cuda()
    check(lib.kdip_guidance_combine(ptr(x0_c), ptr(g_c), ptr(d_c), ptr(coef_c), ptr(cin_c), ptr(hat), B, CHW, st))
    ref = (x0 + coef[:, None, None, None] * (cin[:, None, None, None] * g + d)).clip(-1, 1)
    assert torch.allclose(hat.cpu(), ref, rtol=1e-6, atol=1e-6)


def test_alignment_and_shape_errors():
    from kdip._lib import check, lib, ptr, stream_ptr
    x = torch.zeros(1030, device="cuda")
    with pytest.raises(ValueError):
        check(lib.kdip_churn(ctypes.c_void_p(x.data_ptr() + 4), ptr(x), 1.0, 1.0, 2.0, 1024, stream_ptr()))
    with pytest.raises(ValueError):
        check(lib.kdip_churn(ptr(x), ptr(x), 1.0, 1.0, 2.0, 1030 - 4 + 1, stream_ptr()))
