"""The fused entry point kdip_guided_eval called straight through the C ABI: raw ctypes on libkdip.so with hand-declared structs,
no import of the kdip Python package (torch only owns device memory) - what a maintainer of the reference binds.
Checked against the reference's own golden output of a guided evaluation (PiGDM, Gaussian deblur, tiny UNet)."""
import ctypes
import os

import numpy as np
import pytest
import torch

import inputs as I

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t


class Arch(ctypes.Structure):
    _fields_ = [("image_size", c_int), ("in_channels", c_int), ("model_channels", c_int), ("out_channels", c_int),
                ("num_res_blocks", c_int), ("num_head_channels", c_int), ("n_mult", c_int), ("channel_mult", c_float * 8),
                ("n_att", c_int), ("attention_ds", c_int * 8), ("precision", c_int)]


class OpDesc(ctypes.Structure):
    _fields_ = [("kind", c_int), ("S", c_int), ("sf", c_int), ("sigma_s", c_float), ("psf", c_void_p), ("ksize", c_int),
                ("mask", c_void_p), ("rs_w", c_void_p), ("rs_idx", c_void_p), ("rs_taps", c_int)]


class Pmv(ctypes.Structure):
    _fields_ = [(n, c_float) for n in ("c_in", "recip", "recipm1", "min_log", "max_log", "post_var", "coef1_sq")]


class Cfg(ctypes.Structure):
    _fields_ = [("guidance", c_int), ("sigma", c_float), ("t_model", c_float), ("theta", c_float), ("zeta", c_float), ("sc", Pmv)]


@pytest.mark.parametrize("precision", [1, 0], ids=["fp32", "bf16"])
def test_guided_eval_through_raw_ctypes(precision, golden_small):
    from oracle import diffusion_ref, operators_ref, unet_ref       # test infrastructure: weights, PSF and schedule constants
    lib = ctypes.CDLL(os.path.join(ROOT, "k-diffusion-inverse-problems_b200", "kdip", "libkdip.so"))
    lib.kdip_last_error.restype = ctypes.c_char_p

    def ok(rc):
        assert rc == 0, lib.kdip_last_error().decode()
    dev = torch.device("cuda", 0)
    cfg = unet_ref.tiny_config()
    sd = {k: v.to(dev).contiguous() for k, v in unet_ref.init_state_dict(cfg, seed=0).items()}
    arch = Arch()
    arch.image_size, arch.in_channels, arch.model_channels, arch.out_channels = 64, 3, 64, 6
    arch.num_res_blocks, arch.num_head_channels, arch.precision = 1, 64, precision
    mult, att = cfg.resolved_channel_mult(), cfg.attention_ds()
    arch.n_mult, arch.n_att = len(mult), len(att)
    for i, m in enumerate(mult):
        arch.channel_mult[i] = float(m)
    for i, d in enumerate(att):
        arch.attention_ds[i] = int(d)
    names = list(sd)
    unet = c_void_p()
    ok(lib.kdip_unet_create(ctypes.byref(arch), len(names), (ctypes.c_char_p * len(names))(*[n.encode() for n in names]),
                            (c_void_p * len(names))(*[sd[n].data_ptr() for n in names]),
                            (ctypes.c_int64 * len(names))(*[sd[n].numel() for n in names]), ctypes.byref(unet)))
    psf = np.ascontiguousarray(operators_ref.gaussian_psf().numpy(), dtype=np.float32)
    od = OpDesc()
    od.kind, od.S, od.sf, od.sigma_s, od.psf, od.ksize = 1, 64, 1, 0.05, psf.ctypes.data, psf.shape[-1]
    op = c_void_p()
    ok(lib.kdip_op_create(ctypes.byref(od), ctypes.byref(op)))
    try:
        # measurement exactly as the golden generator drew it: y = A x0 + sigma_s * randn (CPU generator, seed 2)
        x0 = I.image(64, batch=1, seed=1).to(dev)
        torch.manual_seed(2)
        noise = torch.randn(1, 3, 64, 64).to(dev)
        y = torch.empty(1, 3, 64, 64, device=dev)
        n = c_size_t()
        ok(lib.kdip_guided_eval_workspace_bytes(unet, op, 1, ctypes.byref(n)))
        ws = torch.empty(n.value + 256, dtype=torch.uint8, device=dev)
        wsp = c_void_p(ws.data_ptr() + (-ws.data_ptr()) % 256)
        ok(lib.kdip_op_forward(op, c_void_p(x0.data_ptr()), c_void_p(noise.data_ptr()), c_void_p(y.data_ptr()), 1, wsp, c_size_t(n.value), None))
        sched = diffusion_ref.Schedule()
        for sigma in (10.0, 0.1):
            t = int(sched.sigma_to_t(torch.tensor([sigma])).long())
            c = Cfg()
            c.guidance, c.sigma, c.t_model, c.theta, c.zeta = 2, sigma, float(t), float(np.float32(sigma) ** 2 / (1 + np.float32(sigma) ** 2)), 0.0
            c.sc.c_in = float(1.0 / np.sqrt(np.float32(sigma) ** 2 + np.float32(1)))
            c.sc.recip, c.sc.recipm1 = float(sched.sqrt_recip_alphas_cumprod[t]), float(sched.sqrt_recipm1_alphas_cumprod[t])
            c.sc.min_log, c.sc.max_log = float(sched.posterior_log_variance_clipped[t]), float(sched.log_betas[t])
            c.sc.post_var, c.sc.coef1_sq = float(sched.posterior_variance[t]), float(np.float32(sched.posterior_mean_coef1[t]) ** 2)
            xt = I.xt(64, sigma, seed=21).to(dev)
            hat = torch.empty_like(xt)
            ok(lib.kdip_guided_eval(unet, op, ctypes.byref(c), c_void_p(xt.data_ptr()), c_void_p(y.data_ptr()), c_void_p(hat.data_ptr()), 1,
                                    wsp, c_size_t(n.value), None))
            torch.cuda.synchronize()
            gold = torch.as_tensor(golden_small[f"guid.gaussian_blur.pgdm.pgdm.{sigma}"])
            l2 = ((hat.cpu() - gold).norm() / gold.norm()).item()
            print(f"raw C ABI kdip_guided_eval precision={precision} sigma={sigma}: rel-L2 {l2:.3e}")
            assert l2 <= (2e-3 if precision == 1 else (3e-2 if sigma <= 1.5 else 0.6))
    finally:
        lib.kdip_op_destroy(op)
        lib.kdip_unet_destroy(unet)
