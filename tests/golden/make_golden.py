"""Generate golden vectors by EXECUTING THE REFERENCE (/root/reference) on CPU.  Build container only.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference ships no tests or fixtures; these vectors are outputs of its own unmodified code
(imported through oracle/refshim.py) on seeded synthetic inputs and weights.  tests/test_oracle_golden.py pins
the oracle restatement against them; the GPU parity tests then compare the CUDA path with the oracle.
Inputs are regenerated from seeds by ``tests/golden/inputs.py`` (shared with the tests), so only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import refshim  # noqa: E402
from oracle import unet_ref  # noqa: E402
import inputs as I  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    mods = refshim.install()
    CC, CM, K = mods["CC"], mods["CM"], mods["K"]
    torch.set_grad_enabled(True)
    G = {}

    # ---- 0. data fixtures -----------------------------------------------------------------------
    g_ref = np.load("condition/kernels/gaussian_ks61_std3.0.npy")
    from oracle.operators_ref import gaussian_psf
    assert np.array_equal(torch.Tensor(g_ref).numpy(), gaussian_psf().numpy()), "gaussian PSF regeneration differs"

    # ---- 1. tiny UNet ---------------------------------------------------------------------------
    cfg = unet_ref.tiny_config()
    model, diffusion = refshim.build_reference_unet(mods, dict(
        image_size=64, num_channels=64, num_res_blocks=1, attention_resolutions="16,8"))
    sd = unet_ref.init_state_dict(cfg, seed=0)
    model.load_state_dict(sd, strict=True)            # same key names & shapes as the reference module
    G["tiny.sd_checksum"] = np.array([float(sum(v.double().sum() for v in sd.values()))])
    x = I.unet_input(64, batch=2, seed=11).requires_grad_()
    t = torch.tensor([37, 801])
    out, feat = model(x, t, return_feature=True)
    v = I.unet_seed(out.shape, seed=12)
    (gx,) = torch.autograd.grad((out * v).sum(), x)
    G["tiny.out"], G["tiny.feat_mean"], G["tiny.vjp"] = out.detach().numpy(), feat.detach().mean((2, 3)).numpy(), gx.numpy()
    tf = torch.tensor([12.25, 640.5])               # fractional t (v2 path, external.py:163)
    G["tiny.out_fract"] = model(x.detach(), tf).detach().numpy()

    # ---- 2. schedule ----------------------------------------------------------------------------
    den = K.external.OpenAIDenoiser(model, diffusion)
    sig = I.SIGMA_PROBE
    G["sched.sigma_to_t"] = den.sigma_to_t(sig).numpy()
    c_out, c_in = den.get_scalings(sig)
    G["sched.c_in"] = c_in.numpy()
    for name in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                 "posterior_log_variance_clipped", "posterior_mean_coef1", "betas"):
        G["sched." + name] = getattr(diffusion, name)
    G["sched.karras100"] = K.sampling.get_sigmas_karras(100, 0.01, 80, rho=7.).numpy()
    xs = I.unet_input(64, batch=1, seed=13)
    pm = diffusion.p_mean_variance(model, xs, torch.tensor([55]))
    G["pmv.pred_xstart"], G["pmv.variance"] = pm["pred_xstart"].detach().numpy(), pm["variance"].detach().numpy()

    # ---- 3. operators (64x64 and 256x256) ---------------------------------------------------------
    for size in (64, 256):
        x0 = I.image(size, batch=1, seed=1)
        # 256x256 outputs are stored 4x4-subsampled (I.sub) to keep the fixture small; 64x64 ones in full
        sub = (lambda a: I.sub(a)) if size == 256 else (lambda a: a)
        opers = {
            "gaussian_blur": CM.get_operator(name="gaussian_blur", in_shape=(1, 3, size, size), kernel_size=61,
                                             intensity=3.0, sigma_s=0.05, device="cpu"),
            "motion_blur": CM.get_operator(name="motion_blur", in_shape=(1, 3, size, size), kernel_size=61,
                                           intensity=0.5, sigma_s=0.05, device="cpu"),
            "super_resolution": CM.get_operator(name="super_resolution", in_shape=(1, 3, size, size), scale_factor=4,
                                                sigma_s=0.05, device="cpu"),
        }
        np.random.seed(0)
        opers["inpainting"] = CM.get_operator(name="inpainting", sigma_s=0.05, device="cpu", mask_opt=dict(
            mask_type="box", mask_len_range=(size // 2, size // 2 + 1), image_size=size))
        np.random.seed(7)
        op_rand = CM.get_operator(name="inpainting", sigma_s=0.05, device="cpu", mask_opt=dict(
            mask_type="random", mask_prob_range=(0.5, 0.5), image_size=size))
        G[f"op{size}.random_mask"] = np.packbits(op_rand.mask.numpy().astype(np.uint8)[0, 0])
        for name, op in opers.items():
            torch.manual_seed(2)
            y, yf = op.forward(x0.clone(), flatten=True)
            G[f"op{size}.{name}.y"] = sub(y.numpy())
            G[f"op{size}.{name}.yflat_sum"] = np.array([yf.double().sum().item(), yf.shape[1]])
            G[f"op{size}.{name}.y_noiseless"] = sub(op.forward(x0.clone(), noiseless=True).numpy())
            G[f"op{size}.{name}.At_y"] = sub(op.transpose(y).numpy())
            # ---- 4. mat solvers ----
            solver = CC.__MAT_SOLVER__[name]
            op.forward(x0.clone(), flatten=True)        # pre_calculated as in sample_condition_openai.py:169
            xm = I.image(size, batch=1, seed=4) * 0.8
            G[f"mat{size}.{name}.scalar"] = sub(solver(op, y, xm, torch.tensor([0.37])).numpy())
            if size == 64:
                th = I.theta_map(size, seed=5)
                for ot in (None, "dct", "dwt"):
                    G[f"mat{size}.{name}.cg.{ot}"] = solver(op, y, xm, th, CC.OrthoTransform(ot)).numpy()
        if size == 64:
            xx = I.image(size, batch=1, seed=6)
            for ot in ("dct", "dwt"):
                W = CC.OrthoTransform(ot)
                G[f"ot.{ot}.fwd"] = W(xx).numpy()
                G[f"ot.{ot}.inv"] = W.inv(xx).numpy()

        # ---- 5. guided evals with the tiny UNet ----------------------------------------------------
        if size != 64:
            continue
        sigmas100 = K.sampling.get_sigmas_karras(100, 0.01, 80, rho=7.)
        recon = lambda: {"sigmas": sigmas100[:-1].clone(), "mse_list": 0.5 * sigmas100[:-1] ** 2 / (1 + sigmas100[:-1] ** 2)}
        combos = I.GUIDANCE_COMBOS
        for (opname, guidance, cov, sigma, extra) in combos:
            op = opers[opname]
            torch.manual_seed(2)
            meas = op.forward(x0.clone(), flatten=True)
            cm = CC.ConditionOpenAIDenoiser(inner_model=model, diffusion=diffusion, x0_cov_type=cov,
                                            recon_mse=recon(), operator=op, measurement=meas, guidance=guidance,
                                            device="cpu", mle_sigma_thres=0.2, **extra).eval()
            xt = I.xt(size, sigma, seed=21)
            hat = cm(xt, torch.tensor([sigma]))
            G[f"guid.{opname}.{guidance}.{cov}.{sigma}"] = hat.numpy()

        # ---- 6. sampler trajectories (tiny) ------------------------------------------------------------
        for (tag, opname, guidance, cov, sampler, n, churn) in I.SAMPLER_RUNS:
            op = opers[opname]
            torch.manual_seed(2)
            meas = op.forward(x0.clone(), flatten=True)
            cm = CC.ConditionOpenAIDenoiser(inner_model=model, diffusion=diffusion, x0_cov_type=cov, recon_mse=None,
                                            operator=op, measurement=meas, guidance=guidance, device="cpu").eval()
            sig = K.sampling.get_sigmas_karras(n, 0.01, 80, rho=7.)
            xT = I.xT(size, seed=3)
            fn = K.sampling.sample_euler if sampler == "euler" else K.sampling.sample_heun
            torch.manual_seed(5)
            kw = dict(s_churn=80, s_tmin=0.05, s_tmax=50, s_noise=1.003) if churn else {}
            G[f"traj.{tag}"] = fn(cm, xT, sig, disable=True, **kw).detach().numpy()

    np.savez_compressed(os.path.join(OUT, "golden_small.npz"), **G)
    print("golden_small:", len(G), "arrays", sum(v.nbytes for v in G.values()) / 1e6, "MB")

    # ---- 7. full-size FFHQ UNet: one UNet call + one guided eval (target config: gaussian blur, PiGDM) ----
    F_ = {}
    cfg = unet_ref.ffhq_config()
    model, diffusion = refshim.build_reference_unet(mods, dict(num_channels=128, num_res_blocks=1, attention_resolutions="16"))
    sd = unet_ref.init_state_dict(cfg, seed=0)
    model.load_state_dict(sd, strict=True)
    F_["ffhq.sd_checksum"] = np.array([float(sum(v.double().sum() for v in sd.values()))])
    F_["ffhq.n_params"] = np.array([sum(v.numel() for v in sd.values())])
    x0 = I.image(256, batch=1, seed=1)
    op = CM.get_operator(name="gaussian_blur", in_shape=(1, 3, 256, 256), kernel_size=61, intensity=3.0, sigma_s=0.05, device="cpu")
    torch.manual_seed(2)
    meas = op.forward(x0.clone(), flatten=True)
    cm = CC.ConditionOpenAIDenoiser(inner_model=model, diffusion=diffusion, x0_cov_type="pgdm", recon_mse=None,
                                    operator=op, measurement=meas, guidance="pgdm", device="cpu").eval()
    sigma = 1.5
    xt = I.xt(256, sigma, seed=21)
    F_["ffhq.hat_x0.pgdm"] = cm(xt, torch.tensor([sigma])).numpy()
    with torch.no_grad():
        c_in = 1 / (sigma ** 2 + 1) ** 0.5
        t = cm.denoiser.sigma_to_t(torch.tensor([sigma])).long()
        F_["ffhq.t"] = t.numpy()
        F_["ffhq.unet_out"] = model(xt * c_in, t).numpy().astype(np.float16)   # fp16 storage: parity tol is >= 1e-3
    np.savez_compressed(os.path.join(OUT, "golden_ffhq.npz"), **F_)
    print("golden_ffhq:", sum(v.nbytes for v in F_.values()) / 1e6, "MB")


if __name__ == "__main__":
    main()
