"""How reproducible is the reference's own arithmetic on the parity cases?  (CPU oracle only; developer tool, not a test.)

    python tests/tools/sensitivity.py > profiles/r2_reference_sensitivity.txt

Perturbs the WEIGHTS of the CPU oracle (fp32 arithmetic throughout) by a relative eps * N(0,1) and reports how far every tiny-UNet
guided evaluation and sampler trajectory of tests/golden/inputs.py moves from the reference's golden output: eps = 6e-8 is one fp32
ulp (what a different but equally valid fp32 summation order amounts to), eps = 4e-6 the accuracy of a 3 x bf16 operand split.
This is what fixes the tolerances of tests/test_guidance_gpu.py and the choice of an fp32-FMA engine for the tight mode."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from oracle import guidance_ref, operators_ref as ops, sampler_ref, unet_ref  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))
cfg = unet_ref.tiny_config()
sd = unet_ref.init_state_dict(cfg, seed=0)


def pert(eps, seed=0):
    g = torch.Generator().manual_seed(100 + seed)
    return {k: v * (1 + eps * torch.randn(v.shape, generator=g)) for k, v in sd.items()}


def make_ref(name, size=64):
    return {"gaussian_blur": lambda: ops.BlurOperator("gaussian_blur", 0.05, in_shape=(1, 3, size, size)),
            "motion_blur": lambda: ops.BlurOperator("motion_blur", 0.05, intensity=0.5, in_shape=(1, 3, size, size)),
            "super_resolution": lambda: ops.SuperResolutionOperator(0.05, 4, in_shape=(1, 3, size, size)),
            "inpainting": lambda: ops.InpaintingOperator(0.05, ops.box_mask(size, size // 2))}[name]()


def meas(name):
    ref, x0 = make_ref(name), I.image(64, batch=1, seed=1)
    torch.manual_seed(2)
    noise = torch.randn_like(ref.down_sample(x0)).contiguous() if name == "super_resolution" else torch.randn(*x0.shape)
    return ref, ref.forward(x0, flatten=True, noise=noise)


def recon():
    s = sampler_ref.get_sigmas_karras(100, 0.01, 80)
    return {"sigmas": s[:-1].clone(), "mse_list": 0.5 * s[:-1] ** 2 / (1 + s[:-1] ** 2)}


def l2(a, b):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    return ((a - b).norm() / b.norm()).item(), (a - b).abs().max().item()


for eps in (0.0, 6e-8, 4e-6):
    sdp = pert(eps) if eps else sd
    for (opname, guid, cov, sigma, extra) in I.GUIDANCE_COMBOS:
        ref, m = meas(opname)
        cm = guidance_ref.ConditionDenoiserRef(sdp, cfg, ref, m, guid, cov, recon_mse=recon(), mle_sigma_thres=0.2, **extra)
        out = cm(I.xt(64, sigma, seed=21), torch.tensor([sigma]))
        print(f"eps={eps:g} {opname}/{guid}/{cov}/{sigma}: rel-L2, max = {l2(out, G[f'guid.{opname}.{guid}.{cov}.{sigma}'])}", flush=True)
    for (tag, opname, guid, cov, sampler, n, churn) in I.SAMPLER_RUNS:
        ref, m = meas(opname)
        cm = guidance_ref.ConditionDenoiserRef(sdp, cfg, ref, m, guid, cov)
        torch.manual_seed(5)
        noises = [torch.randn(1, 3, 64, 64) for _ in range(n)]
        kw = dict(s_churn=80, s_tmin=0.05, s_tmax=50, s_noise=1.003) if churn else {}
        rfn = sampler_ref.sample_euler if sampler == "euler" else sampler_ref.sample_heun
        out = rfn(cm, I.xT(64, seed=3), sampler_ref.get_sigmas_karras(n, 0.01, 80), noise_fn=lambda i, x: noises[i], **kw)
        print(f"eps={eps:g} trajectory {tag}: rel-L2, max = {l2(out, G['traj.' + tag])}", flush=True)
