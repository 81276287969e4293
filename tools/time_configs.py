"""Device time of one guided model evaluation for the BASELINE.json configs (FFHQ UNet, synthetic weights) at a sigma above and
below the MLE threshold.  Usage: python tools/time_configs.py [B] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]
import torch
import k_diffusion as K
from condition.condition import ConditionOpenAIDenoiser, ConditionOpenAIDenoiserV2
from condition.diffpir_utils.utils_model import create_argparser
from condition.measurements import get_operator
from guided_diffusion.script_util import args_to_dict, create_model_and_diffusion, model_and_diffusion_defaults
from k_diffusion.external import OpenAIDenoiserV2
from kdip.synth import synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
margs = create_argparser({"num_channels": 128, "num_res_blocks": 1, "attention_resolutions": "16"}).parse_args([])
model, diffusion = create_model_and_diffusion(**args_to_dict(margs, model_and_diffusion_defaults().keys()))
model.load_state_dict(synthetic_state_dict(model, seed=0))
model = model.eval().to(dev)
x0 = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
sig100 = K.sampling.get_sigmas_karras(100, 0.01, 80.0, rho=7.0)
recon = {"sigmas": sig100[:-1].clone(), "mse_list": 0.5 * sig100[:-1] ** 2 / (1 + sig100[:-1] ** 2)}


def timed(cm, sigma):
    xt = x0 + sigma * torch.randn_like(x0)
    sg = torch.full((B,), sigma, device=dev)
    for _ in range(4):      # the engine captures its CUDA graphs on the third call of a shape
        cm(xt, sg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = cm(xt, sg)
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    return e0.elapsed_time(e1) / iters


cases = [
    ("cfg1 inpainting box / pgdm", dict(name="inpainting", sigma_s=0.05, mask_opt=dict(mask_type="box", mask_len_range=(128, 129), image_size=256)), "pgdm", "pgdm", {}),
    ("cfg2 gaussian deblur / I convert", dict(name="gaussian_blur", in_shape=(1, 3, 256, 256), kernel_size=61, intensity=3.0, sigma_s=0.05), "I", "convert", {}),
    ("cfg3 SR x4 / I analytic", dict(name="super_resolution", in_shape=(1, 3, 256, 256), scale_factor=4, sigma_s=0.05), "I", "analytic", {}),
    ("cfg4-like motion deblur / dps", dict(name="motion_blur", in_shape=(1, 3, 256, 256), kernel_size=61, intensity=0.5, sigma_s=0.05), "dps", "dps", {"zeta": 1.0}),
]
for tag, okw, guidance, cov, extra in cases:
    op = get_operator(device=dev, **okw)
    y = op.forward(x0, flatten=True)
    cm = ConditionOpenAIDenoiser(inner_model=model, diffusion=diffusion, x0_cov_type=cov, recon_mse=dict(recon), operator=op,
                                 measurement=y, guidance=guidance, mle_sigma_thres=0.2, device=dev, **extra).eval()
    print(f"{tag:36s}: sigma 1.5 {timed(cm, 1.5):7.2f} ms   sigma 0.1 {timed(cm, 0.1):7.2f} ms   (B={B})", flush=True)
# cfg5: v2 denoiser, type II, DWT
for ot in ("dwt", "dct"):
    den = OpenAIDenoiserV2(model, diffusion, device=dev, ortho_tf_type=ot).to(dev)
    op = get_operator(device=dev, name="gaussian_blur", in_shape=(1, 3, 256, 256), kernel_size=61, intensity=3.0, sigma_s=0.05)
    y = op.forward(x0, flatten=True)
    cm = ConditionOpenAIDenoiserV2(denoiser=den, operator=op, measurement=y, guidance="II", device=dev, mle_sigma_thres=1.0, ortho_tf_type=ot).eval()
    print(f"{'cfg5 v2 gaussian deblur / II ' + ot:36s}: sigma 1.5 {timed(cm, 1.5):7.2f} ms   sigma 0.5 {timed(cm, 0.5):7.2f} ms   (B={B}, UNet forward only) cg_iters={max(getattr(op.handle, 'last_cg_iters', None) or [0])}", flush=True)
