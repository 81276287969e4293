"""Device time of one guided model evaluation (ConditionOpenAIDenoiser.forward) at a closed-form sigma and at a CG sigma, next to
the bare UNet forward + VJP: the difference is the guidance / operator / solver overhead per evaluation.  Usage: [B] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]
import torch
from condition.condition import ConditionOpenAIDenoiser
from condition.diffpir_utils.utils_model import create_argparser
from condition.measurements import get_operator
from guided_diffusion.script_util import args_to_dict, create_model_and_diffusion, model_and_diffusion_defaults
from kdip.synth import synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
margs = create_argparser({"num_channels": 128, "num_res_blocks": 1, "attention_resolutions": "16"}).parse_args([])
model, diffusion = create_model_and_diffusion(**args_to_dict(margs, model_and_diffusion_defaults().keys()))
model.load_state_dict(synthetic_state_dict(model, seed=0))
model = model.eval().to(dev)
op = get_operator(name="gaussian_blur", in_shape=(1, 3, 256, 256), kernel_size=61, intensity=3.0, sigma_s=0.05, device=dev)
x0 = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
y = op.forward(x0, flatten=True)[0]
for guidance, cov in (("I", "convert"), ("pgdm", "pgdm")):
    cm = ConditionOpenAIDenoiser(inner_model=model, diffusion=diffusion, x0_cov_type=cov, recon_mse=None, operator=op,
                                 measurement=(y, y.reshape(B, -1)), guidance=guidance, mle_sigma_thres=0.2, device=dev).eval()
    for sigma in (1.5, 0.1):
        xt = x0 + sigma * torch.randn_like(x0)
        sg = torch.full((B,), sigma, device=dev)
        for _ in range(2):
            cm(xt, sg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            cm(xt, sg)
        e1.record()
        torch.cuda.synchronize()
        extra = getattr(getattr(cm.operator, "handle", None), "last_cg_iters", None)
        print(f"guidance={guidance} cov={cov} sigma={sigma}: {e0.elapsed_time(e1)/iters:.2f} ms per guided eval (B={B}) cg_iters={extra}", flush=True)
eng = model.engine()
xs = torch.randn(B, 3, 256, 256, device=dev); tt = torch.full((B,), 338.0, device=dev); sd6 = torch.randn(B, 6, 256, 256, device=dev)
eng.profile(xs, tt, sd6)
pr = eng.profile(xs, tt, sd6)
print("bare UNet fwd+vjp %.2f ms (conv %.2f ms, %.0f TF/s)" % (pr["total_ms"], pr["conv_ms"], pr["conv_flops"] / pr["conv_ms"] / 1e9))
