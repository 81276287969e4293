"""ImageNet 256x256 ADM UNet (configs[3]: 256 channels, 2 res-blocks, attention at 32/16/8) forward + VJP timing. Usage: [B] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]
import torch
from condition.diffpir_utils.utils_model import create_argparser
from guided_diffusion.script_util import args_to_dict, create_model_and_diffusion, model_and_diffusion_defaults
from kdip.synth import synthetic_state_dict

UNET_FWD_FLOPS = 2239.67e9     # ImageNet 256x256 UNet forward, per image (2 x MAC)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
margs = create_argparser({"num_channels": 256, "num_res_blocks": 2, "attention_resolutions": "32,16,8"}).parse_args([])
model, _ = create_model_and_diffusion(**args_to_dict(margs, model_and_diffusion_defaults().keys()))
model.load_state_dict(synthetic_state_dict(model, seed=0))
print("params M", sum(v.numel() for v in model.state_dict().values()) / 1e6, "GF fwd/img", UNET_FWD_FLOPS / 1e9, flush=True)
eng = model.eval().cuda().engine()
print("workspace GB", eng.workspace_bytes(B) / 1e9, flush=True)
x = torch.randn(B, 3, 256, 256, device="cuda"); t = torch.full((B,), 500.0, device="cuda"); seed = torch.randn(B, 6, 256, 256, device="cuda")
out = torch.empty(B, 6, 256, 256, device="cuda"); g = torch.empty(B, 3, 256, 256, device="cuda")
for _ in range(4):     # the engine captures its CUDA graphs on the third call of a shape
    eng.forward(x, t, out=out); eng.vjp(seed, out=g)
torch.cuda.synchronize()
assert torch.isfinite(out).all() and torch.isfinite(g).all()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
for _ in range(iters):
    e[0].record(); eng.forward(x, t, out=out); e[1].record(); eng.vjp(seed, out=g); e[2].record()
    torch.cuda.synchronize()
    tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
tf /= iters; tb /= iters
fl = UNET_FWD_FLOPS * B
print(f"ImageNet UNet B={B}: fwd {tf:.2f} ms ({fl/tf/1e9:.1f} TF/s)  vjp {tb:.2f} ms ({fl/tb/1e9:.1f} TF/s)")
pr = eng.profile(x, t, seed)
print("profile: total %.1f ms conv %.1f ms (%.0f TF/s) other %.1f ms" % (pr["total_ms"], pr["conv_ms"], pr["conv_flops"] / pr["conv_ms"] / 1e9, pr["other_ms"]))
