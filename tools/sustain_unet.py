"""Sustained (power-capped) timing of the FFHQ UNet forward + input-VJP: back-to-back evaluations for [seconds], no host sync in
between, the first third discarded (the 1 kW cap takes a second or two to pull the clocks down; 0.2 s bursts such as
tools/time_unet.py 32 5 run at boost clocks and flatter power-hungry variants).  Usage: python tools/sustain_unet.py [B] [seconds]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]
import torch
from condition.diffpir_utils.utils_model import create_argparser
from guided_diffusion.script_util import args_to_dict, create_model_and_diffusion, model_and_diffusion_defaults
from kdip.synth import synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 9.0
margs = create_argparser({"num_channels": 128, "num_res_blocks": 1, "attention_resolutions": "16"}).parse_args([])
model, _ = create_model_and_diffusion(**args_to_dict(margs, model_and_diffusion_defaults().keys()))
model.load_state_dict(synthetic_state_dict(model, seed=0))
eng = model.eval().cuda().engine()
x = torch.randn(B, 3, 256, 256, device="cuda")
t = torch.full((B,), 500.0, device="cuda")
seed = torch.randn(B, 6, 256, 256, device="cuda")
out = torch.empty(B, 6, 256, 256, device="cuda")
g = torch.empty(B, 3, 256, 256, device="cuda")
for _ in range(4):
    eng.forward(x, t, out=out); eng.vjp(seed, out=g)
torch.cuda.synchronize()
n = max(12, int(secs / 0.037))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ev[0].record()
for i in range(n):
    eng.forward(x, t, out=out); eng.vjp(seed, out=g)
    ev[i + 1].record()
torch.cuda.synchronize()
k = n // 3
late = ev[k].elapsed_time(ev[n]) / (n - k)
early = ev[0].elapsed_time(ev[k]) / k
print(f"B={B}: sustained fwd+vjp {late:.2f} ms per evaluation over the last {n - k} of {n} (first {k}: {early:.2f} ms) -> {B / late * 1000 / 199:.3f} img/s @199 evals", flush=True)
