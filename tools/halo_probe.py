"""Probe: halo-pipeline conv vs torch, with and without the descriptor base offset (KDIP_HALO_BASEOFF)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200"), os.path.join(ROOT, "tests")]
import torch, torch.nn.functional as F
from gpu_util import pack_weight, run_conv, to_nchw_f32, to_nhwc_bf16, relerr
N, H, W, Ci, Co = 1, 4, 128, 64, 64
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(N, Ci, H, W, device="cuda", generator=g)
w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (Ci * 9) ** 0.5
bf = lambda t: t.to(torch.bfloat16).float()
ref = F.conv2d(bf(x), bf(w), padding=1)
for bo in ("1", "0"):
    os.environ["KDIP_HALO_BASEOFF"] = bo
    got = to_nchw_f32(run_conv([(to_nhwc_bf16(x), pack_weight(w)[0], 9)], N, H, W, Co))
    bad = ((got - ref).abs() > ref.abs().max() / 100)
    print(f"base_offset={bo}: rel err {relerr(got, ref):.3e} finite={torch.isfinite(got).all().item()} bad frac {bad.float().mean().item():.3f}")
    # per-tap diagnosis: single-tap weights
    for tap in range(9):
        w1 = torch.zeros_like(w); w1[:, :, tap // 3, tap % 3] = w[:, :, tap // 3, tap % 3]
        r1 = F.conv2d(bf(x), bf(w1), padding=1)
        g1 = to_nchw_f32(run_conv([(to_nhwc_bf16(x), pack_weight(w1)[0], 9)], N, H, W, Co))
        print(f"   tap {tap}: rel err {relerr(g1, r1):.3e}")
