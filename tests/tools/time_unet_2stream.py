"""Experiment: one B-image UNet forward+VJP vs two concurrent half-batches on two CUDA streams (tensor-bound convs of one half
overlap the HBM-bound GroupNorm passes of the other).  Usage: python tests/tools/time_unet_2stream.py [B] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]
import torch
from oracle import unet_ref
from kdip.unet import UNetEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = unet_ref.ffhq_config()
sd = unet_ref.init_state_dict(cfg, seed=0)

def bufs(b):
    return (torch.randn(b, 3, 256, 256, device="cuda"), torch.full((b,), 500.0, device="cuda"), torch.randn(b, 6, 256, 256, device="cuda"),
            torch.empty(b, 6, 256, 256, device="cuda"), torch.empty(b, 3, 256, 256, device="cuda"))

def run(engs, data, streams, iters):
    for _ in range(2):
        for e, d, s in zip(engs, data, streams):
            with torch.cuda.stream(s):
                e.forward(d[0], d[1], out=d[3]); e.vjp(d[2], out=d[4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for e, d, s in zip(engs, data, streams):
            with torch.cuda.stream(s):
                e.forward(d[0], d[1], out=d[3]); e.vjp(d[2], out=d[4])
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

one = UNetEngine(sd)
t1 = run([one], [bufs(B)], [torch.cuda.current_stream()], iters)
print(f"1 stream  B={B}: {t1:.2f} ms per evaluation of {B} images", flush=True)
del one
torch.cuda.empty_cache()
engs = [UNetEngine(sd), UNetEngine(sd)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for s in streams:
    s.wait_stream(torch.cuda.current_stream())
t2 = run(engs, [bufs(B // 2), bufs(B // 2)], streams, iters)
print(f"2 streams B={B//2}+{B//2}: {t2:.2f} ms per evaluation of {B} images  ({100*(t1/t2-1):+.1f}% throughput)", flush=True)
