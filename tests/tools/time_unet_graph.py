"""Experiment: UNet forward+VJP launched kernel by kernel vs replayed from a captured CUDA graph.  Usage: [B] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]
import torch
from oracle import unet_ref
from kdip.unet import UNetEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cfg = unet_ref.ffhq_config()
eng = UNetEngine(unet_ref.init_state_dict(cfg, seed=0))
x = torch.randn(B, 3, 256, 256, device="cuda"); t = torch.full((B,), 500.0, device="cuda"); seed = torch.randn(B, 6, 256, 256, device="cuda")
out = torch.empty(B, 6, 256, 256, device="cuda"); g = torch.empty(B, 3, 256, 256, device="cuda")
def step():
    eng.forward(x, t, out=out); eng.vjp(seed, out=g)
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
print(f"stream launches: {timeit(step):.2f} ms", flush=True)
ref = g.clone()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step(); step()
torch.cuda.current_stream().wait_stream(s)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
graph.replay(); torch.cuda.synchronize()
def rel(a, b):
    return ((a - b).norm() / b.norm()).item(), ((a - b).abs().max() / b.abs().max()).item()
g_graph, out_graph = g.clone(), out.clone()
step(); torch.cuda.synchronize()
g_eager2, out_eager2 = g.clone(), out.clone()
step(); torch.cuda.synchronize()
print("eager vs eager   : vjp rel-L2 %.2e max %.2e | fwd rel-L2 %.2e max %.2e" % (rel(g, g_eager2) + rel(out, out_eager2)), flush=True)
print("graph vs eager   : vjp rel-L2 %.2e max %.2e | fwd rel-L2 %.2e max %.2e" % (rel(g_graph, g_eager2) + rel(out_graph, out_eager2)), flush=True)
print(f"graph replay:    {timeit(graph.replay):.2f} ms", flush=True)
