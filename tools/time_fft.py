"""Device time of the spectral operator entry points at the bench batch (CUDA events, back-to-back calls through the product API):
A x, A^T y, the closed-form mat, one CG solve.  Also the ncu target for the FFT kernels.  Usage: [B] [iters] [operator]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]
import torch
from condition.measurements import get_operator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
name = sys.argv[3] if len(sys.argv) > 3 else "gaussian_blur"
dev = torch.device("cuda", 0)
op = get_operator(name=name, in_shape=(1, 3, 256, 256), kernel_size=61, intensity=3.0 if name == "gaussian_blur" else 0.5, sigma_s=0.05, device=dev)
h = op.handle
g = torch.Generator().manual_seed(0)
x0 = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).to(dev)
y = h.forward(x0, torch.randn(B, 3, 256, 256, generator=g).to(dev))
theta = torch.full((B,), 0.3, device=dev)
tmap = (torch.rand(B, 3, 256, 256, generator=g) * 0.05 + 1e-4).to(dev)
PLANE = 3 * 256 * 256 * 4


def timeit(label, planes, fn, reps=iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{label:40s} {us:9.1f} us   {planes * B * PLANE / us / 1e3:8.1f} GB/s algorithmic ({planes} planes)", flush=True)
    return us


timeit("A x", 2, lambda: h.forward(x0, None))
timeit("A^T y", 2, lambda: h.transpose(y))
timeit("mat closed form", 3, lambda: h.mat_closed(y, x0, theta))
timeit("dps_grad", 3, lambda: h.dps_grad(y, x0))
for ot in (None, "dwt", "dct"):
    h.mat_cg(y, x0, tmap, ot=ot)
    its = max(h.last_cg_iters)
    us = timeit(f"CG solve ot={ot} ({its} it)", 8 * its, lambda: h.mat_cg(y, x0, tmap, ot=ot), reps=3)
    print(f"    -> {us / its:.1f} us per iteration")
