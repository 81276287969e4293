"""Device time of the attention layer kernels at the ImageNet UNet's 32x32 level (T=1024, 8 heads) and others; streamed tcgen05
kernels (attention_tcs.cu) vs the CUDA-core kernels (KDIP_ATTN_TCS=0).  Usage: python tools/time_attn.py [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]
import torch
from kdip._lib import check, lib, ptr, stream_ptr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for T, heads in ((1024, 8), (512, 8), (4096, 4)):
    C = heads * 64
    n = N if T <= 1024 else max(1, N // 8)
    qkv = torch.randn(n, T, 3 * C, device="cuda").to(torch.bfloat16)
    do = torch.randn(n, T, C, device="cuda").to(torch.bfloat16)
    res = {}
    for mode in ("1", "0"):
        os.environ["KDIP_ATTN_TCS"] = mode
        out = torch.empty(n, T, C, dtype=torch.bfloat16, device="cuda")
        lse = torch.empty(n, heads, T, device="cuda")
        dqkv = torch.empty(n, T, 3 * C, dtype=torch.bfloat16, device="cuda")
        st = stream_ptr()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        iters = 5
        for i in range(iters + 2):
            ev[0].record()
            check(lib.kdip_layer_attention_fwd(ptr(qkv), n, T, heads, ptr(out), ptr(lse), st))
            ev[1].record()
            check(lib.kdip_layer_attention_bwd(ptr(qkv), ptr(out), ptr(do), ptr(lse), n, T, heads, ptr(dqkv), st))
            ev[2].record()
            torch.cuda.synchronize()
            if i >= 2:
                tf += ev[0].elapsed_time(ev[1]) / iters
                tb += ev[1].elapsed_time(ev[2]) / iters
        res[mode] = (out.float(), dqkv.float(), tf, tb)
        fl = 4.0 * T * T * 64 * heads * n
        print(f"T={T} heads={heads} N={n} {'tcgen05 streamed' if mode == '1' else 'CUDA cores':>16}: fwd {tf*1e3:8.1f} us ({fl/tf/1e9:6.1f} TF/s)  "
              f"bwd {tb*1e3:8.1f} us ({2.5*fl/tb/1e9:6.1f} TF/s)", flush=True)
    eo = ((res["1"][0] - res["0"][0]).norm() / res["0"][0].norm()).item()
    eg = ((res["1"][1] - res["0"][1]).norm() / res["0"][1].norm()).item()
    print(f"   tcgen05 vs CUDA-core results: out rel-L2 {eo:.2e}  dqkv rel-L2 {eg:.2e}")
os.environ.pop("KDIP_ATTN_TCS", None)
