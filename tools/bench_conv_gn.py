"""Hot-loop timing of a dgrad-style conv_gemm launch WITH the fused GroupNorm-backward reduction in its epilogue (kdip_conv_desc.gn_*),
next to the same conv without it.  Usage: python tools/bench_conv_gn.py [B] [iters] [H] [Cin] [Cout] [two]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200"), os.path.join(ROOT, "tests")]
import torch
from kdip._lib import ConvDesc, check, lib, stream_ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
Ci = int(sys.argv[4]) if len(sys.argv) > 4 else 128
Co = int(sys.argv[5]) if len(sys.argv) > 5 else 128
two = int(sys.argv[6]) if len(sys.argv) > 6 else 0
taps = 9
x = torch.randn(B, H, H, Ci, device="cuda").to(torch.bfloat16)
w = (torch.randn(taps * Co, Ci, device="cuda") / (Ci * taps) ** 0.5).to(torch.bfloat16)
out = torch.empty(B, H, H, Co, device="cuda", dtype=torch.bfloat16)
C0 = Co // 2 if two else Co
gx0 = torch.randn(B, H, H, C0, device="cuda").to(torch.bfloat16)
gx1 = torch.randn(B, H, H, Co - C0, device="cuda").to(torch.bfloat16) if two else None
ab = torch.stack([1 + 0.1 * torch.randn(B, Co, device="cuda"), 0.1 * torch.randn(B, Co, device="cuda")], -1).contiguous()
red = torch.zeros(B, Co, 2, device="cuda")
for fused in (0, 1):
    d = ConvDesc()
    d.N, d.H, d.W, d.Cout_pad, d.Cout, d.nseg = B, H, H, Co, Co, 1
    d.seg[0].act, d.seg[0].C, d.seg[0].wgt, d.seg[0].taps = x.data_ptr(), Ci, w.data_ptr(), taps
    d.out, d.out_mode, d.out_scale = out.data_ptr(), 0, 1.0
    if os.environ.get("KDIP_BENCH_XF") == "1":   # + the operand transform on the conv's input (dgrad conv1 of a two-source experiment)
        abx = torch.stack([1 + 0.1 * torch.randn(B, Ci, device="cuda"), 0.1 * torch.randn(B, Ci, device="cuda")], -1).contiguous()
        d.in_ab[0], d.in_ab_C, d.in_silu = abx.data_ptr(), Ci, 1
    if fused:
        d.gn_x0, d.gn_C0, d.gn_silu = gx0.data_ptr(), C0, 1
        d.gn_x1 = gx1.data_ptr() if two else None
        d.gn_ab, d.gn_red = ab.data_ptr(), red.data_ptr()
    plan = ctypes.c_void_p()
    check(lib.kdip_conv_plan_create(ctypes.byref(d), ctypes.byref(plan)))
    for _ in range(3):
        check(lib.kdip_conv_plan_run(plan, stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        check(lib.kdip_conv_plan_run(plan, stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * B * H * H * Co * Ci * taps
    print(f"{H}^2 {Ci}->{Co} gn_reduce={fused} two={two}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TF/s", flush=True)
    lib.kdip_conv_plan_destroy(plan)
