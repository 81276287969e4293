"""Hot-loop timing of single conv_gemm launches on UNet shapes (CUDA events, inputs > L2 at B=32).
Usage: python tools/bench_conv.py [B] [iters] ; env KDIP_CONV_PAIR / KDIP_CONV_TMAEPI / KDIP_CONV_STAGES / KDIP_CONV_BN select variants."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200"), os.path.join(ROOT, "tests")]
import torch
from kdip._lib import ConvDesc, check, lib, stream_ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
SHAPES = [  # H, Cin, Cout, taps, residual
    (256, 128, 128, 9, 1), (256, 256, 128, 9, 0), (128, 256, 256, 9, 1), (128, 128, 128, 9, 0), (64, 256, 256, 9, 1),
    (64, 512, 256, 9, 0), (32, 512, 512, 9, 0), (16, 512, 512, 9, 1), (256, 128, 128, 1, 0), (256, 128, 128, 9, 0),
    (8, 512, 512, 9, 1), (16, 1024, 512, 9, 0), (32, 256, 256, 9, 1), (64, 128, 128, 9, 0), (32, 768, 256, 9, 0),
]
only = os.environ.get("KDIP_BENCH_SHAPES")
if only:
    SHAPES = [SHAPES[int(i)] for i in only.split(",")]
for (H, Ci, Co, taps, res) in SHAPES:
    x = torch.randn(B, H, H, Ci, device="cuda").to(torch.bfloat16)
    w = (torch.randn(taps * Co, Ci, device="cuda") / (Ci * taps) ** 0.5).to(torch.bfloat16)
    r = torch.randn(B, H, H, Co, device="cuda").to(torch.bfloat16) if res else None
    out = torch.empty(B, H, H, Co, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(Co, device="cuda")
    d = ConvDesc()
    d.N, d.H, d.W, d.Cout_pad, d.Cout, d.nseg = B, H, H, Co, Co, 1
    d.seg[0].act, d.seg[0].C, d.seg[0].wgt, d.seg[0].taps = x.data_ptr(), Ci, w.data_ptr(), taps
    d.bias = bias.data_ptr()
    if res:
        d.residual, d.res_mode = r.data_ptr(), 1
    d.out, d.out_mode, d.out_scale = out.data_ptr(), 0, 1.0
    if os.environ.get("KDIP_BENCH_XF") == "1" and taps == 9 and H >= 128:   # fused GroupNorm apply on the operand path
        ab = torch.stack([1 + 0.1 * torch.randn(B, Ci, device="cuda"), 0.1 * torch.randn(B, Ci, device="cuda")], -1).contiguous()
        d.in_ab[0], d.in_ab_C, d.in_silu = ab.data_ptr(), Ci, 1
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
    print(f"{H:4d}^2 {Ci:4d}->{Co:4d} taps{taps} res{res}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TF/s", flush=True)
    lib.kdip_conv_plan_destroy(plan)
