"""LPIPS (net='vgg') on the libkdip kernels - the perceptual metric of sample_condition_openai.py:46,161
(``loss_fn_vgg = lpips.LPIPS(net='vgg').to(device)``; ``loss_fn_vgg(to_eval(x0), to_eval(hat_x0))[0, 0, 0, 0].item()``).

The `lpips` package (v0.1, richzhang/PerceptualSimilarity) is a third-party dependency of the reference and is not vendored; its
published algorithm is restated here (and, on the CPU, in oracle/lpips_ref.py): ScalingLayer -> torchvision VGG16 `features`
with taps after relu1_2 / relu2_2 / relu3_3 / relu4_3 / relu5_3 -> unit-normalise over channels -> squared difference ->
non-negative 1x1 "lin" weights -> spatial mean -> sum over the five taps.  The thirteen 3x3 convolutions run on the tcgen05
implicit-GEMM kernel (bf16 operands, fp32 accumulation); ReLU, max pooling and the per-tap reduction are csrc/lpips.cu.

Weights: the package downloads them; here they are read from a state_dict in the package's own layout
(``net.slice{1..5}.{torchvision index}.weight / .bias``, ``lin{0..4}.model.1.weight``, optional ``scaling_layer.shift / .scale``)
or in torchvision's (``features.{index}.*`` / ``{index}.*``) plus the lin weights.  Without weights the constructor raises - a
perceptual metric on random features would be silently meaningless.  No CPU fallback."""
import ctypes
import os

import torch

from ._lib import ConvDesc, check, lib, ptr, stream_ptr

VGG_CONVS = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]       # torchvision vgg16.features indices of the convolutions
VGG_POOL_BEFORE = {5, 10, 17, 24}                                   # MaxPool2d(2, 2) sits in front of these convolutions
VGG_TAPS = {2: 0, 7: 1, 14: 2, 21: 3, 28: 4}                        # conv index -> feature tap (after its ReLU)
VGG_SLICE = {0: 1, 2: 1, 5: 2, 7: 2, 10: 3, 12: 3, 14: 3, 17: 4, 19: 4, 21: 4, 24: 5, 26: 5, 28: 5}
CHNS = [64, 128, 256, 512, 512]
SHIFT = (-0.030, -0.088, -0.188)                                    # lpips ScalingLayer
SCALE = (0.458, 0.448, 0.450)


def _pad_rows(n):
    return n if n in (16, 32) else (n + 63) // 64 * 64


def split_state_dict(sd):
    """-> ({conv index: (weight OIHW, bias)}, [lin weight [C] x 5], shift, scale) from a lpips- or torchvision-layout state_dict."""
    convs, lins = {}, [None] * 5
    for i in VGG_CONVS:
        for pre in (f"net.slice{VGG_SLICE[i]}.{i}", f"features.{i}", f"{i}"):
            if f"{pre}.weight" in sd:
                convs[i] = (sd[f"{pre}.weight"].detach().float(), sd[f"{pre}.bias"].detach().float())
                break
        else:
            raise KeyError(f"LPIPS weights: VGG16 convolution {i} not found (tried net.slice{VGG_SLICE[i]}.{i}.*, features.{i}.*, {i}.*)")
    for k in range(5):
        for key in (f"lin{k}.model.1.weight", f"lins.{k}.model.1.weight", f"lin{k}.weight"):
            if key in sd:
                lins[k] = sd[key].detach().float().reshape(-1)
                break
        else:
            raise KeyError(f"LPIPS weights: lin{k}.model.1.weight not found")
        if lins[k].numel() != CHNS[k]:
            raise ValueError(f"LPIPS weights: lin{k} has {lins[k].numel()} channels, expected {CHNS[k]}")
    shift = sd["scaling_layer.shift"].detach().float().reshape(3) if "scaling_layer.shift" in sd else torch.tensor(SHIFT)
    scale = sd["scaling_layer.scale"].detach().float().reshape(3) if "scaling_layer.scale" in sd else torch.tensor(SCALE)
    return convs, lins, shift, scale


class LPIPS:
    """Callable like ``lpips.LPIPS(net='vgg')``: ``loss(in0, in1, normalize=False) -> [N, 1, 1, 1]`` (fp32, on the inputs' device)."""

    def __init__(self, net="vgg", state_dict=None, weights=None, device="cuda"):
        if net != "vgg":
            raise NotImplementedError("kdip LPIPS implements net='vgg' (the one the reference's scripts use)")
        if not torch.cuda.is_available():
            raise RuntimeError("kdip LPIPS needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if state_dict is None:
            path = weights or os.environ.get("KDIP_LPIPS_WEIGHTS")
            if not path or not os.path.exists(path):
                raise FileNotFoundError(
                    "kdip LPIPS needs the lpips(net='vgg') weights: pass state_dict=lpips.LPIPS(net='vgg').state_dict(), or weights= / "
                    "KDIP_LPIPS_WEIGHTS= a file saved with torch.save(that state_dict) (the package downloads them; this one does not)")
            state_dict = torch.load(path, map_location="cpu")
        self.device = torch.device(device)
        convs, lins, shift, scale = split_state_dict(state_dict)
        dev = self.device
        self.shift = shift.to(dev).view(1, 3, 1, 1)
        self.scale = scale.to(dev).view(1, 3, 1, 1)
        self.lins = [w.to(dev).contiguous() for w in lins]
        self.w0 = convs[0][0].to(dev).contiguous()              # conv1_1 (3 -> 64): the small-Cin direct conv
        self.b0 = convs[0][1].to(dev).contiguous()
        self.w0_scratch = torch.empty(9 * 3 * 64, device=dev)
        self.packed = {}
        for i in VGG_CONVS[1:]:
            w, b = convs[i]
            O, Iin = w.shape[0], w.shape[1]
            rp, cp = _pad_rows(O), (Iin + 63) // 64 * 64
            dst = torch.empty(9 * rp, cp, dtype=torch.bfloat16, device=dev)
            with torch.cuda.device(dev):
                check(lib.kdip_pack_conv_weight(ptr(w.to(dev).contiguous()), O, Iin, 9, rp, cp, 0, ptr(dst), stream_ptr()))
            self.packed[i] = (dst, b.to(dev).contiguous(), O, Iin)
        self._plans = {}

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError("kdip LPIPS lives on the device it was created on")
        return self

    def eval(self):
        return self

    def _conv(self, i, x, N, H, W):
        """3x3 conv + bias on bf16 NHWC through a cached kdip_conv_plan (keyed by the tensors it was encoded for)."""
        wp, b, O, Iin = self.packed[i]
        out = torch.empty(N, H, W, O, dtype=torch.bfloat16, device=x.device)
        d = ConvDesc()
        d.N, d.H, d.W, d.Cout_pad, d.Cout, d.nseg = N, H, W, O, O, 1
        d.seg[0].act, d.seg[0].C, d.seg[0].wgt, d.seg[0].taps = x.data_ptr(), Iin, wp.data_ptr(), 9
        d.bias = b.data_ptr()
        d.out, d.out_mode, d.out_scale = out.data_ptr(), 0, 1.0
        plan = ctypes.c_void_p()
        check(lib.kdip_conv_plan_create(ctypes.byref(d), ctypes.byref(plan)))
        try:
            check(lib.kdip_conv_plan_run(plan, stream_ptr()))
        finally:
            lib.kdip_conv_plan_destroy(plan)     # the launch has copied the plan's parameters
        return out

    def features(self, x):
        """x [N,3,H,W] fp32 (already scaled) -> the five taps, bf16 NHWC."""
        N, _, H, W = x.shape
        if H % 16 or W % 16 or H != W or H < 128:
            raise ValueError("kdip LPIPS: square images with a side that is a multiple of 16 and at least 128 (got %dx%d)" % (H, W))
        h = torch.empty(N, H, W, 64, dtype=torch.bfloat16, device=x.device)
        check(lib.kdip_layer_conv_small_cin(ptr(x), None, ptr(self.w0), ptr(self.b0), N, 64, 3, 0, H, W, ptr(self.w0_scratch), ptr(h), stream_ptr()))
        check(lib.kdip_relu_bf16(ptr(h), h.numel(), stream_ptr()))
        taps = []
        for i in VGG_CONVS[1:]:
            if i in VGG_POOL_BEFORE:
                C = h.shape[-1]
                p = torch.empty(N, H // 2, W // 2, C, dtype=torch.bfloat16, device=x.device)
                check(lib.kdip_maxpool2_bf16(ptr(h), ptr(p), N, H, W, C, stream_ptr()))
                h, H, W = p, H // 2, W // 2
            h = self._conv(i, h, N, H, W)
            check(lib.kdip_relu_bf16(ptr(h), h.numel(), stream_ptr()))
            if i in VGG_TAPS:
                taps.append((h, H * W))
        return taps

    def __call__(self, in0, in1, retPerLayer=False, normalize=False):
        assert in0.is_cuda and in1.is_cuda, "kdip LPIPS: CUDA tensors only (no CPU fallback)"
        if in0.dim() == 3:
            in0, in1 = in0[None], in1[None]           # the reference passes [3,H,W] images (to_eval(x)[0])
        in0, in1 = in0.float(), in1.float()
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        N = in0.shape[0]
        with torch.cuda.device(in0.device):
            x0 = ((in0 - self.shift) / self.scale).contiguous()
            x1 = ((in1 - self.shift) / self.scale).contiguous()
            f0, f1 = self.features(x0), self.features(x1)
            per_layer = []
            total = torch.zeros(N, dtype=torch.float64, device=in0.device)
            for k in range(5):
                acc = torch.zeros(N, dtype=torch.float64, device=in0.device)
                (a, hw), (b, _) = f0[k], f1[k]
                check(lib.kdip_lpips_layer(ptr(a), ptr(b), ptr(self.lins[k]), N, hw, CHNS[k], ptr(acc), stream_ptr()))
                per_layer.append(acc.float().view(N, 1, 1, 1))
                total = total + acc
        val = total.float().view(N, 1, 1, 1)
        return (val, per_layer) if retPerLayer else val

    forward = __call__
