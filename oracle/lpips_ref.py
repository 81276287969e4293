"""TEST INFRASTRUCTURE ONLY - CPU restatement of LPIPS(net='vgg') (the `lpips` package v0.1 the reference imports,
sample_condition_openai.py:11,46,161), imported by tests/ alone.

PARITY UNPINNED for the package itself: `lpips` is a third-party dependency that is neither vendored in /root/reference nor
installed here, and its weights are a download, so there are no golden vectors of the real metric.  What IS pinned: the backbone
is torchvision's own vgg16 `features` module (an independent implementation, run here), and the head below restates the package's
published code path: ScalingLayer ((x - shift) / scale), taps after relu1_2 / 2_2 / 3_3 / 4_3 / 5_3, normalize_tensor
(x / (sqrt(sum_c x^2) + 1e-10)), (f0 - f1)^2, NetLinLayer (1x1 conv, no bias), spatial_average, sum over the taps."""
import torch

SHIFT = torch.tensor([-0.030, -0.088, -0.188]).view(1, 3, 1, 1)
SCALE = torch.tensor([0.458, 0.448, 0.450]).view(1, 3, 1, 1)
TAP_AFTER = [3, 8, 15, 22, 29]          # indices of relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 in vgg16.features
CHNS = [64, 128, 256, 512, 512]


def synthetic_state_dict(seed=0):
    """A state_dict in the lpips package's layout with seeded random weights (He-scaled convolutions so the activations keep
    their scale through 13 layers; non-negative lin weights as the trained ones are)."""
    import torchvision
    g = torch.Generator().manual_seed(seed)
    feats = torchvision.models.vgg16(weights=None).features
    sd = {}
    slice_of = lambda i: 1 if i < 4 else 2 if i < 9 else 3 if i < 16 else 4 if i < 23 else 5
    for i, m in enumerate(feats):
        if isinstance(m, torch.nn.Conv2d):
            fan_in = m.in_channels * 9
            sd[f"net.slice{slice_of(i)}.{i}.weight"] = torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5
            sd[f"net.slice{slice_of(i)}.{i}.bias"] = torch.randn(m.bias.shape, generator=g) * 0.05
    for k, c in enumerate(CHNS):
        sd[f"lin{k}.model.1.weight"] = torch.rand(1, c, 1, 1, generator=g) / c
    sd["scaling_layer.shift"] = SHIFT.clone()
    sd["scaling_layer.scale"] = SCALE.clone()
    return sd


def lpips_ref(sd, in0, in1, normalize=False, bf16_operands=False):
    """-> ([N,1,1,1], per-layer list).  bf16_operands: round the conv inputs and weights to bf16 (the product's operand precision)
    so that a kernel bug is not hidden behind the bf16 tolerance."""
    import torchvision
    feats = torchvision.models.vgg16(weights=None).features.eval()
    with torch.no_grad():
        for i, m in enumerate(feats):
            if isinstance(m, torch.nn.Conv2d):
                s = 1 if i < 4 else 2 if i < 9 else 3 if i < 16 else 4 if i < 23 else 5
                w = sd[f"net.slice{s}.{i}.weight"]
                m.weight.copy_(w.bfloat16().float() if (bf16_operands and i > 0) else w)
                m.bias.copy_(sd[f"net.slice{s}.{i}.bias"])
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        outs = []
        for x in (in0, in1):
            h = (x - SHIFT) / SCALE
            taps = []
            for i, m in enumerate(feats[:30]):
                if bf16_operands and isinstance(m, torch.nn.Conv2d) and i > 0:
                    h = h.bfloat16().float()
                h = m(h)
                if bf16_operands and isinstance(m, torch.nn.ReLU):
                    h = h.bfloat16().float()          # activations are stored as bf16
                if i in TAP_AFTER:
                    taps.append(h)
            outs.append(taps)
        per = []
        for k in range(5):
            f0, f1 = outs[0][k], outs[1][k]
            n0 = f0 / (f0.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            n1 = f1 / (f1.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            d = (n0 - n1) ** 2
            w = sd[f"lin{k}.model.1.weight"].view(1, -1, 1, 1)
            per.append((d * w).sum(1, keepdim=True).mean((2, 3), keepdim=True))
        return sum(per), per
