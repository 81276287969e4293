"""LPIPS(net='vgg') on the libkdip kernels (kdip/lpips.py, csrc/lpips.cu) against the CPU restatement oracle/lpips_ref.py
(torchvision's own VGG16 module + the lpips package's published head).  Tolerances: the three small kernels are exact (ReLU, max
pool) or fp32-accurate (per-tap reduction, 1e-5); the whole metric carries the bf16 operand rounding of thirteen convolutions:
relative 5e-3 against the fp32 oracle (measured on the B200: 5e-4 .. 7e-4), 1e-3 against the oracle run on bf16-rounded operands
(measured 3e-5 .. 1.4e-4)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _imgs(n, s, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(n, 3, s, s, generator=g)
    b = (a + 0.15 * torch.randn(n, 3, s, s, generator=g)).clip(0, 1)
    return a, b


def test_relu_and_maxpool_bit_exact():
    from kdip._lib import check, lib, ptr, stream_ptr
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(2, 16, 32, 72, device="cuda", generator=g).to(torch.bfloat16)           # NHWC
    y = x.clone()
    check(lib.kdip_relu_bf16(ptr(y), y.numel(), stream_ptr()))
    assert torch.equal(y, torch.relu(x))
    p = torch.empty(2, 8, 16, 72, dtype=torch.bfloat16, device="cuda")
    check(lib.kdip_maxpool2_bf16(ptr(x), ptr(p), 2, 16, 32, 72, stream_ptr()))
    ref = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2).float(), 2, 2).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(p, ref)
    with pytest.raises(ValueError):
        check(lib.kdip_maxpool2_bf16(ptr(x), ptr(p), 2, 15, 32, 72, stream_ptr()))


@pytest.mark.parametrize("C", [64, 256, 512])
def test_lpips_layer_reduction(C):
    from kdip._lib import check, lib, ptr, stream_ptr
    g = torch.Generator(device="cuda").manual_seed(C)
    N, HW = 3, 37 * 5
    f0 = torch.relu(torch.randn(N, HW, C, device="cuda", generator=g)).to(torch.bfloat16)
    f1 = torch.relu(torch.randn(N, HW, C, device="cuda", generator=g)).to(torch.bfloat16)
    f0[0, 3] = 0                                             # an all-zero feature vector: 0 / (0 + 1e-10) = 0, no NaN
    w = torch.rand(C, device="cuda", generator=g)
    out = torch.zeros(N, dtype=torch.float64, device="cuda")
    check(lib.kdip_lpips_layer(ptr(f0), ptr(f1), ptr(w), N, HW, C, ptr(out), stream_ptr()))
    a, b = f0.double(), f1.double()
    na = a / (a.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    nb = b / (b.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    ref = (((na - nb) ** 2) * w.double()).sum(-1).mean(-1)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-9), (out, ref)


@pytest.mark.parametrize("size,batch", [(128, 2), (256, 1)])
def test_lpips_vgg_matches_oracle(size, batch):
    from oracle import lpips_ref as L
    from kdip.lpips import LPIPS
    sd = L.synthetic_state_dict(0)
    a, b = _imgs(batch, size, 5)
    ref, ref_per = L.lpips_ref(sd, a, b)
    ref16, _ = L.lpips_ref(sd, a, b, bf16_operands=True)
    loss = LPIPS(net="vgg", state_dict=sd)
    got, per = loss(a.cuda(), b.cuda(), retPerLayer=True)
    assert got.shape == (batch, 1, 1, 1)
    rel = ((got.cpu() - ref).abs() / ref.abs()).max().item()
    rel16 = ((got.cpu() - ref16).abs() / ref16.abs()).max().item()
    print(f"LPIPS {size}^2 x{batch}: {got.flatten().tolist()} vs oracle {ref.flatten().tolist()}: rel {rel:.3e}, vs bf16-operand oracle {rel16:.3e}")
    assert rel < 5e-3 and rel16 < 1e-3, (rel, rel16)
    for k in range(5):
        r = ((per[k].cpu() - ref_per[k]).abs() / ref_per[k].abs()).max().item()
        assert r < 1e-2, (k, r)
    # the reference's call shape: one [3,H,W] image pair, scalar read with [0,0,0,0]; identical images -> 0; torchvision-layout weights
    one = loss(a[0].cuda(), b[0].cuda())[0, 0, 0, 0].item()
    assert abs(one - got[0, 0, 0, 0].item()) <= 1e-3 * abs(one)
    assert loss(a.cuda(), a.cuda()).abs().max().item() == 0.0
    tv = {k.replace(f"net.slice{s}.", "features."): v for k, v in sd.items() for s in range(1, 6) if k.startswith(f"net.slice{s}.")}
    tv.update({k: v for k, v in sd.items() if k.startswith("lin")})
    got_tv = LPIPS(net="vgg", state_dict=tv)(a.cuda(), b.cuda())
    assert torch.equal(got_tv, got)


def test_lpips_needs_weights(monkeypatch):
    import lpips
    monkeypatch.delenv("KDIP_LPIPS_WEIGHTS", raising=False)
    with pytest.raises(FileNotFoundError):
        lpips.LPIPS(net="vgg")
    with pytest.raises(NotImplementedError):
        lpips.LPIPS(net="alex", state_dict={})
