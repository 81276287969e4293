"""CPU side of LPIPS: the oracle restatement's invariants (the `lpips` package is absent: parity with it is unpinned, see
oracle/lpips_ref.py) and the weight-layout parser of kdip/lpips.py."""
import pytest
import torch


def test_lpips_oracle_invariants():
    from oracle import lpips_ref as L
    sd = L.synthetic_state_dict(3)
    g = torch.Generator().manual_seed(4)
    a = torch.rand(1, 3, 64, 64, generator=g)
    b = (a + 0.2 * torch.randn(1, 3, 64, 64, generator=g)).clip(0, 1)
    v_ab, per = L.lpips_ref(sd, a, b)
    v_ba, _ = L.lpips_ref(sd, b, a)
    assert v_ab.shape == (1, 1, 1, 1) and v_ab.item() > 0
    assert torch.allclose(v_ab, v_ba, rtol=1e-5)                       # symmetric
    assert torch.allclose(sum(per), v_ab)
    assert L.lpips_ref(sd, a, a)[0].abs().max().item() == 0.0
    # normalize=True is the [0,1] -> [-1,1] map in front of the scaling layer (lpips.LPIPS.forward)
    assert torch.allclose(L.lpips_ref(sd, a, b, normalize=True)[0], L.lpips_ref(sd, 2 * a - 1, 2 * b - 1)[0])
    # head restated independently: unit-normalised features, weighted squared distance, mean over pixels
    import torchvision
    feats = torchvision.models.vgg16(weights=None).features.eval()
    with torch.no_grad():
        for i, m in enumerate(feats):
            if isinstance(m, torch.nn.Conv2d):
                s = 1 if i < 4 else 2 if i < 9 else 3 if i < 16 else 4 if i < 23 else 5
                m.weight.copy_(sd[f"net.slice{s}.{i}.weight"]); m.bias.copy_(sd[f"net.slice{s}.{i}.bias"])
        f = lambda x: feats[:4]((x - L.SHIFT) / L.SCALE)               # relu1_2
        fa, fb = f(a)[0].flatten(1).T, f(b)[0].flatten(1).T            # [HW, 64]
        ua = fa / (fa.norm(dim=1, keepdim=True) + 1e-10)
        ub = fb / (fb.norm(dim=1, keepdim=True) + 1e-10)
        l0 = (((ua - ub) ** 2) @ sd["lin0.model.1.weight"].flatten()).mean()
    assert torch.allclose(l0, per[0].flatten()[0], rtol=1e-4)


def test_lpips_weight_layouts():
    from oracle import lpips_ref as L
    from kdip.lpips import CHNS, VGG_CONVS, split_state_dict
    sd = L.synthetic_state_dict(1)
    convs, lins, shift, scale = split_state_dict(sd)
    assert sorted(convs) == VGG_CONVS and [w.numel() for w in lins] == CHNS
    assert convs[0][0].shape == (64, 3, 3, 3) and convs[28][0].shape == (512, 512, 3, 3)
    tv = {}
    for k, v in sd.items():
        if k.startswith("net.slice"):
            tv["features." + k.split(".", 2)[2]] = v
        elif k.startswith("lin"):
            tv[k] = v
    convs2, lins2, shift2, _ = split_state_dict(tv)
    assert all(torch.equal(convs[i][0], convs2[i][0]) for i in VGG_CONVS) and torch.allclose(shift2, shift.flatten())
    bad = dict(sd)
    del bad["lin3.model.1.weight"]
    with pytest.raises(KeyError):
        split_state_dict(bad)
