"""GPU debug: tiny UNet input-VJP vs the fp32 oracle under several call patterns."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200"), os.path.join(ROOT, "tests", "golden")]
import torch
import inputs as I
from oracle import unet_ref
from kdip.unet import UNetEngine

cfg = unet_ref.tiny_config()
sd = unet_ref.init_state_dict(cfg, seed=0)
eng = UNetEngine(sd, image_size=64, num_channels=64, num_res_blocks=1, attention_resolutions="16,8")


def oracle(x, t, seed, scale=None):
    xs = x if scale is None else x * scale[:, None, None, None]
    xs = xs.detach().requires_grad_()
    out = unet_ref.unet_forward(sd, cfg, xs, t)
    (g,) = torch.autograd.grad((out * seed).sum(), xs)
    return out.detach(), g


def report(tag, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    d = (got - ref).abs()
    print(f"{tag}: max-rel {d.max() / ref.abs().max():.3e} l2-rel {(got - ref).norm() / ref.norm():.3e}  | ref absmax {ref.abs().max():.3e}")
    if (got - ref).norm() / ref.norm() > 5e-2:
        print("   per-image l2:", [((got[b] - ref[b]).norm() / ref[b].norm()).item() for b in range(got.shape[0])])
        print("   per-channel l2:", [((got[:, c] - ref[:, c]).norm() / ref[:, c].norm()).item() for c in range(got.shape[1])])
        rows = d.amax((0, 1, 3))
        cols = d.amax((0, 1, 2))
        print("   worst rows:", rows.topk(5).indices.tolist(), "worst cols:", cols.topk(5).indices.tolist())


def run(tag, x, t, seed, scale=None, junk=False):
    out = eng.forward(x.cuda(), t.cuda(), x_scale=None if scale is None else scale.cuda())
    if junk:
        z = [torch.randn(64, 1024, 1024, device="cuda") for _ in range(4)]
        torch.cuda.synchronize()
    g = eng.vjp(seed.cuda())
    o_ref, g_ref = oracle(x, t, seed, scale)
    report(tag + " fwd", out, o_ref)
    report(tag + " vjp", g, g_ref)


x = I.unet_input(64, batch=2, seed=11)
t = torch.tensor([37, 801])
seed = I.unet_seed((2, 6, 64, 64), seed=12)
run("A direct", x, t, seed)
run("B x_scale", x * 3, t, seed, scale=torch.tensor([0.5, 0.25]))
s0 = seed.clone(); s0[:, 3:] = 0
run("C zero var-seed", x, t, s0)
run("D t=338", x, torch.tensor([338, 338]), seed)
run("E junk between", x, t, seed, junk=True)
xt = I.xt(64, 1.5, seed=21, batch=2)
run("F xt sigma1.5 c_in", xt, torch.tensor([338, 338]), s0, scale=torch.tensor([1 / (1.5 ** 2 + 1) ** 0.5] * 2))
xt = I.xt(64, 10.0, seed=21, batch=2)
run("G xt sigma10 c_in", xt, torch.tensor([673, 673]), s0, scale=torch.tensor([1 / (10.0 ** 2 + 1) ** 0.5] * 2))
# smooth seed (like mat): low-pass noise
sm = torch.nn.functional.avg_pool2d(seed, 9, 1, 4)
sm[:, 3:] = 0
run("H smooth seed", xt, torch.tensor([673, 673]), sm, scale=torch.tensor([1 / (10.0 ** 2 + 1) ** 0.5] * 2))
run("I batch1", x[:1], t[:1], seed[:1])
run("J batch3", torch.cat([x, x[:1]]), torch.tensor([37, 801, 37]), torch.cat([seed, seed[:1]]))
run("K back to 2", x, t, seed)

# module + autograd path
from guided_diffusion.unet import UNetModel
model = UNetModel(image_size=64, in_channels=3, model_channels=64, out_channels=6, num_res_blocks=1,
                  attention_resolutions=cfg.attention_ds(), channel_mult=cfg.resolved_channel_mult(), num_head_channels=64,
                  use_scale_shift_norm=True, resblock_updown=True)
model.load_state_dict(sd, strict=True)
model = model.eval().cuda()
xg = x.cuda().requires_grad_()
out = model(xg, t.cuda())
(gx,) = torch.autograd.grad((out * seed.cuda()).sum(), xg)
o_ref, g_ref = oracle(x, t, seed)
report("M module fwd", out.detach(), o_ref)
report("M module autograd vjp", gx, g_ref)
e2 = model.engine()
o2 = e2.forward(x.cuda(), t.cuda())
g2 = e2.vjp(seed.cuda())
report("N module engine direct fwd", o2, o_ref)
report("N module engine direct vjp", g2, g_ref)
