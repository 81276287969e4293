"""Oracle: functional torch-CPU fp32 restatement of the ADM UNet (test infrastructure only).

Follows guided_diffusion/unet.py:143-668 (ResBlock :237-257, AttentionBlock :301-307,
QKVAttentionLegacy :339-356, UNetModel.forward :636-668), guided_diffusion/nn.py:17-19,103-121
(GroupNorm32, timestep_embedding) and the construction recipe guided_diffusion/script_util.py:130-184
with the defaults of condition/diffpir_utils/utils_model.py:353-387.

The model is described by a flat *block plan* (list of dicts) derived from the hyper-parameters and a
state_dict with exactly the reference's key names, so the same dict loads into the reference's
``UNetModel`` (checked by tests/golden/make_golden.py) and into the CUDA weight packer.
"""
import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass
class UNetConfig:
    """Hyper-parameters (utils_model.py:353-387 defaults; script_util.py:148-164 channel_mult/attention)."""
    image_size: int = 256
    num_channels: int = 128
    num_res_blocks: int = 1
    attention_resolutions: str = "16"
    num_head_channels: int = 64
    channel_mult: tuple = ()
    in_channels: int = 3
    out_channels: int = 6          # learn_sigma=True -> 6 (script_util.py:168)
    resblock_updown: bool = True
    use_scale_shift_norm: bool = True

    def resolved_channel_mult(self):
        if self.channel_mult:
            return tuple(self.channel_mult)
        return {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4),
                128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}[self.image_size]

    def attention_ds(self):
        return tuple(self.image_size // int(r) for r in self.attention_resolutions.split(","))


def ffhq_config():
    """configs/test_ffhq.json:13-17."""
    return UNetConfig(256, 128, 1, "16")


def imagenet_config():
    """configs/test_imagenet.json:12-16."""
    return UNetConfig(256, 256, 2, "32,16,8")


def tiny_config():
    """Small geometry used for golden vectors (64x64, channels 64..256, attention at 16x16 and 8x8)."""
    return UNetConfig(64, 64, 1, "16,8")


def block_plan(cfg: UNetConfig):
    """Flat list of blocks in execution order, mirroring UNetModel.__init__ (unet.py:482-618).

    Each entry: {'prefix': state-dict prefix, 'kind': 'conv_in'|'res'|'attn', 'cin', 'cout', 'updown',
    'stage': 'in'|'mid'|'out', 'block': index of the enclosing TimestepEmbedSequential}.
    """
    mult = cfg.resolved_channel_mult()
    att = cfg.attention_ds()
    mc = cfg.num_channels
    plan = []
    ch = int(mult[0] * mc)
    plan.append(dict(prefix="input_blocks.0.0", kind="conv_in", cin=cfg.in_channels, cout=ch,
                     stage="in", block=0))
    chans = [ch]
    ds = 1
    bi = 1
    for level, m in enumerate(mult):
        for _ in range(cfg.num_res_blocks):
            cout = int(m * mc)
            plan.append(dict(prefix=f"input_blocks.{bi}.0", kind="res", cin=ch, cout=cout, updown=None,
                             stage="in", block=bi))
            ch = cout
            if ds in att:
                plan.append(dict(prefix=f"input_blocks.{bi}.1", kind="attn", cin=ch, cout=ch,
                                 stage="in", block=bi))
            chans.append(ch)
            bi += 1
        if level != len(mult) - 1:
            plan.append(dict(prefix=f"input_blocks.{bi}.0", kind="res", cin=ch, cout=ch, updown="down",
                             stage="in", block=bi))
            chans.append(ch)
            bi += 1
            ds *= 2
    plan.append(dict(prefix="middle_block.0", kind="res", cin=ch, cout=ch, updown=None, stage="mid", block=0))
    plan.append(dict(prefix="middle_block.1", kind="attn", cin=ch, cout=ch, stage="mid", block=0))
    plan.append(dict(prefix="middle_block.2", kind="res", cin=ch, cout=ch, updown=None, stage="mid", block=0))
    bo = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            cout = int(mc * m)
            li = 0
            plan.append(dict(prefix=f"output_blocks.{bo}.{li}", kind="res", cin=ch + ich, cout=cout,
                             updown=None, stage="out", block=bo, skip_ch=ich))
            li += 1
            ch = cout
            if ds in att:
                plan.append(dict(prefix=f"output_blocks.{bo}.{li}", kind="attn", cin=ch, cout=ch,
                                 stage="out", block=bo))
                li += 1
            if level and i == cfg.num_res_blocks:
                plan.append(dict(prefix=f"output_blocks.{bo}.{li}", kind="res", cin=ch, cout=ch, updown="up",
                                 stage="out", block=bo))
                ds //= 2
            bo += 1
    return plan


def param_shapes(cfg: UNetConfig):
    """Ordered {key: shape} with the reference's state_dict names."""
    mc = cfg.num_channels
    ted = 4 * mc
    shapes = {}
    shapes["time_embed.0.weight"] = (ted, mc)
    shapes["time_embed.0.bias"] = (ted,)
    shapes["time_embed.2.weight"] = (ted, ted)
    shapes["time_embed.2.bias"] = (ted,)
    for b in block_plan(cfg):
        p = b["prefix"]
        if b["kind"] == "conv_in":
            shapes[p + ".weight"] = (b["cout"], b["cin"], 3, 3)
            shapes[p + ".bias"] = (b["cout"],)
        elif b["kind"] == "res":
            ci, co = b["cin"], b["cout"]
            shapes[p + ".in_layers.0.weight"] = (ci,)
            shapes[p + ".in_layers.0.bias"] = (ci,)
            shapes[p + ".in_layers.2.weight"] = (co, ci, 3, 3)
            shapes[p + ".in_layers.2.bias"] = (co,)
            shapes[p + ".emb_layers.1.weight"] = (2 * co, ted)
            shapes[p + ".emb_layers.1.bias"] = (2 * co,)
            shapes[p + ".out_layers.0.weight"] = (co,)
            shapes[p + ".out_layers.0.bias"] = (co,)
            shapes[p + ".out_layers.3.weight"] = (co, co, 3, 3)
            shapes[p + ".out_layers.3.bias"] = (co,)
            if ci != co:
                shapes[p + ".skip_connection.weight"] = (co, ci, 1, 1)
                shapes[p + ".skip_connection.bias"] = (co,)
        else:
            c = b["cin"]
            shapes[p + ".norm.weight"] = (c,)
            shapes[p + ".norm.bias"] = (c,)
            shapes[p + ".qkv.weight"] = (3 * c, c, 1)
            shapes[p + ".qkv.bias"] = (3 * c,)
            shapes[p + ".proj_out.weight"] = (c, c, 1)
            shapes[p + ".proj_out.bias"] = (c,)
    shapes["out.0.weight"] = (int(cfg.resolved_channel_mult()[0] * mc),)
    shapes["out.0.bias"] = (int(cfg.resolved_channel_mult()[0] * mc),)
    shapes["out.2.weight"] = (cfg.out_channels, int(cfg.resolved_channel_mult()[0] * mc), 3, 3)
    shapes["out.2.bias"] = (cfg.out_channels,)
    return shapes


def init_state_dict(cfg: UNetConfig, seed=0):
    """Synthetic weights (no checkpoints offline).  Every tensor is non-zero — in
    particular the reference's ``zero_module`` tensors (unet.py:210,295,617) — so no branch is vacuous.
    Deterministic for a given torch version (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        is_norm = (".in_layers.0." in k or ".out_layers.0." in k or ".norm." in k or k.startswith("out.0."))
        if is_norm:
            if k.endswith("weight"):
                sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
            else:
                sd[k] = 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            if len(shp) > 1:
                for s in shp[1:]:
                    fan_in *= s
                a = 1.0 / math.sqrt(fan_in)
                sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * a * math.sqrt(3.0) * 0.8
            else:
                sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.1
    return sd


def timestep_embedding(t, dim, max_period=10000):
    """nn.py:103-121."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, w, b):
    """GroupNorm32, nn.py:17-19,93-100: 32 groups, eps 1e-5, fp32."""
    return F.group_norm(x.float(), 32, w, b, eps=1e-5)


def _resblock(sd, p, x, emb, updown):
    """unet.py:237-257 (scale-shift norm variant, :249-253)."""
    h = F.silu(_gn(x, sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"]))
    if updown == "down":
        h = F.avg_pool2d(h, 2)
        x = F.avg_pool2d(x, 2)
    elif updown == "up":
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    scale, shift = torch.chunk(emb_out[..., None, None], 2, dim=1)
    h = _gn(h, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"]) * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def _attention(sd, p, x, head_ch):
    """unet.py:301-307 + QKVAttentionLegacy :339-356 (per-head interleaved [q,k,v], scale ch^-1/4 on q and k)."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(xf, sd[p + ".norm.weight"], sd[p + ".norm.bias"]), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    nh = c // head_ch
    q, k, v = qkv.reshape(b * nh, head_ch * 3, -1).split(head_ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(head_ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, hh * ww)
    h = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def unet_forward(sd, cfg: UNetConfig, x, t, return_feature=False, taps=None):
    """UNetModel.forward, unet.py:636-668.  ``taps`` (optional dict) collects per-block outputs."""
    emb = timestep_embedding(t, cfg.num_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    plan = block_plan(cfg)
    hs = []
    h = x.float()
    cur = ("in", 0)
    for i, b in enumerate(plan):
        key = (b["stage"], b["block"])
        if b["stage"] == "out" and b["prefix"].endswith(".0"):
            h = torch.cat([h, hs.pop()], dim=1)
        if b["kind"] == "conv_in":
            h = F.conv2d(h, sd[b["prefix"] + ".weight"], sd[b["prefix"] + ".bias"], padding=1)
        elif b["kind"] == "res":
            h = _resblock(sd, b["prefix"], h, emb, b["updown"])
        else:
            h = _attention(sd, b["prefix"], h, cfg.num_head_channels)
        if taps is not None:
            taps[b["prefix"]] = h
        last_of_block = (i + 1 == len(plan)) or ((plan[i + 1]["stage"], plan[i + 1]["block"]) != key)
        if b["stage"] == "in" and last_of_block:
            hs.append(h)
    feat = h
    out = F.conv2d(F.silu(_gn(h, sd["out.0.weight"], sd["out.0.bias"])), sd["out.2.weight"], sd["out.2.bias"], padding=1)
    if return_feature:
        return out, feat
    return out


def unet_flops(cfg: UNetConfig):
    """Forward FLOPs per image (2 x MAC), counted from the block plan (matches SURVEY.md §6 to <0.1%)."""
    res = cfg.image_size
    fl = 0
    mult = cfg.resolved_channel_mult()
    plan = block_plan(cfg)
    # walk resolution alongside the plan
    for b in plan:
        if b["kind"] == "conv_in":
            fl += 2 * res * res * b["cout"] * b["cin"] * 9
        elif b["kind"] == "res":
            r_in = res
            if b["updown"] == "down":
                res //= 2
            elif b["updown"] == "up":
                res *= 2
            fl += 2 * res * res * b["cout"] * b["cin"] * 9
            fl += 2 * res * res * b["cout"] * b["cout"] * 9
            if b["cin"] != b["cout"]:
                fl += 2 * res * res * b["cout"] * b["cin"]
            fl += 2 * 2 * b["cout"] * 4 * cfg.num_channels
        else:
            T = res * res
            c = b["cin"]
            fl += 2 * T * 3 * c * c + 2 * T * c * c + 2 * 2 * T * T * c
    c0 = int(mult[0] * cfg.num_channels)
    fl += 2 * res * res * cfg.out_channels * c0 * 9
    fl += 2 * (cfg.num_channels * 4 * cfg.num_channels + (4 * cfg.num_channels) ** 2)
    return fl
