"""UNetModel — the ADM UNet denoiser of guided_diffusion/unet.py:398-668, executed by libkdip.

The module declares the reference's parameters (same state_dict keys and shapes, so OpenAI checkpoints load with
``load_state_dict``, sample_condition_openai.py:130-132) but has no PyTorch layers: ``forward`` hands the current
weights to the device engine (``kdip.unet.UNetEngine``: bf16 tcgen05 implicit-GEMM convolutions, fused GroupNorm /
SiLU / FiLM, flash-style attention) and autograd sees one custom Function whose backward is the hand-written
input-VJP (parameters never receive gradients on the sampling path: only ``grad(..., x)`` is taken,
condition/condition.py:136,146,155,172,269).

Supported configuration = what condition/diffpir_utils/utils_model.py:353-387 builds: resblock_updown, scale-shift norm,
head width 64, legacy attention order, fp32 I/O, unconditional.  Anything else raises at construction.
"""
import ctypes

import torch as th
import torch.nn as nn

from kdip._lib import UNetArch, check, lib
from kdip.unet import UNetEngine


def _arch(image_size, model_channels, out_channels, num_res_blocks, attention_ds, channel_mult, num_head_channels):
    a = UNetArch()
    a.image_size, a.in_channels, a.model_channels, a.out_channels = image_size, 3, model_channels, out_channels
    a.num_res_blocks, a.num_head_channels = num_res_blocks, num_head_channels
    a.n_mult = len(channel_mult)
    for i, m in enumerate(channel_mult):
        a.channel_mult[i] = float(m)
    a.n_att = len(attention_ds)
    for i, d in enumerate(attention_ds):
        a.attention_ds[i] = int(d)
    return a


def state_dict_schema(arch):
    """[(name, shape)] of the reference's UNetModel parameters, from the library's own block plan."""
    n = ctypes.c_int()
    check(lib.kdip_unet_schema_count(ctypes.byref(arch), ctypes.byref(n)))
    out = []
    name = ctypes.create_string_buffer(256)
    shape = (ctypes.c_int64 * 4)()
    nd = ctypes.c_int()
    for i in range(n.value):
        check(lib.kdip_unet_schema_entry(ctypes.byref(arch), i, name, 256, shape, ctypes.byref(nd)))
        out.append((name.value.decode(), tuple(int(shape[j]) for j in range(nd.value))))
    return out


class _UNetFn(th.autograd.Function):
    @staticmethod
    def forward(ctx, x, t, module):
        eng = module.engine()
        out = eng.forward(x, t)
        ctx.eng, ctx.token = eng, eng.forward_token
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.eng.forward_token != ctx.token:
            raise RuntimeError("UNetModel backward: the engine ran another forward since this output was produced "
                               "(activations for the VJP live in the engine workspace; one live graph at a time)")
        return ctx.eng.vjp(g.contiguous()), None, None


class UNetModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False):
        super().__init__()
        unsupported = []
        if in_channels != 3 or out_channels != 6: unsupported.append("in/out channels must be 3/6 (learn_sigma=True)")
        if num_classes is not None: unsupported.append("class conditioning")
        if not use_scale_shift_norm: unsupported.append("use_scale_shift_norm=False")
        if not resblock_updown: unsupported.append("resblock_updown=False")
        if num_head_channels != 64: unsupported.append("num_head_channels != 64")
        if use_new_attention_order: unsupported.append("use_new_attention_order=True")
        if use_fp16: unsupported.append("use_fp16=True")
        if dims != 2: unsupported.append("dims != 2")
        if unsupported:
            raise NotImplementedError("kdip UNetModel covers the guided-sampling configuration only; unsupported: "
                                      + "; ".join(unsupported))
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.dropout, self.channel_mult = dropout, tuple(channel_mult)
        self.num_classes, self.use_checkpoint, self.dtype = num_classes, use_checkpoint, th.float32
        self.num_heads, self.num_head_channels, self.num_heads_upsample = num_heads, num_head_channels, num_heads_upsample
        self._arch = _arch(image_size, model_channels, out_channels, num_res_blocks, self.attention_resolutions,
                           self.channel_mult, num_head_channels)
        # parameters under the reference's names: "input_blocks.3.0.in_layers.2.weight" -> nested containers
        self._names = []
        for name, shape in state_dict_schema(self._arch):
            parent = self
            parts = name.split(".")
            for p in parts[:-1]:
                if not hasattr(parent, p):
                    parent.add_module(p, nn.Module())
                parent = getattr(parent, p)
            w = th.empty(shape)
            if parts[-1] == "weight" and len(shape) > 1:
                nn.init.kaiming_uniform_(w, a=5 ** 0.5)
            elif parts[-1] == "weight":
                nn.init.ones_(w)
            else:
                nn.init.zeros_(w)
            parent.register_parameter(parts[-1], nn.Parameter(w, requires_grad=False))
            self._names.append(name)
        self._engine = None
        self._engine_key = None
        self.out_cov = None   # set by OpenAIDenoiserV2 to fuse its covariance head
        # engine arithmetic: "bf16" (tcgen05 fast path) or "fp32" (the reference's use_fp16=False arithmetic, CUDA cores);
        # assign model.precision = "fp32" (or export KDIP_PRECISION=fp32) before the first forward
        import os
        self.precision = os.environ.get("KDIP_PRECISION", "bf16")

    # ---- engine management ---------------------------------------------------------------------------------------
    def _weights_key(self):
        ps = list(self.parameters())
        return (ps[0].device, tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps[:4]),
                None if self.out_cov is None else tuple(p._version for p in self.out_cov), self.precision)

    def invalidate(self):
        """Drop the packed engine: call after editing weights in place through ``p.data`` (EMA-style updates do not bump
        ``Parameter._version``, which is what the engine cache keys on)."""
        self._engine = None
        self._engine_key = None

    def engine(self):
        """(Re)pack the weights into the device engine when they changed (load_state_dict / .to())."""
        key = self._weights_key()
        if self._engine is None or key != self._engine_key:
            dev = key[0]
            if dev.type != "cuda":
                raise RuntimeError("kdip UNetModel runs on CUDA only (B200, sm_100a); call .to('cuda') — there is no CPU fallback")
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith("out_cov")}
            self._engine = None
            self._engine = UNetEngine(sd, image_size=self.image_size, num_channels=self.model_channels,
                                      num_res_blocks=self.num_res_blocks,
                                      attention_resolutions=",".join(str(self.image_size // d) for d in self.attention_resolutions),
                                      num_head_channels=self.num_head_channels, channel_mult=self.channel_mult,
                                      out_cov=self.out_cov, device=dev, precision=self.precision)
            self._engine_key = key
        return self._engine

    def convert_to_fp16(self):
        raise NotImplementedError("kdip UNetModel keeps fp32 I/O with bf16 tensor-core math inside the engine")

    def convert_to_fp32(self):
        return None

    def forward(self, x, timesteps, y=None, return_feature=False):
        """unet.py:636-668.  x [N,3,S,S] fp32, timesteps [N] -> [N,6,S,S] (and the pre-head feature)."""
        assert y is None, "class-conditional models are outside the guided-sampling path"
        need_grad = th.is_grad_enabled() and x.requires_grad
        if need_grad:
            out = _UNetFn.apply(x, timesteps, self)
        else:
            out = self.engine().forward(x.detach(), timesteps)
        if return_feature:
            return out, self.engine().feature(x.shape[0])
        return out
