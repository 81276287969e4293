"""Device-side ADM UNet handle (libkdip kdip_unet_*), the engine behind guided_diffusion.unet.UNetModel.

Reference: guided_diffusion/unet.py:398-668 (module), guided_diffusion/script_util.py:130-184 (construction),
condition/diffpir_utils/utils_model.py:353-387 (hyper-parameter defaults).
"""
import ctypes

import torch

from ._lib import UNetArch, check, lib, ptr, stream_ptr


def channel_mult_for(image_size):
    """script_util.py:148-160."""
    return {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}[image_size]


class UNetEngine:
    """Owns the packed weights (inside libkdip) and the activation workspace (a torch uint8 tensor)."""

    def __init__(self, state_dict, image_size=256, num_channels=128, num_res_blocks=1, attention_resolutions="16",
                 num_head_channels=64, channel_mult=None, out_cov=None, device="cuda", precision="bf16"):
        """precision: "bf16" = bf16 tcgen05 operands with fp32 accumulation (the fast path); "fp32" = the reference's own fp32
        arithmetic on the CUDA cores (csrc/unet_fp32.cu), for tight-tolerance parity."""
        if not torch.cuda.is_available():
            raise RuntimeError("kdip.UNetEngine needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        mult = tuple(channel_mult) if channel_mult else channel_mult_for(image_size)
        att = tuple(image_size // int(r) for r in str(attention_resolutions).split(","))
        arch = UNetArch()
        arch.image_size, arch.in_channels, arch.model_channels, arch.out_channels = image_size, 3, num_channels, 6
        arch.num_res_blocks, arch.num_head_channels = num_res_blocks, num_head_channels
        arch.n_mult = len(mult)
        for i, m in enumerate(mult):
            arch.channel_mult[i] = float(m)
        arch.n_att = len(att)
        for i, d in enumerate(att):
            arch.attention_ds[i] = int(d)
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"UNetEngine precision must be 'bf16' or 'fp32' (got {precision!r})")
        self.precision = precision
        arch.precision = 1 if precision == "fp32" else 0
        self.arch = arch
        self.image_size = image_size
        tensors = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in state_dict.items()}
        if out_cov is not None:   # OpenAIDenoiserV2.out_cov (k_diffusion/external.py:141)
            tensors["out_cov.weight"] = out_cov[0].detach().to(self.device, torch.float32).contiguous()
            tensors["out_cov.bias"] = out_cov[1].detach().to(self.device, torch.float32).contiguous()
        self.has_cov = out_cov is not None
        names = list(tensors.keys())
        n = len(names)
        c_names = (ctypes.c_char_p * n)(*[s.encode() for s in names])
        c_ptrs = (ctypes.c_void_p * n)(*[tensors[s].data_ptr() for s in names])
        c_numel = (ctypes.c_int64 * n)(*[tensors[s].numel() for s in names])
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.kdip_unet_create(ctypes.byref(arch), n, c_names, c_ptrs, c_numel, ctypes.byref(h)))
        self._h = h
        self._ws = None
        self._ws_N = 0
        self._ws_need = 0
        self._fwd_N = 0
        self.forward_token = 0   # bumped by every forward: the VJP is only valid for the latest one
        # CUDA-graph replay of the fixed launch lists (about 140 launches forward, 190 backward): after two eager calls per
        # (pass, batch, options) the launches are captured once on static I/O buffers and replayed; inputs / outputs are copied
        # in and out (25-50 MB, ~10 us).  Same kernels either way; any capture problem falls back to eager launches for good.
        import os
        self._graphs_on = os.environ.get("KDIP_CUDA_GRAPH", "1") != "0"
        self._replays = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.kdip_unet_destroy(h)
            self._h = None

    def workspace_bytes(self, N):
        b = ctypes.c_size_t()
        check(lib.kdip_unet_workspace_bytes(self._h, N, ctypes.byref(b)))
        return b.value

    def _workspace(self, N, at_least=0):
        """(aligned pointer, bytes the UNet plan is keyed on).  The buffer may be larger than the UNet's own need: the fused
        guided evaluation (kdip.ops.FusedGuidedEval) places the UNet workspace first and its operator / guidance scratch behind it,
        so both paths run the SAME launch plan (the library re-plans whenever (N, pointer, size) changes)."""
        if self._ws is None or self._ws_N != N or self._ws.numel() - 256 < at_least:
            need = self.workspace_bytes(N)
            keep = max(need, at_least, (self._ws.numel() - 256) if (self._ws is not None and self._ws_N == N) else 0)
            self._ws = None
            self._ws = torch.empty(keep + 256, dtype=torch.uint8, device=self.device)
            self._ws_N, self._ws_need = N, need
            self._replays.clear()          # captured graphs hold the old pointer
        off = (-self._ws.data_ptr()) % 256
        return ctypes.c_void_p(self._ws.data_ptr() + off), self._ws_need

    def _replay(self, key, statics_fn, launch_fn):
        """Returns the (graph, static buffers) for `key` once two eager calls have happened, else None."""
        if not self._graphs_on:
            return None
        if key not in self._replays and len(self._replays) >= 16:     # bound the static buffers kept for shapes no longer in use
            self._replays.pop(next(iter(self._replays)))
        r = self._replays.setdefault(key, {"calls": 0, "graph": None, "bufs": None, "failed": False})
        if r["failed"]:
            return None
        r["calls"] += 1
        if r["calls"] <= 2:
            return None
        if r["graph"] is None:
            try:
                bufs = statics_fn()
                cur = torch.cuda.current_stream()
                side = torch.cuda.Stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    launch_fn(bufs)                       # settles every lazily built plan for these pointers
                cur.wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = lib.kdip_launch_count()
                with torch.cuda.graph(g):
                    launch_fn(bufs)
                r["graph"], r["bufs"], r["kernels"] = g, bufs, int(lib.kdip_launch_count() - n0)   # kernel nodes per replay
            except Exception as e:                        # noqa: BLE001 - eager launches remain fully functional
                import warnings
                warnings.warn(f"kdip: CUDA-graph capture of {key} failed ({e}); using eager launches")
                r["failed"] = True
                r["graph"] = None
                return None
        return r

    def forward(self, x, t, x_scale=None, out=None, want_cov=False):
        """x [N,3,S,S] fp32 cuda, t [N] (any numeric dtype) -> out [N,6,S,S] fp32 (and cov [N,6,S,S])."""
        N = x.shape[0]
        x = x.contiguous().float()
        t = t.to(self.device, torch.float32).contiguous()
        if x_scale is not None:
            x_scale = x_scale.to(self.device, torch.float32).contiguous()
        if out is None:
            out = torch.empty(N, 6, x.shape[2], x.shape[3], device=self.device, dtype=torch.float32)
        cov = torch.empty_like(out) if want_cov else None
        ws, ws_bytes = self._workspace(N)
        H, W = x.shape[2], x.shape[3]

        def statics():
            mk = lambda *shape: torch.zeros(*shape, device=self.device, dtype=torch.float32)
            return {"x": mk(N, 3, H, W), "t": mk(N), "xs": mk(N) if x_scale is not None else None, "out": mk(N, 6, H, W),
                    "cov": mk(N, 6, H, W) if want_cov else None}

        def launch(b):
            check(lib.kdip_unet_forward(self._h, ptr(b["x"]), ptr(b["xs"]), ptr(b["t"]), N, ptr(b["out"]), ptr(b["cov"]), ws, ws_bytes,
                                        stream_ptr()))

        r = self._replay(("fwd", N, H, W, bool(want_cov), x_scale is not None, self._ws.data_ptr()), statics, launch)
        if r is not None:
            b = r["bufs"]
            b["x"].copy_(x); b["t"].copy_(t)
            if x_scale is not None:
                b["xs"].copy_(x_scale)
            check(lib.kdip_unet_prepare(self._h, N, ws, ws_bytes))   # the replay bypasses the library: keep its (N, workspace) current
            r["graph"].replay()
            lib.kdip_launch_count_add(r["kernels"])
            out.copy_(b["out"])
            if want_cov:
                cov.copy_(b["cov"])
        else:
            check(lib.kdip_unet_forward(self._h, ptr(x), ptr(x_scale), ptr(t), N, ptr(out), ptr(cov), ws, ws_bytes, stream_ptr()))
        self._fwd_N = N
        self.forward_token += 1
        return (out, cov) if want_cov else out

    def feature(self, N):
        """Pre-head feature [N, C0, S, S] fp32 of the last forward (UNetModel.forward(return_feature=True))."""
        assert N == self._fwd_N, "feature must follow forward with the same batch"
        c0 = int(self.arch.channel_mult[0] * self.arch.model_channels)
        feat = torch.empty(N, c0, self.image_size, self.image_size, device=self.device, dtype=torch.float32)
        check(lib.kdip_unet_feature(self._h, N, ptr(feat), stream_ptr()))
        return feat

    def profile(self, x, t, seed, x_scale=None):
        """Instrumented forward + VJP (kdip_unet_profile): per-class device times from CUDA events around every step."""
        from ._lib import UNetProfile
        N = x.shape[0]
        x, seed = x.contiguous().float(), seed.contiguous().float()
        t = t.to(self.device, torch.float32).contiguous()
        out = torch.empty(N, 6, x.shape[2], x.shape[3], device=self.device, dtype=torch.float32)
        g = torch.empty(N, 3, x.shape[2], x.shape[3], device=self.device, dtype=torch.float32)
        ws, ws_bytes = self._workspace(N)
        prof = UNetProfile()
        check(lib.kdip_unet_profile(self._h, ptr(x), ptr(x_scale), ptr(t), ptr(seed), N, ptr(out), ptr(g), ws, ws_bytes,
                                    stream_ptr(), ctypes.byref(prof)))
        self._fwd_N = N
        self.forward_token += 1
        return {k: getattr(prof, k) for k, _ in UNetProfile._fields_}

    def vjp(self, seed, out=None):
        """seed [N,6,S,S] fp32 -> d<seed, unet_out>/d(unet_input) [N,3,S,S] fp32 for the preceding forward."""
        N = seed.shape[0]
        assert N == self._fwd_N, "vjp must follow forward with the same batch"
        seed = seed.contiguous().float()
        if out is None:
            out = torch.empty(N, 3, seed.shape[2], seed.shape[3], device=self.device, dtype=torch.float32)
        ws, ws_bytes = self._workspace(N)
        H, W = seed.shape[2], seed.shape[3]

        def statics():
            return {"seed": torch.zeros(N, 6, H, W, device=self.device, dtype=torch.float32),
                    "grad": torch.zeros(N, 3, H, W, device=self.device, dtype=torch.float32)}

        def launch(b):
            check(lib.kdip_unet_vjp(self._h, ptr(b["seed"]), N, ptr(b["grad"]), ws, ws_bytes, stream_ptr()))

        r = self._replay(("vjp", N, H, W, self._ws.data_ptr()), statics, launch)
        if r is not None:
            r["bufs"]["seed"].copy_(seed)
            check(lib.kdip_unet_prepare(self._h, N, ws, ws_bytes))
            r["graph"].replay()
            lib.kdip_launch_count_add(r["kernels"])
            out.copy_(r["bufs"]["grad"])
        else:
            check(lib.kdip_unet_vjp(self._h, ptr(seed), N, ptr(out), ws, ws_bytes, stream_ptr()))
        return out
