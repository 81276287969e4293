#!/usr/bin/env python
"""Benchmark of the guided-sampling hot path (BASELINE.json: images/sec, 100-step Heun, 256x256).

`--config {0..4}` selects one of BASELINE.json's configs (default 1, the one the metric is quoted on):
  0  FFHQ UNet, inpainting (128x128 box), PiGDM, 20 Euler steps, B=1      (the reference's own CPU-runnable case)
  1  FFHQ UNet, Gaussian deblur, type I / Convert (Eq. 22 + CG below sigma 0.2), 100 Heun steps, B=32
  2  FFHQ UNet, SR x4, type I / Analytic (synthetic recon_mse), 100 Heun steps, B=64
  3  ImageNet UNet (256 ch, 2 res blocks, attention 32/16/8), motion deblur, DPS zeta=1, 100 Heun steps, B=32 per GPU
  4  v2 (DWT-Var) denoiser on the FFHQ UNet, Gaussian deblur, type II, DWT theta + CG below sigma 1, 100 Heun steps (SDE churn,
     as quick_start/eval_guidance_II.sh runs it), B=32 per GPU
One "step" = one complete posterior sampling of the batch through the public API (condition.ConditionOpenAIDenoiser[V2] +
k_diffusion.sampling.sample_heun / sample_euler + evaluation.compute_features).

  value : images/s with the measurements already resident in HBM when the timed region starts
  e2e   : the same through host buffers - pinned host y -> device, sampling, this rank's finished samples -> pinned host, every step
  roofline          : the dominant kernel (tcgen05 implicit-GEMM conv), CUDA-event pairs around every launch of one forward + VJP
  roofline_guidance : every guidance kernel outside the UNet, algorithmic bytes (SURVEY.md 8(d) plane counts) / CUDA-event time
  cpu_baseline      : the reference algorithm on this box's host cores (oracle port), two evaluations extrapolated; plus
                      cpu_anchor_cfg0: configs[0] timed IN FULL (20 evaluations, no extrapolation)

Launch: python bench.py [--config C --gpus N --steps K --warmup W]; for N > 1 under torch.distributed.run (one rank per GPU, NCCL).
`--impl reference` times the reference algorithm's CPU path (the torch-CPU oracle port of the reference's code; the
reference itself is pure Python and cannot travel to the GPU box) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]

METRIC = "images/sec (100-step Heun, 256x256)"
PLANE = 3 * 256 * 256 * 4     # bytes of one fp32 [3,256,256] image ("plane")
FFHQ = {"num_channels": 128, "num_res_blocks": 1, "attention_resolutions": "16"}
IMAGENET = {"num_channels": 256, "num_res_blocks": 2, "attention_resolutions": "32,16,8"}
# per-image FLOPs of one model evaluation (counted on the reference module, 2 x MAC)
GF_FFHQ_FWD_VJP, GF_FFHQ_FWD_V2, GF_IMAGENET_FWD_VJP = 776.26, 388.03, 4491.40

CONFIGS = {
    0: dict(name="configs[0]: FFHQ 256x256 inpainting (128x128 box, sigma_s 0.05), PiGDM isotropic covariance, 20 Euler steps",
            unet=FFHQ, unet_desc="ADM 128ch mult(1,1,2,2,4,4) 1 res-block attn@16 (93.56M params, synthetic weights)",
            operator="inpainting", guidance="pgdm", cov="pgdm", sampler="euler", n_steps=20, batch=1, thres=0.2, gf=GF_FFHQ_FWD_VJP,
            ws_gb=0.516),
    1: dict(name="configs[1]: FFHQ 256x256 gaussian deblur (ks61 std3.0 sigma_s0.05), guidance=I x0_cov=convert mle_sigma_thres=0.2, "
                 "Heun ODE (quick_start --ode) 100 steps sigma 0.01..80 rho 7",
            unet=FFHQ, unet_desc="ADM 128ch mult(1,1,2,2,4,4) 1 res-block attn@16 (93.56M params, synthetic weights)",
            operator="gaussian_blur", guidance="I", cov="convert", sampler="heun", n_steps=100, batch=32, thres=0.2, gf=GF_FFHQ_FWD_VJP,
            ws_gb=0.516),
    2: dict(name="configs[2]: FFHQ 256x256 4x super-resolution (bicubic, sigma_s 0.05), guidance=I x0_cov=analytic (synthetic "
                 "recon_mse 0.5 sigma^2/(1+sigma^2)), Heun ODE 100 steps",
            unet=FFHQ, unet_desc="ADM 128ch mult(1,1,2,2,4,4) 1 res-block attn@16 (93.56M params, synthetic weights)",
            operator="super_resolution", guidance="I", cov="analytic", sampler="heun", n_steps=100, batch=64, thres=0.2,
            gf=GF_FFHQ_FWD_VJP, ws_gb=0.516),
    3: dict(name="configs[3]: ImageNet 256x256 motion deblur (ks61, sigma_s 0.05), DPS guidance zeta=1 (VJP through the UNet), "
                 "Heun ODE 100 steps",
            unet=IMAGENET, unet_desc="ADM 256ch mult(1,1,2,2,4,4) 2 res-blocks attn@32,16,8 (552.8M params, synthetic weights)",
            operator="motion_blur", guidance="dps", cov="dps", sampler="heun", n_steps=100, batch=32, thres=0.2,
            gf=GF_IMAGENET_FWD_VJP, ws_gb=2.2, extra={"zeta": 1.0}),
    4: dict(name="configs[4]: FFHQ 256x256 gaussian deblur, v2 (DWT-Var) denoiser, guidance=II ortho_tf=dwt mle_sigma_thres=1.0 "
                 "(DiffPIR-style closed form above, DWT-domain theta + on-device CG below), Heun SDE (s_churn 80) 100 steps",
            unet=FFHQ, unet_desc="ADM 128ch ... + out_cov Conv2d(128,6,1) head (random init)", operator="gaussian_blur", guidance="II",
            cov=None, sampler="heun", n_steps=100, batch=32, thres=1.0, gf=GF_FFHQ_FWD_V2, ws_gb=0.516, v2=True,
            churn=dict(s_churn=80.0, s_tmin=0.05, s_tmax=50.0, s_noise=1.003)),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="kdip", choices=["kdip", "reference"])
    p.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    p.add_argument("--batch", type=int, default=None, help="images per GPU (default: the config's)")
    p.add_argument("--heun-steps", type=int, default=None, help="sampler steps (default: the config's)")
    p.add_argument("--guidance", default=None, help="override the config's guidance (e.g. pgdm = the north-star target on configs[1])")
    p.add_argument("--cov", default=None)
    p.add_argument("--skip-cpu-baseline", action="store_true")
    p.add_argument("--skip-guidance-roofline", action="store_true")
    args = p.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch is not None:
        cfg["batch"] = args.batch
    if args.heun_steps is not None:
        cfg["n_steps"] = args.heun_steps
    if args.guidance is not None:
        cfg["guidance"] = args.guidance
        cfg["name"] += " [guidance overridden: %s]" % args.guidance
    if args.cov is not None:
        cfg["cov"] = args.cov
    cfg["n_evals"] = cfg["n_steps"] if cfg["sampler"] == "euler" else 2 * cfg["n_steps"] - 1
    args.cfg = cfg
    return args


def workload_config(cfg, n_gpus):
    return {"workload": cfg["name"], "unet": cfg["unet_desc"], "batch_per_gpu": cfg["batch"], "global_batch": cfg["batch"] * n_gpus,
            "model_evals_per_image": cfg["n_evals"],
            "parallelism": "dp%d (independent images, one NCCL all-gather of finished samples)" % n_gpus,
            "l2": "working set >> L2 (activation workspace %.1f GB per rank)" % (cfg["ws_gb"] * cfg["batch"])}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return pk, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [v.strip() for v in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def schedule(n_steps):
    import torch
    ramp = torch.linspace(0, 1, n_steps)
    s = (80 ** (1 / 7.) + ramp * (0.01 ** (1 / 7.) - 80 ** (1 / 7.))) ** 7.
    return s.tolist() + [0.0]


def eval_counts(cfg):
    """How many of the model evaluations run above / below the config's MLE threshold (closed-form vs CG / per-pixel branch)."""
    sig, n = schedule(cfg["n_steps"]), cfg["n_steps"]
    evals = [sig[i] for i in range(n)]
    if cfg["sampler"] == "heun":
        evals += [sig[i + 1] for i in range(n) if sig[i + 1] > 0]
    lo = sum(1 for v in evals if v < cfg["thres"])
    return len(evals) - lo, lo


# ---- the reference's CPU path (oracle port; the only place bench.py executes oracle/) ----------------------------------------
def cpu_model(cfg):
    """-> callable(x [1,3,256,256], sigma [1]) -> hat_x0: the reference algorithm for this config on torch CPU fp32, B = 1."""
    import numpy as np
    import torch
    from oracle import guidance_ref, operators_ref as ops_ref, sampler_ref, unet_ref
    ucfg = unet_ref.imagenet_config() if cfg["unet"] is IMAGENET else unet_ref.ffhq_config()
    sd = unet_ref.init_state_dict(ucfg, seed=0)
    name = cfg["operator"]
    if name == "inpainting":
        op = ops_ref.InpaintingOperator(0.05, ops_ref.box_mask(256, 128))
    elif name == "super_resolution":
        op = ops_ref.SuperResolutionOperator(0.05, 4, in_shape=(1, 3, 256, 256))
    else:
        op = ops_ref.BlurOperator(name, 0.05, 61, 3.0 if name == "gaussian_blur" else 0.5, (1, 3, 256, 256))
    g = torch.Generator().manual_seed(1)
    x0 = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    torch.manual_seed(2)
    y = op.forward(x0, flatten=True)
    s100 = sampler_ref.get_sigmas_karras(100, 0.01, 80)
    recon = {"sigmas": s100[:-1].clone(), "mse_list": 0.5 * s100[:-1] ** 2 / (1 + s100[:-1] ** 2)}
    if cfg.get("v2"):
        gg = torch.Generator().manual_seed(9)
        cov_w, cov_b = torch.randn(6, 128, 1, 1, generator=gg) * 0.05, torch.randn(6, generator=gg) * 0.5 - 1.0
        cm = guidance_ref.ConditionDenoiserV2Ref(sd, ucfg, cov_w, cov_b, op, y, cfg["guidance"], mle_sigma_thres=cfg["thres"],
                                                 ortho_tf_type="dwt")
    else:
        cm = guidance_ref.ConditionDenoiserRef(sd, ucfg, op, y, cfg["guidance"], x0_cov_type=cfg["cov"], recon_mse=recon,
                                               mle_sigma_thres=cfg["thres"], **cfg.get("extra", {}))
    return cm, x0, g, np


def cpu_reference_sample(cfg):
    """Times the reference algorithm on the host cores: one guided model evaluation per sigma branch at B = 1 (the reference
    asserts B == 1), extrapolated to a full trajectory.  Returns (images/s, info)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cm, x0, g, _ = cpu_model(cfg)
    n_hi, n_lo = eval_counts(cfg)
    hi_sigma, lo_sigma = max(1.5, 2 * cfg["thres"]), cfg["thres"] / 2
    t = {}
    cm(x0 * 0.7 + hi_sigma * torch.randn(1, 3, 256, 256, generator=g), torch.tensor([hi_sigma]))   # untimed warm-up (thread pool, oneDNN primitives)
    for name, sigma in (("hi", hi_sigma), ("lo", lo_sigma)):
        xt = x0 * 0.7 + sigma * torch.randn(1, 3, 256, 256, generator=g)
        t0 = time.perf_counter()
        cm(xt, torch.tensor([sigma]))
        t[name] = time.perf_counter() - t0
    t_img = n_hi * t["hi"] + n_lo * t["lo"]
    info = {"cores": cores, "kind": "port",
            "sample": "B=1, after one untimed warm-up eval: one guided eval of this config at sigma=%.2f (%.2f s) and one at sigma=%.2f (%.2f s; below the MLE threshold); "
                      "extrapolated to %d + %d evals per image" % (hi_sigma, t["hi"], lo_sigma, t["lo"], n_hi, n_lo)}
    return 1.0 / t_img, info


def cpu_full_cfg0():
    """configs[0] timed IN FULL on the host cores (20 Euler steps = 20 guided evaluations, B = 1): the non-extrapolated anchor."""
    import torch
    from oracle import sampler_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(CONFIGS[0])
    cm, x0, g, _ = cpu_model(cfg)
    xT = torch.randn(1, 3, 256, 256, generator=g) * 80.0
    t0 = time.perf_counter()
    out = sampler_ref.sample_euler(cm, xT, sampler_ref.get_sigmas_karras(cfg["n_steps"], 0.01, 80))
    dt = time.perf_counter() - t0
    assert bool(torch.isfinite(out).all())
    return {"value": 1.0 / dt, "unit": "images/s", "cores": cores, "kind": "port", "seconds_per_image": dt,
            "sample": "configs[0] in full: 20 Euler steps = 20 guided evaluations (FFHQ UNet fwd + VJP, inpainting, PiGDM), B=1, no extrapolation"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.cfg
    vals, info = [], None
    for i in range(args.warmup + args.steps):
        if args.config == 0:
            a = cpu_full_cfg0()
            v, info = a["value"], {k: a[k] for k in ("cores", "kind", "sample")}
        else:
            v, info = cpu_reference_sample(cfg)
        if i >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, args.gpus),
            "cpu_baseline": dict(value=v, unit="images/s", **info),
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print on stdout (e.g. NCCL's version banner) goes to stderr: stdout carries the ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def guidance_roofline(cfg, operator, dev, B, pk):
    """Per-kernel HBM fraction of the guidance kernels outside the UNet: algorithmic bytes (SURVEY.md 8(d) plane counts x B)
    over the CUDA-event time of the call, measured here on the bench batch."""
    import torch
    from kdip import ops
    peak = pk["hbm_gbs"]
    g = torch.Generator(device="cpu").manual_seed(5)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dev)
    x, x0, v, grad = rnd(B, 3, 256, 256), rnd(B, 3, 256, 256) * 0.5, rnd(B, 3, 256, 256), rnd(B, 3, 256, 256)
    out6 = rnd(B, 6, 256, 256)
    from guided_diffusion.script_util import create_gaussian_diffusion
    sc = ops.pmv_scalars(create_gaussian_diffusion(learn_sigma=True), [40] * B, [0.9] * B, dev)
    coef = torch.full((B,), 0.3, device=dev)
    theta = torch.full((B,), 0.2, device=dev)
    rows = []

    def timeit(name, planes, fn, reps=20, note=None):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        gbs = planes * B * PLANE / (us * 1e-6) / 1e9
        row = {"kernel": name, "us": us, "algorithmic_planes_per_image": planes, "GB/s": gbs, "frac": gbs / peak}
        spectral = ("spectral" in name or "closed form" in name or "dps_grad" in name or "CG" in name) and operator.name != "inpainting"
        # the FFT-based entries are bound by the register FFTs (instruction issue + L2 latency), not by HBM: their
        # HBM fraction is reported for completeness, the instruction floor of one blur application at B = 32 is ~15 us
        row["bound"] = "fp32 ALU / latency (register FFT), not HBM" if spectral else "hbm"
        if note:
            row["note"] = note
        rows.append(row)
        return us

    timeit("pmv_epilogue (Convert Eq. 22)", 5, lambda: ops.pmv_epilogue(out6, x, sc, ops.VAR_CONVERT))
    timeit("pmv_vjp_seed", 3, lambda: ops.pmv_vjp_seed(x0, v, sc), note="writes the 6-channel seed (3 zero channels) + direct term: 7 planes touched")
    timeit("guidance_combine", 3, lambda: ops.guidance_combine(x0, grad, v, coef, coef), note="4 planes touched (direct term)")
    timeit("euler_step", 3, lambda: ops.euler_step(x, x0, 1.5, -0.2))
    timeit("heun_step", 5, lambda: ops.heun_step(x, v, x0, grad, 1.2, -0.2))
    h = operator.handle
    y = h.forward(x0, None)
    y = y + 0.05 * torch.randn(y.shape, generator=g).to(dev)        # a noisy measurement: y - A x0 != 0, the solvers have work to do
    if operator.name in ("gaussian_blur", "motion_blur"):
        timeit("A x (blur, spectral)", 2, lambda: h.forward(x0, None))
        timeit("A^T y (blur, spectral)", 2, lambda: h.transpose(y))
    elif operator.name == "super_resolution":
        timeit("A x (Resizer)", 1 + 1 / 16, lambda: h.forward(x0, None))
        timeit("A^T y (SR, spectral)", 1 + 1 / 16, lambda: h.transpose(y))
    else:
        timeit("A x (mask)", 2, lambda: h.forward(x0, None))
    timeit("mat closed form (%s)" % operator.name, 3 if operator.name != "super_resolution" else 2 + 1 / 16, lambda: h.mat_closed(y, x0, theta))
    timeit("dps_grad (A^T r, ||r||)", 3, lambda: h.dps_grad(y, x0))
    tmap = torch.rand(B, 3, 256, 256, generator=g).to(dev) * 0.05 + 1e-4
    for ot in ((None, "dwt") if cfg.get("v2") else (None,)):
        h.mat_cg(y, x0, tmap, ot=ot)
        its = max(h.last_cg_iters)
        us = timeit("CG mat solve (%s, theta map%s)" % (operator.name, ", dwt" if ot else ""), 8 * its, lambda: h.mat_cg(y, x0, tmap, ot=ot),
                    reps=3, note="%d iterations (all images of the batch iterate until the last one converges)" % its)
        rows[-1]["iterations"] = its
        rows[-1]["us_per_iteration"] = us / its
    if cfg.get("v2"):
        timeit("dwt forward * theta", 2, lambda: ops.ortho("dwt", x0, mul=tmap), note="3 planes touched (theta)")
        timeit("dwt inverse", 2, lambda: ops.ortho("dwt", x0, inverse=True))
    return {"bound": "hbm", "peak": peak, "unit": "GB/s", "batch": B, "kernels": rows,
            "how": "CUDA events around 20 back-to-back calls of each op through the product API at the bench batch; "
                   "achieved = algorithmic planes x B x 786432 B / time (the 25-100 MB working sets partly live in the 126 MB L2)"}


def main():
    args = parse()
    cfg = args.cfg
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    from condition.condition import ConditionOpenAIDenoiser, ConditionOpenAIDenoiserV2
    from condition.diffpir_utils.utils_model import create_argparser
    from condition.measurements import get_operator
    from guided_diffusion.script_util import args_to_dict, create_model_and_diffusion, model_and_diffusion_defaults
    import k_diffusion as K
    from kdip._lib import lib
    from kdip.dist import Accelerator
    from kdip.synth import synthetic_state_dict
    import numpy as np

    acc = Accelerator()
    n_gpus = acc.num_processes
    assert n_gpus == args.gpus or n_gpus == 1, "launch with torch.distributed.run --nproc-per-node N for --gpus N"
    dev = acc.device
    B = cfg["batch"]

    margs = create_argparser(cfg["unet"]).parse_args([])
    model, diffusion = create_model_and_diffusion(**args_to_dict(margs, model_and_diffusion_defaults().keys()))
    model.load_state_dict(synthetic_state_dict(model, seed=0))
    model = model.eval().to(dev)
    name = cfg["operator"]
    if name == "inpainting":
        np.random.seed(0)
        operator = get_operator(name=name, sigma_s=0.05, device=dev, mask_opt=dict(mask_type="box", mask_len_range=(128, 129), image_size=256))
    elif name == "super_resolution":
        operator = get_operator(name=name, in_shape=(1, 3, 256, 256), scale_factor=4, sigma_s=0.05, device=dev)
    else:
        operator = get_operator(name=name, in_shape=(1, 3, 256, 256), kernel_size=61, intensity=3.0 if name == "gaussian_blur" else 0.5,
                                sigma_s=0.05, device=dev)
    sigmas = K.sampling.get_sigmas_karras(cfg["n_steps"], 0.01, 80.0, rho=7.0, device=dev)
    recon = {"sigmas": K.sampling.get_sigmas_karras(100, 0.01, 80.0, rho=7.0)[:-1].clone()}
    recon["mse_list"] = 0.5 * recon["sigmas"] ** 2 / (1 + recon["sigmas"] ** 2)
    denoiser_v2 = None
    if cfg.get("v2"):
        denoiser_v2 = K.external.OpenAIDenoiserV2(model, diffusion, device=dev, ortho_tf_type="dwt").to(dev)
        gg = torch.Generator().manual_seed(9)
        with torch.no_grad():
            denoiser_v2.out_cov.weight.copy_(torch.randn(6, 128, 1, 1, generator=gg) * 0.05)
            denoiser_v2.out_cov.bias.copy_(torch.randn(6, generator=gg) * 0.5 - 1.0)

    # synthetic ground truth / measurements, seeded by GLOBAL image index so results do not depend on the GPU count
    lo, hi = acc.process_index * B, (acc.process_index + 1) * B
    x0 = torch.stack([torch.rand(3, 256, 256, generator=torch.Generator().manual_seed(1000 + i)) * 2 - 1 for i in range(lo, hi)])
    torch.manual_seed(2 + acc.process_index)
    y_dev = operator.forward(x0.to(dev))
    y_host = y_dev.cpu().pin_memory()
    out_host = torch.empty(B, 3, 256, 256).pin_memory()          # this rank's shard of the finished samples
    sampler = K.sampling.sample_euler if cfg["sampler"] == "euler" else K.sampling.sample_heun
    skw = dict(cfg.get("churn", {}))

    def sample(y):
        meas = (y, y.reshape(B, -1))
        if denoiser_v2 is not None:
            cm = ConditionOpenAIDenoiserV2(denoiser=denoiser_v2, operator=operator, measurement=meas, guidance=cfg["guidance"],
                                           mle_sigma_thres=cfg["thres"], ortho_tf_type="dwt", device=dev).eval()
        else:
            cm = ConditionOpenAIDenoiser(inner_model=model, diffusion=diffusion, x0_cov_type=cfg["cov"], recon_mse=dict(recon),
                                         operator=operator, measurement=meas, guidance=cfg["guidance"], mle_sigma_thres=cfg["thres"],
                                         device=dev, **cfg.get("extra", {})).eval()
        local = {}

        def sample_fn(n):
            x = torch.randn([n, 3, 256, 256], device=dev) * 80.0
            local["x"] = sampler(cm, x, sigmas, disable=True, **skw)
            return local["x"]
        return K.evaluation.compute_features(acc, sample_fn, lambda x: x, B * n_gpus, B), local["x"]

    def timed(fn, k):
        acc.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        acc.barrier()
        torch.cuda.synchronize()
        mine = e0.elapsed_time(e1)
        return acc.max_over_ranks(mine), mine

    def step_resident():
        return sample(y_dev)

    def step_e2e():
        y = y_host.to(dev, non_blocking=True)
        _, mine = sample(y)                          # the all-gather of the finished samples still runs (it is part of the path)
        out_host.copy_(mine, non_blocking=True)      # ...but every rank hands only its OWN shard to the host
        torch.cuda.current_stream().synchronize()    # the result is on the host when the step ends

    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    l0 = lib.kdip_launch_count()
    ms, ms_mine = timed(step_resident, args.steps)
    launches = lib.kdip_launch_count() - l0
    clk = clocks.stop()
    ms_e2e, _ = timed(step_e2e, args.steps)
    imgs = B * n_gpus * args.steps
    value = imgs / (ms / 1000.0)
    e2e_value = imgs / (ms_e2e / 1000.0)
    per_rank_ms = [ms_mine / args.steps]
    if n_gpus > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_mine / args.steps], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(n_gpus)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(v.item()) for v in allt]

    # live roofline of the dominant kernel (tcgen05 implicit-GEMM conv): CUDA-event pairs around every launch of one
    # forward + VJP at the bench batch, right after the timed region
    pk, pk_src = peaks()
    eng = model.engine()
    xs = torch.randn(B, 3, 256, 256, device=dev) * 1.2
    tt = torch.full((B,), 338.0, device=dev)
    sd6 = torch.randn(B, 6, 256, 256, device=dev)
    eng.profile(xs, tt, sd6)
    pr = eng.profile(xs, tt, sd6)
    conv_tf = pr["conv_flops"] / (pr["conv_ms"] * 1e-3) / 1e12
    traffic, traffic_note = None, "not measured in this run (ncu only): see profiles/conv_traffic.json for the capture it came from"
    try:
        with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("round") == 2 and B == tj.get("batch") and args.config in tj.get("configs", [1]):
            traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("how")
    except Exception:
        pass
    roofline = {"kernel": "kdip::conv_gemm_kernel (tcgen05 implicit-GEMM conv, fwd + dgrad)", "bound": "tensor",
                "achieved": conv_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": conv_tf / pk["bf16_tflops_sustained"], "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": pk_src + ", sustained (kernel timed inside a long step)",
                "frac_of_burst": conv_tf / pk["bf16_tflops"], "launches_per_eval": pr["conv_launches"],
                "avg_launch_ms": pr["conv_ms"] / max(1, pr["conv_launches"]),
                "algorithmic_flops_per_launch": pr["conv_flops"] / max(1, pr["conv_launches"]),
                "conv_share_of_eval": pr["conv_ms"] / pr["total_ms"], "eval_ms": pr["total_ms"],
                "how": "one CUDA event between consecutive launches of one UNet forward+VJP (B=%d) on the launching stream (eager "
                       "launches, right after the timed region)" % B,
                "note": "since round 2 the conv launches of the 256- and 128-pixel levels also apply the GroupNorm affine + SiLU of their "
                        "input on the operand path (16 launches): that work is charged to the conv kernel here, the separate pass it "
                        "replaced was not"}
    unet_tf = value / n_gpus * cfg["n_evals"] * cfg["gf"] / 1e3
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(cfg, n_gpus), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": y_host.numel() * 4 * n_gpus,
                    "d2h_bytes_per_step": out_host.numel() * 4 * n_gpus, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "roofline": roofline, "ms_per_step_per_rank": per_rank_ms,
            "ms_per_guided_eval": ms / args.steps / cfg["n_evals"],
            "unet_tflops_per_gpu": unet_tf, "unet_frac_of_bf16_peak": unet_tf / pk["bf16_tflops"],
            "unet_frac_of_bf16_sustained": unet_tf / pk["bf16_tflops_sustained"]}
    if acc.is_main_process:
        if n_gpus == 1 and not args.skip_guidance_roofline:
            line["roofline_guidance"] = guidance_roofline(cfg, operator, dev, B, pk)
        if n_gpus == 1 and not args.skip_cpu_baseline:
            v, info = cpu_reference_sample(cfg)
            line["cpu_baseline"] = dict(value=v, unit="images/s", **info)
            if args.config in (0, 1):
                line["cpu_anchor_cfg0"] = cpu_full_cfg0()
        emit(line)
    if n_gpus > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def _trap_report():
    """After a failed run: the conv pipeline's host-mapped wait-timeout record (include/kdip.h: kdip_conv_trap_read), if any."""
    try:
        import ctypes
        from kdip._lib import lib
        buf = (ctypes.c_uint * 64)()
        n = lib.kdip_conv_trap_read(buf, 64)
        if n:
            recs = [tuple(buf[4 + 4 * i: 8 + 4 * i]) for i in range(min(n, 15))]
            print("[bench] conv_gemm wait timeouts: %d; (line, block, thread|parity<<16, barrier) = %s" % (n, recs), file=sys.stderr)
        else:
            print("[bench] no conv_gemm wait-timeout record", file=sys.stderr)
    except Exception as e:                                                   # noqa: BLE001
        print("[bench] trap record unavailable: %r" % (e,), file=sys.stderr)


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        _trap_report()
        raise
