#!/usr/bin/env python
"""Benchmark of the guided-sampling hot path (BASELINE.json: images/sec, 100-step Heun, 256x256).

Workload = BASELINE.json configs[1]: FFHQ 256x256 ADM UNet (synthetic weights), Gaussian deblur (61x61, std 3, sigma_s 0.05),
type-I guidance with Convert covariance (per-pixel Eq. 22 + on-device CG below sigma 0.2), 100 Heun steps (199 guided model
evaluations = UNet forward + input-VJP each), batch 32 per GPU.  One "step" = one complete posterior sampling of the batch
through the public API (condition.ConditionOpenAIDenoiser + k_diffusion.sampling.sample_heun + evaluation.compute_features).

  value : images/s with the measurements already resident in HBM when the timed region starts
  e2e   : the same through host buffers - pinned host y -> device, sampling, finished samples -> pinned host, every step

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torch.distributed.run (one rank per GPU, NCCL).
`--impl reference` times the reference algorithm's CPU path (the torch-CPU oracle port of the reference's code; the
reference itself is pure Python and cannot travel to the GPU box) on this box's host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "k-diffusion-inverse-problems_b200")]

METRIC = "images/sec (100-step Heun, 256x256)"
UNET_FWD_GF = 387.93          # SURVEY.md §6 [probe]: FFHQ UNet forward FLOPs per image (2 x MAC)
UNET_FWD_VJP_GF = 776.26      # forward + input-VJP


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="kdip", choices=["kdip", "reference"])
    p.add_argument("--batch", type=int, default=32, help="images per GPU (configs[1]: 32)")
    p.add_argument("--heun-steps", type=int, default=100)
    p.add_argument("--guidance", default="I")
    p.add_argument("--cov", default="convert")
    p.add_argument("--skip-cpu-baseline", action="store_true")
    return p.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "configs[1]: FFHQ 256x256 gaussian deblur (ks61 std3.0 sigma_s0.05), guidance=%s x0_cov=%s "
                        "mle_sigma_thres=0.2, Heun ODE (quick_start --ode) %d steps sigma 0.01..80 rho 7" % (args.guidance, args.cov, args.heun_steps),
            "unet": "ADM 128ch mult(1,1,2,2,4,4) 1 res-block attn@16 (93.56M params, synthetic weights)",
            "batch_per_gpu": args.batch, "global_batch": args.batch * n_gpus, "model_evals_per_image": 2 * args.heun_steps - 1,
            "parallelism": "dp%d (independent images, one NCCL all-gather of finished samples)" % n_gpus,
            "l2": "working set >> L2 (activation workspace %.1f GB per rank)" % (0.516 * args.batch)}


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


def eval_counts(heun_steps, thres=0.2):
    """How many of the 2N-1 model evaluations run above / below the MLE threshold (closed-form vs CG branch)."""
    import torch
    ramp = torch.linspace(0, 1, heun_steps)
    s = (80 ** (1 / 7.) + ramp * (0.01 ** (1 / 7.) - 80 ** (1 / 7.))) ** 7.
    sig = s.tolist() + [0.0]
    evals = [sig[i] for i in range(heun_steps)] + [sig[i + 1] for i in range(heun_steps) if sig[i + 1] > 0]
    lo = sum(1 for v in evals if v < thres)
    return len(evals) - lo, lo


def cpu_reference_sample(args, n_repeat=1):
    """Times the reference algorithm on the host cores: oracle port (torch CPU fp32) of one guided model evaluation per
    branch at B = 1 (the reference asserts B == 1), extrapolated to a full trajectory.  Returns (images/s, info)."""
    import torch
    from oracle import guidance_ref, operators_ref as ops_ref, unet_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = unet_ref.ffhq_config()
    sd = unet_ref.init_state_dict(cfg, seed=0)
    op = ops_ref.BlurOperator("gaussian_blur", 0.05, 61, 3.0, (1, 3, 256, 256))
    g = torch.Generator().manual_seed(1)
    x0 = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    y = op.forward(x0, noise=torch.randn(1, 3, 256, 256, generator=g))
    cm = guidance_ref.ConditionDenoiserRef(sd, cfg, op, y, args.guidance, x0_cov_type=args.cov, mle_sigma_thres=0.2)
    n_hi, n_lo = eval_counts(args.heun_steps)
    t = {}
    for name, sigma in (("hi", 1.5), ("lo", 0.1)):
        xt = x0 * 0.7 + sigma * torch.randn(1, 3, 256, 256, generator=g)
        best = float("inf")
        for _ in range(n_repeat):
            t0 = time.perf_counter()
            cm(xt, torch.tensor([sigma]))
            best = min(best, time.perf_counter() - t0)
        t[name] = best
    t_img = n_hi * t["hi"] + n_lo * t["lo"]
    info = {"cores": cores, "kind": "port",
            "sample": "B=1, one guided eval (UNet fwd + autograd VJP + mat solver) at sigma=1.5 (closed form, %.2f s) and one at "
                      "sigma=0.1 (per-pixel Convert covariance, scipy CG, %.2f s); extrapolated to %d + %d evals per image"
                      % (t["hi"], t["lo"], n_hi, n_lo)}
    return 1.0 / t_img, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        v, info = cpu_reference_sample(args)
        if i >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    ms = 1000.0 / v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
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


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    from condition.condition import ConditionOpenAIDenoiser
    from condition.diffpir_utils.utils_model import create_argparser
    from condition.measurements import get_operator
    from guided_diffusion.script_util import args_to_dict, create_model_and_diffusion, model_and_diffusion_defaults
    import k_diffusion as K
    from kdip._lib import lib
    from kdip.dist import Accelerator
    from kdip.synth import synthetic_state_dict

    acc = Accelerator()
    n_gpus = acc.num_processes
    assert n_gpus == args.gpus or n_gpus == 1, "launch with torch.distributed.run --nproc-per-node N for --gpus N"
    dev = acc.device
    B = args.batch

    margs = create_argparser({"num_channels": 128, "num_res_blocks": 1, "attention_resolutions": "16"}).parse_args([])
    model, diffusion = create_model_and_diffusion(**args_to_dict(margs, model_and_diffusion_defaults().keys()))
    model.load_state_dict(synthetic_state_dict(model, seed=0))
    model = model.eval().to(dev)
    operator = get_operator(name="gaussian_blur", in_shape=(1, 3, 256, 256), kernel_size=61, intensity=3.0, sigma_s=0.05, device=dev)
    sigmas = K.sampling.get_sigmas_karras(args.heun_steps, 0.01, 80.0, rho=7.0, device=dev)

    # synthetic ground truth / measurements, seeded by GLOBAL image index so results do not depend on the GPU count
    lo, hi = acc.process_index * B, (acc.process_index + 1) * B
    x0 = torch.stack([torch.rand(3, 256, 256, generator=torch.Generator().manual_seed(1000 + i)) * 2 - 1 for i in range(lo, hi)])
    torch.manual_seed(2 + acc.process_index)
    y_dev = operator.forward(x0.to(dev), flatten=True)[0]
    y_host = y_dev.cpu().pin_memory()
    out_host = torch.empty(B * n_gpus, 3, 256, 256).pin_memory()

    def sample(y):
        cm = ConditionOpenAIDenoiser(inner_model=model, diffusion=diffusion, x0_cov_type=args.cov, recon_mse=None,
                                     operator=operator, measurement=(y, y.reshape(B, -1)), guidance=args.guidance,
                                     mle_sigma_thres=0.2, device=dev).eval()

        def sample_fn(n):
            x = torch.randn([n, 3, 256, 256], device=dev) * 80.0
            return K.sampling.sample_heun(cm, x, sigmas, disable=True)
        return K.evaluation.compute_features(acc, sample_fn, lambda x: x, B * n_gpus, B)

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
        return acc.max_over_ranks(e0.elapsed_time(e1))

    def step_resident():
        return sample(y_dev)

    def step_e2e():
        y = y_host.to(dev, non_blocking=True)
        out = sample(y)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the result is on the host when the step ends

    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    l0 = lib.kdip_launch_count()
    ms = timed(step_resident, args.steps)
    launches = lib.kdip_launch_count() - l0
    clk = clocks.stop()
    ms_e2e = timed(step_e2e, args.steps)
    imgs = B * n_gpus * args.steps
    value = imgs / (ms / 1000.0)
    e2e_value = imgs / (ms_e2e / 1000.0)

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
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": "kdip::conv_gemm_kernel (tcgen05 implicit-GEMM conv, fwd + dgrad)", "bound": "tensor",
                "achieved": conv_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": conv_tf / pk["bf16_tflops_sustained"], "traffic": traffic, "peak_source": pk_src + ", sustained (kernel timed inside a long step)",
                "frac_of_burst": conv_tf / pk["bf16_tflops"], "launches_per_eval": pr["conv_launches"],
                "avg_launch_ms": pr["conv_ms"] / max(1, pr["conv_launches"]),
                "algorithmic_flops_per_launch": pr["conv_flops"] / max(1, pr["conv_launches"]),
                "conv_share_of_eval": pr["conv_ms"] / pr["total_ms"], "eval_ms": pr["total_ms"],
                "how": "cudaEvent pairs around each launch of one UNet forward+VJP (B=%d) on the launching stream" % B}
    unet_tf = value / n_gpus * (2 * args.heun_steps - 1) * UNET_FWD_VJP_GF / 1e3
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args, n_gpus), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": y_host.numel() * 4 * n_gpus,
                    "d2h_bytes_per_step": out_host.numel() * 4 * n_gpus, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "roofline": roofline,
            "unet_tflops_per_gpu": unet_tf, "unet_frac_of_bf16_peak": unet_tf / pk["bf16_tflops"]}
    if acc.is_main_process:
        if n_gpus == 1 and not args.skip_cpu_baseline:
            v, info = cpu_reference_sample(args)
            line["cpu_baseline"] = dict(value=v, unit="images/s", **info)
        emit(line)
    if n_gpus > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
