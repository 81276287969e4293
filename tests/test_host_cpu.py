"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol of include/kdip.h, the Python mirror
declares the reference's interface (state_dict schema, schedule constants, sigma<->t, Resizer tables, registries, mask
generator), the product path refuses to run without CUDA, and the N > 1 path (sharding + one all-gather) works under gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import inputs as I

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from kdip import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 45
    for name in protos:
        assert hasattr(_lib.lib, name), f"libkdip.so does not export {name}"
    assert _lib.lib.kdip_version() >= 1
    assert _lib.lib.kdip_launch_count() == 0          # no kernel was launched: nothing here computes without a GPU
    # every prototype cites the reference interface it replaces somewhere in the header
    hdr = open(_lib.HEADER).read()
    for ref in ("sampling.py", "condition.py", "measurements.py", "gaussian_diffusion.py", "unet.py", "utils_sisr.py", "resizer.py", "external.py"):
        assert ref in hdr


def test_unet_module_schema_matches_reference_state_dict():
    from oracle import unet_ref
    from condition.diffpir_utils.utils_model import create_argparser
    from guided_diffusion.script_util import args_to_dict, create_model_and_diffusion, model_and_diffusion_defaults
    for cfg, over in ((unet_ref.ffhq_config(), {"num_channels": 128, "num_res_blocks": 1, "attention_resolutions": "16"}),
                      (unet_ref.imagenet_config(), {"num_channels": 256, "num_res_blocks": 2, "attention_resolutions": "32,16,8"})):
        args = create_argparser(over).parse_args([])
        model, diffusion = create_model_and_diffusion(**args_to_dict(args, model_and_diffusion_defaults().keys()))
        ref = unet_ref.param_shapes(cfg)
        sd = model.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        assert all(tuple(sd[k].shape) == tuple(ref[k]) for k in ref)
    assert sum(v.numel() for v in sd.values()) == 552814086           # SURVEY.md §6 (ImageNet UNet)
    # load_state_dict works with the reference's keys; the CUDA engine refuses a CPU model
    model.load_state_dict(unet_ref.init_state_dict(unet_ref.imagenet_config(), seed=1), strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 3, 256, 256), torch.zeros(1))


def test_unsupported_unet_configuration_is_loud():
    from guided_diffusion.unet import UNetModel
    with pytest.raises(NotImplementedError):
        UNetModel(image_size=64, in_channels=3, model_channels=64, out_channels=6, num_res_blocks=1, attention_resolutions=(4,),
                  num_head_channels=32, use_scale_shift_norm=True, resblock_updown=True)


def test_schedule_constants_and_sigma_to_t(golden_small):
    from guided_diffusion.script_util import create_gaussian_diffusion
    from k_diffusion.external import OpenAIDenoiser
    import k_diffusion as K
    d = create_gaussian_diffusion(learn_sigma=True)
    for name in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                 "posterior_log_variance_clipped", "posterior_mean_coef1", "betas"):
        np.testing.assert_array_equal(getattr(d, name), golden_small["sched." + name])
    den = OpenAIDenoiser(None, d)
    np.testing.assert_allclose(den.sigma_to_t(I.SIGMA_PROBE).numpy(), golden_small["sched.sigma_to_t"], rtol=1e-6, atol=1e-4)
    host = np.array([den.sigma_to_t_host(float(s)) for s in I.SIGMA_PROBE])
    np.testing.assert_allclose(host, golden_small["sched.sigma_to_t"], rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(den.get_scalings(I.SIGMA_PROBE)[1].numpy(), golden_small["sched.c_in"], rtol=1e-6)
    np.testing.assert_allclose(K.sampling.get_sigmas_karras(100, 0.01, 80, rho=7.).numpy(), golden_small["sched.karras100"], rtol=1e-6)
    # timestep respacing bookkeeping (respace.py:63-85)
    from guided_diffusion.respace import space_timesteps
    assert space_timesteps(1000, [1000]) == set(range(1000))
    assert len(space_timesteps(1000, "ddim50")) == 50 and len(space_timesteps(1000, "10,15,20")) == 45


def test_resizer_tables_match_oracle():
    from oracle import operators_ref as ops
    from condition.dps_utils.resizer import Resizer
    for S in (64, 256):
        w, idx = Resizer((1, 3, S, S), 1 / 4).tables
        wr, ir = ops.resizer_contributions(S, S // 4, 0.25)
        np.testing.assert_allclose(w, wr.astype(np.float32), rtol=0, atol=0)
        np.testing.assert_array_equal(idx, ir)
        assert w.shape == (S // 4, 16)


def test_registries_and_mask_generator(golden_small):
    from condition import measurements as M
    from condition.utils import OrthoTransform, register_ot
    with pytest.raises(NameError):
        M.get_operator(name="does_not_exist")
    with pytest.raises(NameError):
        M.register_operator("inpainting")(type("X", (), {}))
    assert set(M.__OPERATOR__) == {"super_resolution", "motion_blur", "gaussian_blur", "inpainting"}
    with pytest.raises(ValueError):
        OrthoTransform("fourier")
    x = torch.randn(1, 3, 8, 8)
    assert OrthoTransform(None)(x) is x and OrthoTransform(None).inv(x) is x
    for size in (64, 256):
        np.random.seed(7)
        m = M.MaskGenerator("random", mask_prob_range=(0.5, 0.5), image_size=size)(torch.zeros(1, 3, size, size))
        np.testing.assert_array_equal(np.packbits(m.numpy().astype(np.uint8)[0, 0]), golden_small[f"op{size}.random_mask"])
    np.random.seed(0)
    box = M.MaskGenerator("box", mask_len_range=(128, 129), image_size=256)(torch.zeros(1, 3, 256, 256))
    assert int((box == 0).sum()) == 3 * 128 * 128 and box[0, 0, 64, 64] == 0 and box[0, 0, 63, 63] == 1
    from condition.condition import __MAT_SOLVER__
    assert set(__MAT_SOLVER__) == {"inpainting", "gaussian_blur", "motion_blur", "super_resolution"}


def test_no_cpu_fallback():
    import k_diffusion as K
    with pytest.raises(RuntimeError, match="CUDA"):
        K.sampling.sample_heun(lambda x, s: x, torch.zeros(1, 3, 8, 8), K.sampling.get_sigmas_karras(3, 0.01, 80))
    from kdip import ops
    with pytest.raises((AssertionError, RuntimeError)):
        ops.euler_step(torch.zeros(4), torch.zeros(4), 1.0, -0.5)
    # the front end either side of the path is CUDA-only as well
    from kdip.data import ImageBatchLoader
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ImageBatchLoader(K.utils.FolderOfImages(ROOT), batch_size=1)
    with pytest.raises((AssertionError, RuntimeError)):
        ops.psnr(torch.zeros(1, 3, 8, 8), torch.ones(1, 3, 8, 8))
    with pytest.raises((AssertionError, RuntimeError)):
        ops.lincomb3(torch.zeros(8), 1.0)
    # nothing under the package (the sub-packages and the top-level analytic_variance.py / train_openai.py) imports the oracle
    checked = 0
    for dirpath, _, files in os.walk(os.path.join(ROOT, "k-diffusion-inverse-problems_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                checked += 1
                assert "import oracle" not in text and "from oracle" not in text, f"{f}: the product must not import the oracle"
    assert checked > 20
    for f in os.listdir(os.path.join(ROOT, "tools")):        # the measurement scripts use the product package alone
        if f.endswith(".py"):
            text = open(os.path.join(ROOT, "tools", f)).read()
            assert "import oracle" not in text and "from oracle" not in text, f"tools/{f} must not import the oracle"


_WORKER = r'''
import os, sys, torch
sys.path[:0] = [r"%(root)s", r"%(pkg)s"]
from kdip.dist import Accelerator
import k_diffusion as K
acc = Accelerator(device="cpu")
assert acc.num_processes == 2
n, bs = 7, 2                       # ragged: ceil(7/2) = 4 per rank, batches of 2
def sample_fn(b):                  # each rank produces samples tagged with its rank and a running counter
    sample_fn.k += 1
    return torch.full((b, 3, 4, 4), float(acc.process_index * 100 + sample_fn.k))
sample_fn.k = 0
out = K.evaluation.compute_features(acc, sample_fn, lambda x: x, n, bs)
assert out.shape == (7, 3, 4, 4), out.shape
tags = out[:, 0, 0, 0].tolist()
assert tags == [1.0, 1.0, 101.0, 101.0, 2.0, 2.0, 102.0], tags     # batch-major, rank-ordered gather, trimmed to n
assert acc.shard(7) == ((0, 4) if acc.process_index == 0 else (4, 7))
assert acc.max_over_ranks(float(acc.process_index)) == 1.0
acc.barrier()
print("rank", acc.process_index, "ok")
'''


def test_two_rank_gloo_shard_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "pkg": os.path.join(ROOT, "k-diffusion-inverse-problems_b200")})
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()


@pytest.mark.parametrize("case", I.STSL_CASES[:2], ids=lambda c: f"{c[0]}-{c[1]}")
def test_stsl_host_composition_against_reference_golden(case, golden_stsl, monkeypatch):
    """ConditionDenoiser._stsl_guidance_impl composes STSL (condition.py:185-208) from explicit input-VJPs instead of autograd on
    the loss.  Here the host code itself runs on CPU with its device pieces (UNet forward / VJP, dps_grad, lincomb, combine)
    replaced by oracle-backed stand-ins, and must reproduce the REFERENCE's output: this pins the decomposition, the
    coefficients and the order of the probe draws without a GPU (the GPU parity test then checks the kernels)."""
    from oracle import guidance_ref, operators_ref, unet_ref
    from condition import condition as C
    opname, sigma, zeta, eta, n, seed = case
    cfg = unet_ref.tiny_config()
    sd = unet_ref.init_state_dict(cfg, seed=0)
    ref_op = {"gaussian_blur": lambda: operators_ref.BlurOperator("gaussian_blur", 0.05, in_shape=(1, 3, 64, 64)),
              "inpainting": lambda: operators_ref.InpaintingOperator(0.05, operators_ref.box_mask(64, 32))}[opname]()
    x0 = I.image(64, batch=1, seed=1)
    torch.manual_seed(2)
    meas = ref_op.forward(x0.clone(), flatten=True)
    oracle = guidance_ref.ConditionDenoiserRef(sd, cfg, ref_op, meas, "stsl", "pgdm")

    class Handle:
        def dps_grad(self, y, x0m):
            x0g = x0m.detach().requires_grad_()
            nrm = torch.linalg.norm(y - ref_op.forward(x0g, noiseless=True))
            return -torch.autograd.grad(nrm, x0g)[0] * nrm.detach(), nrm.detach().reshape(1)

    class Op:
        name, handle = opname, Handle()

    class Host(C.ConditionDenoiser):
        def uncond_pred(self, x, sigma):
            xg = x.detach().requires_grad_()
            x0m = oracle.uncond_pred(xg, sigma)[0]
            self._ctx = dict(c_in_dev=torch.ones(1), xg=xg, x0m=x0m)
            return x0m.detach(), None, None

        def _score(self, x0_mean, v):
            assert torch.equal(x0_mean, self._ctx["x0m"].detach()), "VJP requested for a stale forward"
            return torch.autograd.grad((self._ctx["x0m"] * v).sum(), self._ctx["xg"])[0], torch.zeros_like(v)

    bc = lambda a: torch.as_tensor(a, dtype=torch.float32).reshape(-1, 1, 1, 1)
    monkeypatch.setattr(C.ops, "lincomb", lambda x, y, a, c: bc(a) * x + bc(c) * y)
    monkeypatch.setattr(C.ops, "guidance_combine",
                        lambda x0m, g, d, coef, c_in=None: (x0m + bc(coef) * (bc(c_in) * g + d)).clip(-1, 1))
    torch.set_grad_enabled(True)
    for e, key in ((eta, f"stsl.{opname}.{sigma}"), (0.0, f"stsl.{opname}.{sigma}.eta0")):
        cm = Host(operator=Op(), measurement=meas, guidance="stsl", zeta=zeta, eta=e, num_hutchinson_samples=n)
        torch.manual_seed(seed)
        hat = cm(I.xt(64, sigma, seed=21), torch.tensor([sigma]))
        ref = torch.from_numpy(golden_stsl[key])
        err = (hat - ref).abs().max().item()
        assert err < 2e-3, (key, err)
    assert (torch.from_numpy(golden_stsl[f"stsl.{opname}.{sigma}"]) - ref).norm() / ref.norm() > 0.05   # the probes matter


def test_extra_sampler_host_coefficients():
    """Host-side fp32 coefficient helpers of the remaining samplers (k_diffusion/sampling.py of this package) against the oracle's
    0-dim tensor arithmetic (itself pinned to the reference's outputs in test_oracle_golden.py): ancestral step, geometric
    midpoint, linear-multistep coefficients."""
    import k_diffusion as K
    from oracle import sampler_ref
    sig = sampler_ref.get_sigmas_karras(10, 0.05, 20.0)
    hs = sig.numpy()
    for i in range(10):
        for eta in (1.0, 0.5, 0.0):
            d_ref, u_ref = sampler_ref.ancestral_step(sig[i], sig[i + 1], eta)
            d, u = K.sampling.get_ancestral_step(hs[i], hs[i + 1], eta)
            assert abs(float(d) - float(d_ref)) <= 1e-6 * max(1.0, float(d_ref)) and abs(float(u) - float(u_ref)) <= 1e-6 * max(1.0, float(u_ref))
        if sig[i + 1] > 0:
            mid_ref = float(sig[i].log().lerp(sig[i + 1].log(), 0.5).exp())
            assert abs(float(K.sampling._log_midpoint(hs[i], hs[i + 1])) - mid_ref) <= 2e-7 * mid_ref
        cur = min(i + 1, 4)
        for j in range(cur):
            assert K.sampling.linear_multistep_coeff(cur, hs, i, j) == pytest.approx(sampler_ref.lms_coefficient(cur, hs, i, j), rel=1e-9)
    with pytest.raises(ValueError):
        K.sampling.linear_multistep_coeff(4, hs, 1, 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        K.sampling.sample_dpmpp_2m(lambda x, s: x, torch.zeros(1, 3, 8, 8), sig)


@pytest.mark.parametrize("case", I.EXTRA_SAMPLER_CASES, ids=lambda c: c[0])
def test_extra_sampler_host_loops_against_reference_golden(case, golden_samplers, monkeypatch):
    """Host side of the remaining samplers (step coefficients, branch decisions, noise / callback plumbing) against the
    REFERENCE's own outputs, with the three update kernels replaced by torch stand-ins that exist only in this test (the
    product has no CPU path; the kernels themselves are checked on the GPU in test_guidance_gpu.py::test_extra_samplers)."""
    import k_diffusion as K
    from kdip import ops
    from oracle import sampler_ref
    f32 = lambda v: torch.tensor(np.float32(v))

    def lincomb3(x, a, y=None, b=0.0, z=None, c=0.0, out=None):
        r = f32(a) * x
        if y is not None:
            r = r + f32(b) * y
        if z is not None:
            r = r + f32(c) * z
        return r if out is None else out.copy_(r)

    def euler_step(x, denoised, sigma_hat, dt, want_d=False):
        d = (x - denoised) / f32(sigma_hat)
        return (x + d * f32(dt), d) if want_d else x + d * f32(dt)

    def churn_(x, eps, s_noise, sigma, sigma_hat):
        return x + eps * f32(s_noise) * f32(np.sqrt(np.float32(sigma_hat) ** 2 - np.float32(sigma) ** 2))

    monkeypatch.setattr(ops, "lincomb3", lincomb3)
    monkeypatch.setattr(ops, "euler_step", euler_step)
    monkeypatch.setattr(ops, "churn_", churn_)
    monkeypatch.setattr(K.sampling, "_prep", lambda x: x.detach().contiguous().float())
    name, kw = case
    base = name.split("/")[0]
    noises, sig = I.sampler_noises(), sampler_ref.get_sigmas_karras(I.SAMPLER_N, I.SAMPLER_SIGMA_MIN, I.SAMPLER_SIGMA_MAX)
    kw = dict(kw)
    if "ancestral" in name:
        by_sigma = {float(s): k for k, s in enumerate(sig[:-1])}
        kw["noise_sampler"] = lambda s, s_next: noises[by_sigma[float(s)]]
    elif base == "dpm_2":
        kw["noise_sampler"] = lambda i, x: noises[i]
    seen = []
    out = getattr(K.sampling, "sample_" + base)(I.SAMPLER_MODEL, I.sampler_start(), sig, disable=True,
                                                callback=lambda d: seen.append(d["i"]), **kw)
    ref = golden_samplers["sampler." + name]
    assert seen == list(range(I.SAMPLER_N))
    assert np.abs(out.numpy() - ref).max() <= 2e-6 * np.abs(ref).max()
