"""GPU parity of the front end either side of the sampling path through the C ABI: 8-bit <-> fp32
image transforms (bit-exact), PSNR / SSIM, the analytic-variance Monte-Carlo estimate, and the checkpoint formats end to end."""
import os

import numpy as np
import pytest
import torch

import inputs as I

pytestmark = pytest.mark.gpu


def test_u8_f32_transforms_bit_exact():
    """kdip_images_u8_to_f32 == ToTensor then x*2-1; kdip_images_f32_to_u8 == (clamp(x,-1,1)+1)/2 -> mul(255).byte(): bit-exact,
    including every byte value, values outside [-1,1], exact +-1 and the k/255 boundaries where truncation bites."""
    from kdip import ops
    from oracle import metrics_ref
    rng = np.random.RandomState(0)
    u8 = rng.randint(0, 256, size=(3, 16, 24, 3), dtype=np.uint8)
    u8[0].reshape(-1)[:256] = np.arange(256, dtype=np.uint8)            # every byte value
    got = ops.images_u8_to_f32(torch.from_numpy(u8).cuda()).cpu()
    ref = torch.stack([metrics_ref.to_tensor_pm1(u8[b]) for b in range(3)])
    assert torch.equal(got, ref)
    x = torch.randn(3, 3, 16, 24, generator=torch.Generator().manual_seed(1)) * 0.8
    x[0, 0, 0, :8] = torch.tensor([-1.0, 1.0, -1.5, 2.0, 0.0, -0.0, 1.0 - 2 ** -24, -1.0 + 2 ** -24])
    x[1] = ref[1]                                                        # exact k/255*2-1 values
    back = ops.images_f32_to_u8(x.cuda()).cpu().numpy()
    ref_back = np.stack([metrics_ref.to_u8(x[b]) for b in range(3)])
    assert np.array_equal(back, ref_back)
    # full size, ragged batch of one
    big = rng.randint(0, 256, size=(1, 256, 256, 3), dtype=np.uint8)
    assert torch.equal(ops.images_u8_to_f32(torch.from_numpy(big).cuda()).cpu()[0], metrics_ref.to_tensor_pm1(big[0]))
    with pytest.raises(ValueError):
        ops.images_u8_to_f32(torch.zeros(1, 3, 5, 3, dtype=torch.uint8, device="cuda"))   # H*W not a multiple of 4


def test_image_batch_loader_matches_reference_transform(tmp_path):
    """Files -> pinned uint8 staging -> one copy -> device decode == the reference's per-image ToTensor / x*2-1 route."""
    from PIL import Image
    from kdip.data import ImageBatchLoader
    from oracle import metrics_ref
    import k_diffusion as K
    rng = np.random.RandomState(5)
    arrs = []
    for i in range(5):
        a = rng.randint(0, 256, size=(64, 64, 3), dtype=np.uint8)
        Image.fromarray(a).save(os.path.join(str(tmp_path), f"{i:03d}.png"))
        arrs.append(a)
    ds = K.utils.FolderOfImages(str(tmp_path))
    dl = ImageBatchLoader(ds, batch_size=2)
    assert len(dl) == 3
    batches = [b for (b,) in dl]
    assert [b.shape[0] for b in batches] == [2, 2, 1]                    # ragged last batch
    got = torch.cat(batches).cpu()
    assert torch.equal(got, torch.stack([metrics_ref.to_tensor_pm1(a) for a in arrs]))
    assert len(ImageBatchLoader(ds, batch_size=2, drop_last=True)) == 2
    # to_pil_image of a CUDA tensor goes through the device quantiser and equals the host route
    pil = K.utils.to_pil_image(batches[0][:1])
    assert np.array_equal(np.asarray(pil), metrics_ref.to_u8(batches[0][0].cpu()))


@pytest.mark.parametrize("size", [64, 256])
def test_psnr_ssim(size):
    """compute_metrics of sample_condition_openai.py:41-49 (PSNR, SSIM) per image of a batch vs the oracle's scikit-image
    restatement.  Tolerances: PSNR 1e-9 dB (fp64 sums, different order); SSIM 2e-5 (the oracle filters in float32 like skimage
    does for float32 input, the kernel keeps the window sums in fp64)."""
    from kdip import ops
    from oracle import metrics_ref
    g = torch.Generator().manual_seed(11)
    x0 = I.image(size, batch=3, seed=1)
    hat = x0.clone()
    hat[0] += 0.3 * torch.randn(3, size, size, generator=g)              # noisy, partly outside [-1,1] (to_eval clips)
    hat[1] = hat[1] * 0.5 + 0.1                                           # smooth distortion
    hat[2] += 1e-3 * torch.randn(3, size, size, generator=g)             # nearly identical (SSIM -> 1, PSNR large)
    p = ops.psnr(x0.cuda(), hat.cuda()).cpu()
    s = ops.ssim(x0.cuda(), hat.cuda()).cpu()
    for b in range(3):
        rp, rs = metrics_ref.psnr(x0[b], hat[b]), metrics_ref.ssim(x0[b], hat[b])
        print(f"image {b} @ {size}: psnr {p[b].item():.6f} (ref {rp:.6f})  ssim {s[b].item():.7f} (ref {rs:.7f})")
        assert abs(p[b].item() - rp) < 1e-9 * max(1.0, abs(rp))
        assert abs(s[b].item() - rs) < 2e-5
    assert ops.ssim(x0.cuda(), x0.cuda()).cpu().sub(1).abs().max() < 1e-12


def test_analytic_variance_estimate(golden_small):
    """analytic_variance.py:113-139 on the tiny UNet: 4 noise levels (+ the trailing sigma = 0 the reference also evaluates) x
    3 batches of 2 images, injected noise, vs the oracle loop on the fp32 CPU UNet.  Tolerance 5e-2 relative per entry (one bf16
    UNet forward, error in eps ~1e-2 relative enters the squared error twice); sigma = 0 gives exactly 0."""
    from oracle import diffusion_ref, metrics_ref, sampler_ref, unet_ref
    from guided_diffusion.script_util import create_gaussian_diffusion
    from guided_diffusion.unet import UNetModel
    from k_diffusion.external import OpenAIDenoiser
    import analytic_variance
    cfg = unet_ref.tiny_config()
    sd = unet_ref.init_state_dict(cfg, seed=0)
    model = UNetModel(image_size=64, in_channels=3, model_channels=64, out_channels=6, num_res_blocks=1,
                      attention_resolutions=cfg.attention_ds(), channel_mult=cfg.resolved_channel_mult(), num_head_channels=64,
                      use_scale_shift_norm=True, resblock_updown=True)
    model.load_state_dict(sd, strict=True)
    den = OpenAIDenoiser(model.eval().cuda(), create_gaussian_diffusion(learn_sigma=True), device="cuda")
    sigmas = sampler_ref.get_sigmas_karras(4, 0.05, 20.0)
    batches = [I.image(64, batch=2, seed=100 + j) for j in range(3)]
    g = torch.Generator().manual_seed(7)
    noises = {(i, j): torch.randn(2, 3, 64, 64, generator=g) for i in range(len(sigmas)) for j in range(3)}
    got = analytic_variance.estimate_recon_mse(den, [b.cuda() for b in batches], sigmas, noise_fn=lambda i, j, x: noises[(i, j)])
    sched = diffusion_ref.Schedule()

    def ref_denoise(x, sigma):           # OpenAIDenoiser.forward (external.py:111-132): continuous t, eps = first 3 channels
        c_out, c_in = diffusion_ref.get_scalings(sigma)
        eps = unet_ref.unet_forward(sd, cfg, x * c_in.view(-1, 1, 1, 1), sched.sigma_to_t(sigma))[:, :3]
        return x + eps * c_out.view(-1, 1, 1, 1)

    with torch.no_grad():
        ref = metrics_ref.recon_mse(ref_denoise, batches, sigmas, lambda i, j, x: noises[(i, j)])
    print("mse_list", got["mse_list"].tolist(), "ref", ref["mse_list"].tolist())
    assert torch.equal(got["sigmas"], sigmas) and got["errors"].shape == (5, 3)
    assert torch.allclose(got["errors"][:-1], ref["errors"][:-1], rtol=5e-2, atol=0)
    assert torch.allclose(got["mse_list"][:-1], ref["mse_list"][:-1], rtol=5e-2, atol=0)
    assert float(got["mse_list"][-1]) == 0.0 and float(ref["mse_list"][-1]) == 0.0
    # the file is consumable by x0_cov_type='analytic' (condition.py:250-256): nearest-sigma lookup returns a scalar
    idx = (got["sigmas"] - 0.05).abs().argmin()
    assert got["mse_list"][idx].dim() == 0


def test_lightning_checkpoint_end_to_end(tmp_path, golden_v2):
    """ffhq_dwt.ckpt layout -> train_openai.OpenAIDenoiser.load_from_checkpoint(path).model_ema (sample_condition_openai_v2.py:117)
    -> the same model_output / logvar_ot as the REFERENCE produced with these weights (golden_v2)."""
    from test_frontend_cpu import V2_MODEL_CONFIG, _lightning_ckpt
    from train_openai import OpenAIDenoiser
    path = _lightning_ckpt(str(tmp_path), I.v2_config(), V2_MODEL_CONFIG)[0]
    den = OpenAIDenoiser.load_from_checkpoint(path, map_location="cuda").model_ema.eval()
    sigma = I.V2_SIGMAS[0]
    xt = I.xt(64, sigma, seed=21).cuda()
    mo, lv, lvo = den(xt, torch.tensor([sigma]).cuda(), return_variance=True)
    ref_mo, ref_lvo = torch.from_numpy(golden_v2[f"v2.dwt.{sigma}.model_output"]), torch.from_numpy(golden_v2[f"v2.dwt.{sigma}.logvar_ot"])
    e_mo = ((mo.cpu() - ref_mo).norm() / ref_mo.norm()).item()
    e_lv = ((lvo.cpu() - ref_lvo).norm() / ref_lvo.norm()).item()
    print(f"checkpoint -> v2 denoiser: model_output rel-L2 {e_mo:.3e}, logvar_ot rel-L2 {e_lv:.3e}")
    assert e_mo < 1.5e-2 and e_lv < 1.5e-2


def test_sample_script_flow(tmp_path):
    """The main body of sample_condition_openai.py:112-210 executed on this package, end to end on files: OpenAI ``.pt`` checkpoint
    -> ``create_model_and_diffusion`` + ``load_state_dict(dist_util.load_state_dict(...))``; image folder -> ``FolderOfImages``
    (batched GPU decode); ``get_operator`` -> measurement; ``ConditionOpenAIDenoiser`` -> ``sample_heun`` driven by
    ``K.evaluation.compute_features``; PSNR / SSIM; ``to_pil_image(...).save``.  Structural checks only: 4-step trajectories of the
    synthetic tiny UNet are chaotic (module docstring of test_guidance_gpu.py); numerical parity of every stage is pinned by the
    other tests."""
    from PIL import Image
    import k_diffusion as K
    from condition.condition import ConditionOpenAIDenoiser
    from condition.measurements import get_operator
    from kdip import checkpoint, ops
    from kdip.data import ImageBatchLoader
    from kdip.dist import Accelerator
    from oracle import metrics_ref, unet_ref
    root = str(tmp_path)
    ckpt = os.path.join(root, "tiny.pt")
    torch.save(unet_ref.init_state_dict(unet_ref.tiny_config(), seed=0), ckpt)
    os.makedirs(os.path.join(root, "images"))
    truth = I.image(64, batch=3, seed=1)
    for i in range(3):
        K.utils.to_pil_image(truth[i:i + 1]).save(os.path.join(root, "images", f"{i:05d}.png"))
    config = {"model": {"openai": {"image_size": 64, "num_channels": 64, "num_res_blocks": 1, "attention_resolutions": "16,8",
                                   "channel_mult": "1,2,3,4"}, "sigma_min": 0.01, "sigma_max": 80.0},
              "dataset": {"location": os.path.join(root, "images")},
              "operator": {"name": "gaussian_blur", "in_shape": (1, 3, 64, 64), "kernel_size": 61, "intensity": 3.0, "sigma_s": 0.05}}
    accelerator = Accelerator()
    device = accelerator.device
    inner_model, diffusion = checkpoint.load_openai_unet(ckpt, config["model"]["openai"], device=device)
    operator = get_operator(device=device, **config["operator"])
    sigmas = K.sampling.get_sigmas_karras(4, config["model"]["sigma_min"], config["model"]["sigma_max"], rho=7., device=device)
    test_set = K.utils.FolderOfImages(config["dataset"]["location"])
    assert len(test_set) == 3
    hats, x0s, seen = [], [], 0
    for (x0,) in ImageBatchLoader(test_set, batch_size=2, device=device):          # batches of 2 and 1
        b = x0.shape[0]
        measurement = operator.forward(x0, flatten=True)
        assert measurement[0].shape == x0.shape and measurement[1].shape == (b, x0[0].numel())
        model = ConditionOpenAIDenoiser(inner_model=inner_model, diffusion=diffusion, x0_cov_type="convert", recon_mse=None,
                                        operator=operator, measurement=measurement, guidance="I", device=device,
                                        mle_sigma_thres=0.2).eval()

        def sample_fn(n):
            x = torch.randn([n, 3, 64, 64], device=device) * config["model"]["sigma_max"]
            return K.sampling.sample_heun(model, x, sigmas, disable=True)

        hat = K.evaluation.compute_features(accelerator, sample_fn, lambda x: x, b, b)
        assert hat.shape == x0.shape and torch.isfinite(hat).all() and hat.abs().max() <= 1.0 + 1e-5
        hats.append(hat)
        x0s.append(x0)
        seen += b
    assert seen == 3
    hat, x0 = torch.cat(hats), torch.cat(x0s)
    assert torch.equal(x0.cpu(), torch.stack([metrics_ref.to_tensor_pm1(metrics_ref.to_u8(truth[i])) for i in range(3)]))   # decode
    psnr, ssim = ops.psnr(x0, hat).cpu(), ops.ssim(x0, hat).cpu()
    assert psnr.shape == (3,) and torch.isfinite(psnr).all() and torch.isfinite(ssim).all() and (ssim.abs() <= 1.0 + 1e-9).all()
    for i in range(3):
        assert abs(psnr[i].item() - metrics_ref.psnr(x0[i].cpu(), hat[i].cpu())) < 1e-6
        path = os.path.join(root, f"hat_{i:05d}.png")
        K.utils.to_pil_image(hat[i:i + 1]).save(path)
        assert np.array_equal(np.asarray(Image.open(path)), metrics_ref.to_u8(hat[i].cpu()))
    print("sample flow: psnr", [round(v, 2) for v in psnr.tolist()], "ssim", [round(v, 3) for v in ssim.tolist()])
