"""CPU checks of the front end either side of the sampling path: checkpoint formats, the image
dataset, the oracle's transform / metric restatements.  No CUDA compute here; the kernels are checked in test_frontend_gpu.py."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import inputs as I

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lightning_ckpt(tmp_path, cfg, model_config):
    """A checkpoint file laid out like train_openai.py's LightningModule writes it: model.* (training copy), model_ema.* (EMA
    copy, different values), schedule buffers, hyper_parameters."""
    from oracle import unet_ref
    from guided_diffusion.script_util import create_gaussian_diffusion
    sd_train, sd_ema = unet_ref.init_state_dict(cfg, seed=1), unet_ref.init_state_dict(cfg, seed=0)
    cov_w, cov_b = I.v2_out_cov(seed=9)
    ac = torch.tensor(create_gaussian_diffusion(learn_sigma=True).alphas_cumprod, dtype=torch.float32)
    sigmas = ((1 - ac) / ac) ** 0.5
    state = {}
    for copy, sd, scale in (("model", sd_train, 2.0), ("model_ema", sd_ema, 1.0)):
        for k, v in sd.items():
            state[f"{copy}.inner_model.{k}"] = v
        state[f"{copy}.out_cov.weight"], state[f"{copy}.out_cov.bias"] = cov_w * scale, cov_b * scale
        state[f"{copy}.sigmas"], state[f"{copy}.log_sigmas"] = sigmas, sigmas.log()
    path = os.path.join(tmp_path, "dwt.ckpt")
    torch.save({"state_dict": state, "hyper_parameters": {"model_config": model_config, "train_config": {"lr": 1e-4}},
                "epoch": 3, "global_step": 1234, "pytorch-lightning_version": "2.1.0"}, path)
    return path, sd_ema, sd_train, cov_w, cov_b


V2_MODEL_CONFIG = {"input_size": [64, 64], "input_channels": 3, "sigma_min": 0.01, "sigma_max": 80, "ortho_tf_type": "dwt",
                   "openai": {"image_size": 64, "num_channels": 128, "num_res_blocks": 1, "attention_resolutions": "16,8",
                              "channel_mult": "1,2,3,4"}}


def test_lightning_checkpoint_to_denoiser_v2(tmp_path):
    """Format 2 (ffhq_dwt.ckpt layout, train_openai.py:77-135) -> OpenAIDenoiserV2 copies, as sample_condition_openai_v2.py:117."""
    from kdip import checkpoint
    from train_openai import OpenAIDenoiser
    cfg = I.v2_config()
    path, sd_ema, sd_train, cov_w, cov_b = _lightning_ckpt(str(tmp_path), cfg, V2_MODEL_CONFIG)
    mod = OpenAIDenoiser.load_from_checkpoint(path, map_location="cpu")
    den = mod.model_ema.eval()
    assert den.ortho_tf_type == "dwt"
    got = den.inner_model.state_dict()
    assert list(got.keys()) == list(sd_ema.keys())
    assert all(torch.equal(got[k], sd_ema[k]) for k in sd_ema)
    assert torch.equal(den.out_cov.weight, cov_w) and torch.equal(den.out_cov.bias, cov_b)
    assert torch.equal(mod.model.out_cov.weight, cov_w * 2.0)                       # the training copy stays separate
    assert torch.equal(mod.model.inner_model.state_dict()["out.2.bias"], sd_train["out.2.bias"])
    assert mod.hparams["model_config"]["ortho_tf_type"] == "dwt"
    # malformed inputs are loud
    sd = checkpoint.unwrap_state_dict(torch.load(path, weights_only=False))
    with pytest.raises(KeyError, match="checkpoint has no"):
        checkpoint.split_denoiser_state_dict(sd, "ema")
    bad = {k: v for k, v in sd.items() if "out_cov" not in k}
    with pytest.raises(KeyError, match="out_cov"):
        checkpoint.split_denoiser_state_dict(bad, "model_ema")
    sd2 = dict(sd)
    sd2["model_ema.log_sigmas"] = sd["model_ema.log_sigmas"] + 0.1
    with pytest.raises(ValueError, match="log_sigmas"):
        checkpoint.build_denoiser_v2(V2_MODEL_CONFIG, sd2, device="cpu")
    with pytest.raises(KeyError, match="hyper_parameters"):
        p2 = os.path.join(str(tmp_path), "flat.pt")
        torch.save(sd_ema, p2)
        OpenAIDenoiser.load_from_checkpoint(p2)
    with pytest.raises(NotImplementedError):
        OpenAIDenoiser(V2_MODEL_CONFIG, {})


def test_openai_pt_checkpoint_to_unet(tmp_path):
    """Format 1 (flat UNetModel state_dict .pt): sample_condition_openai.py:128-132."""
    from oracle import unet_ref
    from kdip import checkpoint
    cfg = unet_ref.tiny_config()
    sd = unet_ref.init_state_dict(cfg, seed=0)
    path = os.path.join(str(tmp_path), "tiny.pt")
    torch.save(sd, path)
    over = {"image_size": 64, "num_channels": 64, "num_res_blocks": 1, "attention_resolutions": "16,8", "channel_mult": "1,2,3,4"}
    model, diffusion = checkpoint.load_openai_unet(path, over, device="cpu")
    got = model.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd) and not model.training
    assert len(diffusion.betas) == 1000
    sd.pop("out.2.bias")
    torch.save(sd, path)
    with pytest.raises(RuntimeError, match="out.2.bias"):
        checkpoint.load_openai_unet(path, over, device="cpu")


def test_load_state_dict_broadcast_two_rank_gloo(tmp_path):
    """dist_util.load_state_dict: rank 0 reads, the bytes are broadcast (dist_util.py:54-74) - world size 2 over gloo; the file is
    deleted from under rank 1 before it would read it, so it can only have come through the broadcast."""
    path = os.path.join(str(tmp_path), "w.pt")
    torch.save({"a": torch.arange(7.0), "b": {"c": torch.ones(2, 3)}}, path)
    script = f"""
import os, sys, time
sys.path[:0] = [{os.path.join(ROOT, 'k-diffusion-inverse-problems_b200')!r}]
import torch, torch.distributed as dist
from guided_diffusion import dist_util
rank = int(os.environ['RANK'])
dist.init_process_group('gloo', rank=rank, world_size=2)
path = {path!r} if rank == 0 else {path + '.not-there'!r}
sd = dist_util.load_state_dict(path, map_location='cpu')
assert torch.equal(sd['a'], torch.arange(7.0)) and torch.equal(sd['b']['c'], torch.ones(2, 3))
dist.barrier()
print('rank', rank, 'ok')
"""
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
        procs.append(subprocess.Popen([sys.executable, "-c", script], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "rank 0 ok" in outs[0] and "rank 1 ok" in outs[1]


def test_folder_of_images_and_pil_round_trip(tmp_path):
    """FolderOfImages (k_diffusion/utils.py:274-297): recursive, sorted, RGB, 1-tuples; from_pil_image / to_pil_image (:16-31)
    against the oracle's ToTensor / mul(255).byte() restatement."""
    from PIL import Image
    from oracle import metrics_ref
    import k_diffusion as K
    rng = np.random.RandomState(0)
    os.makedirs(os.path.join(str(tmp_path), "b", "deep"))
    imgs = {}
    for name in ("z.png", "a.png", os.path.join("b", "m.PNG"), os.path.join("b", "deep", "c.bmp")):
        arr = rng.randint(0, 256, size=(16, 20, 3), dtype=np.uint8)
        Image.fromarray(arr).save(os.path.join(str(tmp_path), name))
        imgs[name] = arr
    Image.fromarray(rng.randint(0, 256, size=(16, 20), dtype=np.uint8)).save(os.path.join(str(tmp_path), "gray.png"))
    open(os.path.join(str(tmp_path), "notes.txt"), "w").write("not an image")
    ds = K.utils.FolderOfImages(str(tmp_path), transform=lambda im: metrics_ref.to_tensor_pm1(np.asarray(im)))
    rel = [str(p.relative_to(tmp_path)) for p in ds.paths]
    assert rel == sorted(rel) and len(ds) == 5 and "notes.txt" not in rel
    for i, r in enumerate(rel):
        (x,) = ds[i]
        assert x.shape == (3, 16, 20) and x.dtype == torch.float32
        if r in imgs:
            assert torch.equal(x, metrics_ref.to_tensor_pm1(imgs[r]))
            assert torch.equal(K.utils.from_pil_image(Image.fromarray(imgs[r])), x)
            # to_pil_image(from_pil_image(im)) reproduces the bytes except where (v/255*2-1+1)/2*255 truncates below v
            back = np.asarray(K.utils.to_pil_image(x[None]))
            assert np.array_equal(back, metrics_ref.to_u8(x))
            assert np.abs(back.astype(int) - imgs[r].astype(int)).max() <= 1
    assert "FolderOfImages" in repr(ds)


def test_metric_oracle_known_answers():
    """PSNR / SSIM restatements on cases with known answers: identical images (SSIM 1), a constant offset (PSNR closed form),
    SSIM symmetric, SSIM of float32 vs float64 evaluation agree to 1e-5."""
    from oracle import metrics_ref
    x = I.image(64, batch=1, seed=1)[0]
    y = (x + 0.2 * torch.randn(x.shape, generator=torch.Generator().manual_seed(3))).clamp(-1, 1)
    assert metrics_ref.ssim(x, x) == pytest.approx(1.0, abs=1e-6)
    assert metrics_ref.ssim(x, y) == pytest.approx(metrics_ref.ssim(y, x), abs=1e-7)
    assert metrics_ref.ssim(x, y) == pytest.approx(metrics_ref.ssim(x, y, dtype=np.float64), abs=1e-5)
    assert 0.0 < metrics_ref.ssim(x, y) < 0.99
    c = torch.full_like(x, 0.1)
    d = torch.full_like(x, 0.3)           # to_eval: 0.55 vs 0.65 -> mse 0.01 -> 20 dB
    assert metrics_ref.psnr(c, d) == pytest.approx(20.0, abs=1e-4)
