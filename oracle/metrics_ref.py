"""Oracle: dataset transform, PNG quantisation and the PSNR / SSIM of the sample scripts (test infrastructure only).

Follows sample_condition_openai.py:41-49 (compute_metrics), :140-144 (dataset transform), k_diffusion/utils.py:24-31
(to_pil_image) and analytic_variance.py:124-131.  Third-party arithmetic restated because scikit-image is not installed in this
container (**parity unpinned** for SSIM: no reference run of skimage is possible here; PSNR is the textbook formula):
  * skimage.metrics.peak_signal_noise_ratio(a, b, data_range=1): both images to float64, 10 log10(1 / mean((a-b)^2)).
  * skimage.metrics.structural_similarity(a, b, channel_axis=0, data_range=1) with its defaults: win_size 7,
    scipy.ndimage.uniform_filter per channel in the image's own float type, use_sample_covariance=True
    (cov_norm = 49/48), K1 = 0.01, K2 = 0.03, S = ((2 ux uy + C1)(2 vxy + C2)) / ((ux^2 + uy^2 + C1)(vx + vy + C2)),
    mean (float64) over the map cropped by (win_size-1)//2 = 3 on each side, then the mean over channels.
"""
import numpy as np
import torch
from scipy.ndimage import uniform_filter


def to_tensor_pm1(u8_hwc):
    """torchvision ToTensor (HWC uint8 -> CHW float32 / 255) followed by x*2-1."""
    x = torch.from_numpy(np.array(u8_hwc, dtype=np.uint8)).permute(2, 0, 1).to(torch.float32).div(255)
    return x * 2 - 1


def to_u8(x_chw):
    """to_pil_image: (clamp(x,-1,1)+1)/2 -> torchvision: mul(255).byte() -> HWC."""
    return ((x_chw.clamp(-1, 1) + 1) / 2).mul(255).byte().permute(1, 2, 0).numpy()


def to_eval(x):
    return (x / 2 + 0.5).clip(0, 1)


def psnr(x0, hat_x0):
    a, b = to_eval(x0).numpy().astype(np.float64), to_eval(hat_x0).numpy().astype(np.float64)
    return float(10 * np.log10(1.0 / np.mean((a - b) ** 2)))


def ssim(x0, hat_x0, dtype=np.float32):
    a, b = to_eval(x0).numpy().astype(dtype), to_eval(hat_x0).numpy().astype(dtype)
    NP, C1, C2, pad = 49, 0.01 ** 2, 0.03 ** 2, 3
    cov_norm = NP / (NP - 1)
    per_channel = []
    for c in range(a.shape[0]):
        x, y = a[c], b[c]
        ux, uy = uniform_filter(x, size=7), uniform_filter(y, size=7)
        uxx, uyy, uxy = uniform_filter(x * x, size=7), uniform_filter(y * y, size=7), uniform_filter(x * y, size=7)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        per_channel.append(S[pad:-pad, pad:-pad].mean(dtype=np.float64))
    return float(np.mean(per_channel))


def recon_mse(denoise_fn, batches, sigmas, noise_fn):
    """analytic_variance.py:117-137: for every sigma (the trailing 0 included) and every batch, mse of the denoiser on x0 + n*sigma.
    denoise_fn(x, sigma[B]) -> hat_x0; noise_fn(i, j, x0) -> the N(0, I) draw of (sigma i, batch j)."""
    errors = torch.zeros(len(sigmas), len(batches))
    mse_list = []
    for i, sigma in enumerate(sigmas):
        mse = 0
        for j, x0 in enumerate(batches):
            hat = denoise_fn(x0 + noise_fn(i, j, x0) * sigma, sigma.repeat(x0.shape[0]))
            cur = (x0 - hat).pow(2).mean()
            errors[i, j] = cur
            mse = mse + cur
        mse_list.append(mse / len(batches))
    return {"sigmas": sigmas, "mse_list": torch.stack(mse_list), "errors": errors}
