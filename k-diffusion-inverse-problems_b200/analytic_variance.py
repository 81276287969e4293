"""Monte-Carlo estimate of the optimal isotropic posterior variance per noise level (analytic_variance.py:113-139):
``mse(sigma) = E || x0 - D(x0 + sigma n, sigma) ||^2 / numel`` over a data subset, saved as ``{'sigmas', 'mse_list', 'errors'}``
(the ``recon_mse.pt`` that ``x0_cov_type='analytic'`` reads, condition/condition.py:250-256).

UNet FORWARD only: per (sigma, batch) one noising pass (kdip_lincomb), one UNet evaluation with continuous t
(OpenAIDenoiser.forward, k_diffusion/external.py:111-132 - no ``.long()``, no clamp) and one fused denoise + squared-error
reduction (kdip_denoise_sqerr).  Nothing is synchronised per batch: the per-image fp64 sums stay on the device until the end.
Multi-GPU: batches are independent; ranks take disjoint batches and the [n_sigma, n_batch] error table is all-reduced once.
"""
import numpy as np
import torch

from kdip import ops


@torch.no_grad()
def estimate_recon_mse(denoiser, batches, sigmas, accelerator=None, noise_fn=None):
    """denoiser: k_diffusion.external.OpenAIDenoiser on a kdip UNetModel; batches: list of [B,3,H,W] CUDA tensors (equal B, as
    DataLoader(drop_last=True)); sigmas: 1-D tensor (get_sigmas_karras output, trailing 0 included as in the reference).
    noise_fn(i, j, x0) -> N(0, I) draw for (sigma i, batch j); default torch.randn_like.  -> dict like recon_mse_test.pt."""
    eng = denoiser.inner_model.engine()
    sig_host = [float(s) for s in sigmas.detach().cpu().tolist()]
    n_sig, n_b = len(sig_host), len(batches)
    rank, world = (0, 1) if accelerator is None else (accelerator.process_index, accelerator.num_processes)
    dev = batches[0].device
    sums = torch.zeros(n_sig, n_b, dtype=torch.float64, device=dev)
    for i, s in enumerate(sig_host):
        s32 = np.float32(s)
        c_in = float(np.float32(1) / np.sqrt(s32 * s32 + np.float32(1)))
        t = denoiser.sigma_to_t_host(s)
        for j, x0 in enumerate(batches):
            if j % world != rank:
                continue
            B = x0.shape[0]
            noise = torch.randn_like(x0) if noise_fn is None else noise_fn(i, j, x0).to(dev, torch.float32)
            one, sg = torch.ones(B, device=dev), torch.full((B,), s, device=dev, dtype=torch.float32)
            xn = ops.lincomb(x0, noise, one, sg)                                  # x0 + n * sigma
            out = eng.forward(xn, torch.full((B,), t, device=dev), x_scale=torch.full((B,), c_in, device=dev))
            sums[i, j] = ops.denoise_sqerr(out, xn, x0, sg).sum() / (B * x0[0].numel())
    if world > 1:
        torch.distributed.all_reduce(sums)
    errors = sums.to(torch.float32).cpu()
    return {"sigmas": sigmas.detach().cpu(), "mse_list": errors.mean(dim=1), "errors": errors}
