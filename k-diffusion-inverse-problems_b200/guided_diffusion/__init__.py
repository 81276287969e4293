"""guided_diffusion — B200-native mirror of the reference's ADM UNet + DDPM schedule interface (hot path only).

Same import paths and names as /root/reference/guided_diffusion for the pieces the guided-sampling path touches:
``unet.UNetModel``, ``gaussian_diffusion.GaussianDiffusion`` (schedule constants, ``p_mean_variance``),
``respace.SpacedDiffusion`` and the factories of ``script_util``.  All device work runs in libkdip.
"""
