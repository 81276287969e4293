"""Drop-in for ``import lpips`` as the reference's sample scripts use it (sample_condition_openai.py:11,161:
``loss_fn_vgg = lpips.LPIPS(net='vgg').to(device)``): the same call on the libkdip kernels (kdip/lpips.py).
The package's downloaded weights are read from ``KDIP_LPIPS_WEIGHTS`` (a torch.save'd ``lpips.LPIPS(net='vgg').state_dict()``)."""
from kdip.lpips import LPIPS  # noqa: F401
