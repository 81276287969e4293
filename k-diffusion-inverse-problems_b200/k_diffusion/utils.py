"""k_diffusion/utils.py — the helpers the sampling path and its data front end use:
``append_dims`` (:40-45), ``to_pil_image`` / ``from_pil_image`` (:16-31) and ``FolderOfImages`` (:274-297).

``FolderOfImages`` keeps the reference's contract (recursive, sorted, RGB, item = 1-tuple after ``transform``) so the sample
scripts' ``DataLoader(FolderOfImages(location, transform=tf), batch_size)`` works unchanged; ``kdip.data.ImageBatchLoader`` is the
batched route that decodes on the GPU (pinned uint8 staging -> one copy -> kdip_images_u8_to_f32)."""
from pathlib import Path

from torch.utils import data


def append_dims(x, target_dims):
    """Appends dimensions to the end of a tensor until it has target_dims dimensions."""
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


def from_pil_image(x):
    """PIL image -> [C,H,W] fp32 in [-1,1] (utils.py:16-21: to_tensor, add a channel axis for 2-D, x*2-1)."""
    import numpy as np
    import torch
    t = torch.from_numpy(np.array(x))
    t = t[None] if t.ndim == 2 else t.permute(2, 0, 1)
    return t.to(torch.float32).div(255) * 2 - 1


def to_pil_image(x):
    """[1,C,H,W] / [C,H,W] tensor in [-1,1] -> PIL image (utils.py:24-31); CUDA tensors are quantised by kdip_images_f32_to_u8."""
    from PIL import Image
    if x.ndim == 4:
        assert x.shape[0] == 1
        x = x[0]
    if x.shape[0] == 3 and x.is_cuda:
        from kdip import ops
        return Image.fromarray(ops.images_f32_to_u8(x[None])[0].cpu().numpy(), mode="RGB")
    u8 = ((x.detach().float().cpu().clamp(-1, 1) + 1) / 2).mul(255).byte()
    if u8.shape[0] == 1:
        return Image.fromarray(u8[0].numpy(), mode="L")
    return Image.fromarray(u8.permute(1, 2, 0).contiguous().numpy(), mode="RGB")


class FolderOfImages(data.Dataset):
    """Every image file below ``root`` (recursive, sorted by path); no classes / targets."""

    IMG_EXTENSIONS = {'.jpg', '.jpeg', '.png', '.ppm', '.bmp', '.pgm', '.tif', '.tiff', '.webp'}

    def __init__(self, root, transform=None):
        super().__init__()
        self.root = Path(root)
        self.transform = transform if transform is not None else (lambda im: im)
        self.paths = sorted(p for p in self.root.rglob('*') if p.suffix.lower() in self.IMG_EXTENSIONS)

    def __repr__(self):
        return f'FolderOfImages(root="{self.root}", len: {len(self)})'

    def __len__(self):
        return len(self.paths)

    def load(self, key):
        from PIL import Image
        with open(self.paths[key], 'rb') as f:
            return Image.open(f).convert('RGB')

    def __getitem__(self, key):
        return self.transform(self.load(key)),
