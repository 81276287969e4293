"""Batched image front end: files -> [B,3,H,W] fp32 in [-1,1] on the GPU.

The reference feeds one image at a time through ``DataLoader(FolderOfImages(..., transform=ToTensor -> x*2-1), batch_size=1)``
and ``x0.to(device)`` (sample_condition_openai.py:139-163), i.e. it ships fp32 (12 bytes / pixel) over PCIe.  Here a batch is
decoded by PIL into ONE pinned uint8 HWC staging buffer (3 bytes / pixel), copied once, and expanded on the device by
``kdip_images_u8_to_f32`` - bit-identical to ToTensor followed by x*2-1.  Ranks take contiguous shards of the file list
(k_diffusion/evaluation.py:54).
"""
import numpy as np
import torch

from . import ops


class ImageBatchLoader:
    def __init__(self, dataset, batch_size, device="cuda", accelerator=None, drop_last=False):
        """``dataset``: a k_diffusion.utils.FolderOfImages (its ``transform`` is NOT applied: the device kernel is the transform)."""
        if not torch.cuda.is_available():
            raise RuntimeError("kdip.data.ImageBatchLoader decodes on the GPU; there is no CPU fallback")
        self.ds, self.bs, self.device, self.drop_last = dataset, int(batch_size), torch.device(device), drop_last
        lo, hi = (0, len(dataset)) if accelerator is None else accelerator.shard(len(dataset))
        self.index = list(range(lo, hi))
        self._stage = None
        self._copied = None     # CUDA event recorded after the last H2D copy out of the pinned staging buffer

    def __len__(self):
        n = len(self.index)
        return n // self.bs if self.drop_last else -(-n // self.bs)

    def _staging(self, B, H, W):
        if self._stage is None or self._stage.shape != (B, H, W, 3):
            self._stage = torch.empty(B, H, W, 3, dtype=torch.uint8).pin_memory()
        return self._stage

    def __iter__(self):
        for k in range(len(self)):
            keys = self.index[k * self.bs:(k + 1) * self.bs]
            first = np.array(self.ds.load(keys[0]))
            H, W = first.shape[:2]
            stage = self._staging(len(keys), H, W)
            if self._copied is not None:
                self._copied.synchronize()      # the previous batch's DMA still reads the staging buffer until this event
            stage[0].copy_(torch.from_numpy(first))
            for i, key in enumerate(keys[1:], 1):
                im = np.array(self.ds.load(key))
                if im.shape[:2] != (H, W):
                    raise ValueError(f"{self.ds.paths[key]}: {im.shape[:2]} differs from the batch's {(H, W)}")
                stage[i].copy_(torch.from_numpy(im))
            dev = stage.to(self.device, non_blocking=True)
            self._copied = torch.cuda.Event()
            self._copied.record()
            yield ops.images_u8_to_f32(dev),
