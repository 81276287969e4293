"""One process per GPU over torch.distributed (NCCL on NVLink 5 / NVSwitch; gloo for CPU tests).

``Accelerator`` is the small slice of ``accelerate.Accelerator`` the sampling path uses
(sample_condition_openai.py:124, k_diffusion/evaluation.py:53-63): ``device``, ``num_processes``, ``process_index``,
``is_main_process``, ``is_local_main_process`` and ``gather``.  Images are independent units: the batch is sharded in
contiguous blocks (rank r owns images [r*ceil(n/G), ...)), weights / OTF tables / masks are replicated, and the ONLY
communication is one all-gather of the finished samples — the last sampler update has already written them into a
contiguous [b,3,H,W] tensor, which NCCL reads in place (no staging copy).
"""
import os

import torch
import torch.distributed as dist


class Accelerator:
    def __init__(self, backend=None, device=None):
        self.num_processes = int(os.environ.get("WORLD_SIZE", "1"))
        self.process_index = int(os.environ.get("RANK", "0"))
        self.local_process_index = int(os.environ.get("LOCAL_RANK", "0"))
        if device is None:
            device = torch.device("cuda", self.local_process_index) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        if self.num_processes > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            backend = backend or ("nccl" if self.device.type == "cuda" else "gloo")
            kw = {"device_id": self.device} if backend == "nccl" else {}
            dist.init_process_group(backend, rank=self.process_index, world_size=self.num_processes, **kw)

    @property
    def is_main_process(self):
        return self.process_index == 0

    @property
    def is_local_main_process(self):
        return self.local_process_index == 0

    def shard(self, n):
        """Contiguous block of global indices owned by this rank (evaluation.py:54: ceil(n / world) per process)."""
        per = -(-n // self.num_processes)
        lo = min(n, self.process_index * per)
        return lo, min(n, lo + per)

    def gather(self, tensor):
        """All-gather along dim 0 (accelerate's ``gather``): every rank receives the concatenation in rank order."""
        if self.num_processes == 1:
            return tensor
        tensor = tensor.contiguous()
        out = torch.empty((self.num_processes * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
        dist.all_gather_into_tensor(out, tensor)
        return out

    def barrier(self):
        if self.num_processes > 1:
            dist.barrier()

    def max_over_ranks(self, value):
        if self.num_processes == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
