"""Checkpoint loading for one-process-per-GPU jobs (guided_diffusion/dist_util.py:24-74).

The reference reads the file on MPI rank 0 and broadcasts the bytes in 1 GiB chunks (dist_util.py:54-74) so that N ranks do
not hit the filesystem N times.  Same contract here over torch.distributed (NCCL when the process group is NCCL, gloo in
the CPU tests): rank 0 reads, everyone receives one uint8 tensor, ``torch.load`` runs on the bytes.  With no process group
(or world size 1) it is a plain ``torch.load``.  ``dev()`` mirrors dist_util.py:45-51.
"""
import io
import os

import torch as th
import torch.distributed as dist


def dev():
    """dist_util.py:45-51: the CUDA device of this rank (LOCAL_RANK), else cpu."""
    if th.cuda.is_available():
        return th.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    return th.device("cpu")


def _read_broadcast(path):
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        with open(path, "rb") as f:
            return f.read()
    rank = dist.get_rank()
    on_gpu = dist.get_backend() == "nccl"
    device = dev() if on_gpu else th.device("cpu")
    if rank == 0:
        with open(path, "rb") as f:
            data = f.read()
        n = th.tensor([len(data)], dtype=th.int64, device=device)
    else:
        data, n = None, th.zeros(1, dtype=th.int64, device=device)
    dist.broadcast(n, src=0)
    if rank == 0:
        buf = th.frombuffer(bytearray(data), dtype=th.uint8).to(device)
    else:
        buf = th.empty(int(n.item()), dtype=th.uint8, device=device)
    dist.broadcast(buf, src=0)
    return data if rank == 0 else buf.cpu().numpy().tobytes()


def load_state_dict(path, **kwargs):
    """Load a PyTorch file without redundant fetches across ranks (dist_util.py:54-74).  ``kwargs`` go to ``torch.load``
    (the sample scripts pass ``map_location="cpu"``, sample_condition_openai.py:130-132)."""
    kwargs.setdefault("weights_only", False)      # Lightning checkpoints carry plain-python hyper-parameters
    return th.load(io.BytesIO(_read_broadcast(path)), **kwargs)
