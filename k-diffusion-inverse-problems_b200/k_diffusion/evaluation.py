"""Batch driver + the path's only collective (k_diffusion/evaluation.py:53-63).

``compute_features(accelerator, sample_fn, extractor_fn, n, batch_size)`` keeps the reference signature.  Each rank
samples its own ceil(n / world) images in batches and the finished samples are all-gathered ONCE per batch
(``accelerator.gather`` in the reference).  ``accelerator`` may be an ``accelerate.Accelerator`` (duck-typed:
``num_processes``, ``is_main_process``, ``gather``) or the torch.distributed-backed ``kdip.dist.Accelerator``
(NCCL over NVLink; one process per GPU).
"""
import math

import torch

try:
    from tqdm.auto import trange
except Exception:  # pragma: no cover
    def trange(a, b=None, c=1, disable=None):
        return range(a, b, c)


def compute_features(accelerator, sample_fn, extractor_fn, n, batch_size):
    n_per_proc = math.ceil(n / accelerator.num_processes)
    feats_all = []
    try:
        for i in trange(0, n_per_proc, batch_size, disable=not accelerator.is_main_process):
            cur_batch_size = min(n - i, batch_size)
            samples = sample_fn(cur_batch_size)[:cur_batch_size]
            feats_all.append(accelerator.gather(extractor_fn(samples)))
    except StopIteration:
        pass
    return torch.cat(feats_all)[:n]
