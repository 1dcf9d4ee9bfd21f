"""Image-sharded data parallelism for the hot path.

One process per GPU (``torchrun``).  a1–a6 have no cross-image dependency, so ranks work on
``images[rank::world]`` independently; the only exchange is the centroid 'mean' pass: every rank accumulates
``acc [C, D+1]`` = (sum of per-image class means, number of contributing images) and ONE all-reduce (NCCL over
NVLink on GPUs, gloo in the CPU tests) of that 19 x 2049 fp32 buffer finishes the pass.  This equals the
reference's sequential running mean (calc_centroids.py:157-161) while every class count stays below the 3000
clamp, i.e. for one pass over the 2975 Cityscapes training images.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int):
    """Indices of the items rank ``rank`` owns: ``range(n_items)[rank::world]``."""
    return list(range(n_items))[rank::world]


def new_mean_accumulator(class_numbers: int, feat_dim: int, device) -> torch.Tensor:
    return torch.zeros((class_numbers, feat_dim + 1), dtype=torch.float32, device=device)


def finish_mean_pass(acc: torch.Tensor, group=None):
    """All-reduce ``acc`` (in place, SUM) and return ``(objective_vectors [C,D], objective_vectors_num [C])``.

    ``objective_vectors[c] = sum / max(n, 1)``; ``objective_vectors_num[c] = min(n, 3000)`` (the reference's clamp).
    Works on CUDA tensors with the NCCL backend and on CPU tensors with gloo (host-logic tests).
    """
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    d = acc.shape[1] - 1
    n = acc[:, d]
    vectors = acc[:, :d] / n.clamp(min=1.0).unsqueeze(1)
    return vectors, n.clamp(max=3000.0)
