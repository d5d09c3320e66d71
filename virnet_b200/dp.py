"""Data-parallel plumbing of the training path (reference: DDP wrap at train_denoising_syn.py:70-71,
batch split at :133).  The hot path shards by patches; the only exchange step per iteration is ONE
sum all-reduce of the flat fp32 gradient buffer.  DDP's averaging (1/world) is not applied here: it is
folded into vk_adam_clip_step's `grad_scale`, which reads the reduced bucket once for
norm -> clip -> Adam.  Works on any backend (`nccl` on the B200 box, `gloo` in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def per_rank_batch(global_batch: int, world: int) -> int:
    """The reference splits its global batch: `batch_size // num_gpus` per rank (train_denoising_syn.py:133)."""
    if world <= 0 or global_batch < world:
        raise ValueError("global batch smaller than the number of ranks")
    return global_batch // world


def all_reduce_flat_grads(flat_grads: torch.Tensor, group=None) -> float:
    """Sum-all-reduce the flat gradient bucket in place; returns the factor (1/world) that turns the sum
    into DDP's average and must be applied by the consumer (vk_adam_clip_step grad_scale)."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / w


def broadcast_flat_params(flat_params: torch.Tensor, src: int = 0, group=None) -> None:
    """DDP broadcasts rank 0's parameters at wrap time (train_denoising_syn.py:70-71)."""
    if world_size(group) > 1:
        dist.broadcast(flat_params, src=src, group=group)
