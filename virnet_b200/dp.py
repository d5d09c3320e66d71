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


class BucketedGradSync:
    """K11 of SURVEY.md §2.1: the gradient exchange of a data-parallel step, overlapped with the backward pass like DDP's
    reducer (train_denoising_syn.py:70-71,179) — but over the engine's ONE flat gradient buffer.

    The backward pass produces weight gradients in reverse layer order, i.e. from the END of the flat buffer towards its
    start.  Each time the engine reports that a level of the U-Net is complete (`bucket_ready`), the bucket's layers
    are converted from the split-K workspace to parameter layout and its slice of the flat buffer is sum-all-reduced,
    both on a side stream, while the main stream keeps running the remaining dgrad / wgrad kernels.  `finish` makes
    the main stream wait for the last (small) bucket.  The 1/world average and the per-sub-network norms stay fused in
    vk_adam_clip_step (grad_scale), which reads the reduced buffer once.

    Buckets of the denoising network (5): [up block 1 .. tail], [up block 0], [down level 2], [down level 1],
    [SNet, head, down level 0] — 2.4 MB for the last, exposed one; 42 MB in total."""

    def __init__(self, engine, group=None):
        self.engine, self.group = engine, group
        dev = engine.flat_params.device
        self.comm = torch.cuda.Stream(device=dev)
        self._ev = [torch.cuda.Event() for _ in range(32)]
        self._ei = 0
        self._hi = len(engine.layers)            # layers [hi, end) are already handed over
        self._works = []
        self.buckets_last_step = []

    def begin(self):
        self._hi = len(self.engine.layers)
        self._works = []
        self.buckets_last_step = []

    def _event(self):
        ev = self._ev[self._ei % len(self._ev)]
        self._ei += 1
        return ev

    def bucket_ready(self, i0: int, wg_stream):
        eng = self.engine
        i1 = self._hi
        if i0 >= i1:
            return
        main = torch.cuda.current_stream()
        ev = self._event()
        ev.record(main)                          # bias gradients written on the main stream (channel_sum)
        self.comm.wait_event(ev)
        if wg_stream is not None:
            ev2 = self._event()
            ev2.record(wg_stream)                # the bucket's weight-gradient kernels
            self.comm.wait_event(ev2)
        begin, end = eng.layer_flat_range(i0, i1)
        with torch.cuda.stream(self.comm):
            eng.unpack_range(i0, i1)
            self._works.append(dist.all_reduce(eng.flat_grads[begin:end], op=dist.ReduceOp.SUM, group=self.group,
                                               async_op=True))
        self.buckets_last_step.append((i0, i1, begin, end))
        self._hi = i0

    def finish(self):
        assert self._hi == 0, "a bucket was never handed over"
        main = torch.cuda.current_stream()
        for w in self._works:
            w.wait()                             # stream-level wait of the current (main) stream, no host block on NCCL
        main.wait_stream(self.comm)
        self._works = []
