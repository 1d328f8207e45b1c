"""Data parallelism of the hot path: one process per GPU, the (already ordered)
batch is split into contiguous shards, site tensors are replicated, and the only exchange is ONE
all-reduce (sum) of the packed gradient buffer plus the scalar loss per step -- K6.

The reference has no multi-device code at all (tn4ml/models/model.py picks one jax device), so
this module is new; it keeps ``train_step`` semantics: every rank ends the step with identical
parameters because it applies the same optimizer update to the same summed gradient.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [lo, hi) of an n-row batch owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, targets=None, group=None):
    """This rank's contiguous shard of a batch (and of its targets)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(x), rank, world)
    return (x[lo:hi], None if targets is None else targets[lo:hi])


def allreduce_step(loss_sum: torch.Tensor, grads: torch.Tensor, group=None):
    """Sum the local loss sum and the packed gradient over the batch shards (in place)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group)
    return loss_sum, grads


class OverlappedAllReduce:
    """The step's gradient all-reduce, issued per site range while the rest of the weight-gradient work is still running.

    The backward pass runs in ``nparts`` site ranges (``tnb_backward_part``); after part k an event is recorded on the compute
    stream, and the NCCL all-reduce of that range's slice of the packed gradient is issued on a communication stream that
    waits for the event -- so the transfer of range k overlaps the GEMMs of range k+1 and only the last slice's collective
    (1/nparts of the bytes) is exposed.  ``finish`` makes the compute stream wait for every collective of the step.
    """

    def __init__(self, group=None, device=None, nparts: int = 4, grad_bytes: int | None = None):
        # Small gradients are not worth splitting: at cfg 4 (1.8 MB, 0.3 ms of work per rank on 8 GPUs) four ranges cost more
        # in launches and collective latency than they hide (0.81 ms per step against 0.58 ms with one trailing all-reduce,
        # profiles/r02g_scaling.md), so below 8 MB the step keeps one range.
        if grad_bytes is not None and grad_bytes < (8 << 20):
            nparts = 1
        self.group, self.nparts = group, nparts
        self.stream = torch.cuda.Stream(device)
        self._works = []

    def reduce(self, tensor: torch.Tensor):
        """All-reduce (sum, in place) ``tensor`` once everything issued so far on the current stream has finished."""
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self._works.append(dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        for w in self._works:
            w.wait()  # stream-level: the current stream waits for the collective, the host does not block
        self._works = []
