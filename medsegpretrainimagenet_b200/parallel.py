"""Data parallelism for the hot path: one process per GPU, NCCL over NVLink 5 / NVSwitch.

The reference's only strategy is single-process `nn.DataParallel` (train_model.py:192-194: replicate,
scatter, gather, reduce-add to GPU 0 every step, BatchNorm per replica).  Here every rank owns its shard of
the batch and three exchanges exist: BatchNorm statistics (functional._BnAct), batchwise
Dice sums (losses._Dice) and — this module — the parameter gradients, reduced in reverse-order buckets
that are launched as soon as their last gradient has been produced, so the NCCL kernels overlap the rest
of the backward pass.

Gradients live directly in the flat bucket storage (`param.grad` is a view), so no packing copies are
needed: autograd accumulates into the bucket, NCCL reduces it in place, the optimizer reads it.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("flat", "params", "pending", "work")

    def __init__(self, flat, params):
        self.flat, self.params = flat, params
        self.pending, self.work = 0, None


class GradReducer:
    """Bucketed, overlapped gradient averaging.

        reducer = GradReducer(model.parameters(), bucket_mb=32)
        loop:  reducer.zero_grad();  loss.backward();  reducer.finish();  optimizer.step()
    """

    def __init__(self, params, bucket_mb: float = 32.0, group=None, average: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.average = average
        params = [p for p in params if p.requires_grad]
        self.params = params
        self.buckets: List[_Bucket] = []
        self._bucket_of = {}
        self._hooks = []
        if self.world == 1:
            # single process: nothing to exchange — gradients stay the tensors the backward kernels produced
            # (`param.grad = None` before backward lets autograd adopt them without an accumulate kernel each)
            self.zero_grad()
            return
        # autograd produces gradients roughly in reverse parameter order: bucket in that order so the
        # first bucket to fill is the first whose all-reduce can start
        order = list(reversed(params))
        limit = int(bucket_mb * (1 << 20))
        cur, cur_bytes = [], 0
        for p in order:
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > limit or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        for b in self.buckets:
            for p in b.params:
                self._bucket_of[p] = b
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.zero_grad()

    def _close(self, params):
        total = sum(p.numel() for p in params)
        flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.buckets.append(_Bucket(flat, params))

    @property
    def grad_bytes(self) -> int:
        return sum(b.flat.numel() * b.flat.element_size() for b in self.buckets)

    def zero_grad(self) -> None:
        """Replaces optimizer.zero_grad(): keeps `param.grad` aliased to the bucket storage."""
        if self.world == 1:
            for p in self.params:
                p.grad = None
            return
        for b in self.buckets:
            b.flat.zero_()
            b.pending = len(b.params)
            b.work = None

    def _on_grad(self, p) -> None:
        b = self._bucket_of[p]
        b.pending -= 1
        if b.pending == 0 and self.world > 1:
            # async_op: NCCL's own stream waits for the producer stream, then reduces while the
            # remaining backward kernels keep running on the compute stream
            if self.average and dist.get_backend(self.group) == "nccl":
                b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            else:
                b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> None:
        """Call after backward: waits for the in-flight reductions (and launches any bucket whose
        parameters received no gradient this step, e.g. frozen branches)."""
        if self.world == 1:
            return
        for b in self.buckets:
            if b.work is None:
                op = dist.ReduceOp.AVG if (self.average and dist.get_backend(self.group) == "nccl") \
                    else dist.ReduceOp.SUM
                b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
        for b in self.buckets:
            b.work.wait()
            if self.average and dist.get_backend(self.group) != "nccl":
                b.flat.div_(self.world)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shard_rows(n_global: int, rank: int, world: int):
    """Rank-local row range [lo, hi) of a global batch (rank r gets rows [r*B, (r+1)*B))."""
    per = n_global // world
    return rank * per, (rank + 1) * per
