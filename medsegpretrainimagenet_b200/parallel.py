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

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib


class _Bucket:
    __slots__ = ("flat", "params", "pending", "work")

    def __init__(self, flat, params):
        self.flat, self.params = flat, params
        self.pending, self.work = 0, None


class GradReducer:
    """Bucketed, overlapped gradient averaging.

        reducer = GradReducer(model.parameters(), bucket_mb=32)
        loop:  reducer.zero_grad();  loss.backward();  reducer.finish();  optimizer.step()

    Gradient accumulation (the reference's "virtual batch", train_model.py:53-55: `accumulation_scale` backward passes
    per optimizer step): run every micro-batch but the last under `with reducer.no_sync():` — their gradients only
    accumulate in the bucket storage — and call `finish()` after the last one, which is the only pass that exchanges.
    A second backward outside `no_sync()` without a `finish()` in between, or a `param.grad` that no longer aliases
    the bucket (e.g. after `optimizer.zero_grad(set_to_none=True)`; use `reducer.zero_grad()`), raises instead of
    silently reducing stale or empty buckets.
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
        self._sync = True
        self._grad_ptr = {}
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
                # the hook also fires for the convolution weights, whose gradient bypasses autograd's accumulation
                # (ops._WgradQueue adds it into the bucket view itself and hands autograd None): AccumulateGrad runs its
                # post hooks for an undefined gradient too (torch 2.11; tests/test_parallel_cpu.py pins that behaviour)
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.zero_grad()

    def _close(self, params):
        total = sum(p.numel() for p in params)
        flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            self._grad_ptr[p] = p.grad.data_ptr()
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

    def no_sync(self):
        """Context manager for every micro-batch of an accumulated step except the last: backward passes inside it
        add into the buckets without exchanging them."""
        reducer = self

        class _NoSync:
            def __enter__(self_inner):
                self_inner.prev, reducer._sync = reducer._sync, False
                return reducer

            def __exit__(self_inner, *exc):
                reducer._sync = self_inner.prev
                return False

        return _NoSync()

    def _launch(self, b) -> None:
        # async_op: NCCL's own stream waits for the producer stream, then reduces while the remaining backward
        # kernels keep running on the compute stream
        op = dist.ReduceOp.AVG if (self.average and dist.get_backend(self.group) == "nccl") else dist.ReduceOp.SUM
        b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def _flush_and_launch(self, b) -> None:
        """The bucket's queued weight-gradient unpacks, then its all-reduce.  With the wgrad side stream (ops.
        set_wgrad_stream) both go ON that stream — it first waits for the main stream (the bucket's BatchNorm / bias
        gradients), then carries the unpack behind the wgrad kernels it depends on and hands the bucket to NCCL, while
        the main stream's dgrad / BatchNorm-backward chain runs on without waiting for any of it."""
        from . import ops
        side = ops.wgrad_stream() if b.flat.is_cuda else None
        if side is None:
            ops.flush_wgrad()
            self._launch(b)
            return
        side.wait_stream(torch.cuda.current_stream())
        ops.flush_wgrad(on_side=True)
        with torch.cuda.stream(side):
            self._launch(b)

    def _on_grad(self, p) -> None:
        if p.grad is None or p.grad.data_ptr() != self._grad_ptr[p]:
            raise RuntimeError("GradReducer: param.grad no longer aliases the reduction bucket (was the gradient reset "
                               "with optimizer.zero_grad(set_to_none=True)? use reducer.zero_grad())")
        if not self._sync:
            return                      # accumulating micro-batch: the exchange happens after the last one
        b = self._bucket_of[p]
        if b.work is not None or b.pending <= 0:
            raise RuntimeError("GradReducer: a second backward pass reached a bucket that was already exchanged; run "
                               "all but the last micro-batch of an accumulated step under `with reducer.no_sync():` and "
                               "call finish() after the last one")
        b.pending -= 1
        if b.pending == 0 and self.world > 1:
            self._flush_and_launch(b)   # the queued weight-gradient unpacks of this bucket's convolutions, then NCCL

    def finish(self) -> None:
        """Call after the (last) backward: launches every bucket that is not in flight yet (parameters without a
        gradient this step, e.g. frozen branches), waits for all of them and re-arms the buckets."""
        if self.world == 1:
            return
        from . import ops
        ops.flush_wgrad()               # joins the wgrad side stream when anything is still queued
        for b in self.buckets:
            if b.work is None:
                self._launch(b)
        ops.join_wgrad_stream()         # (a CUDA-graph capture must see every forked stream joined again)
        for b in self.buckets:
            b.work.wait()
            if self.average and dist.get_backend(self.group) != "nccl":
                b.flat.div_(self.world)
            b.work = None
            b.pending = 0               # exchanged: a further backward before zero_grad() is an error (see _on_grad)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shard_rows(n_global: int, rank: int, world: int):
    """Rank-local row range [lo, hi) of a global batch (rank r gets rows [r*B, (r+1)*B))."""
    per = n_global // world
    return rank * per, (rank + 1) * per


class PeerAllReduce:
    """In-place sum of small fp32 vectors across the ranks of one node with ONE kernel over NVLink peer memory
    (csrc/msp_p2p.cu) instead of a NCCL call: SyncBN statistics and global Dice sums.

        par = PeerAllReduce(group)          # collective: every rank of `group`, once
        par.allreduce_sum_(stats)           # fp32, contiguous, numel <= max_floats; current stream; graph capturable

    Set-up allocates this rank's communication buffer with cudaMalloc, exchanges the cudaIpc handles with
    `all_gather_object` and maps every peer's buffer.  Every rank must issue the same sequence of calls."""

    def __init__(self, group=None, max_floats: int = 8192, device=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.max_floats = int(max_floats)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = int(_lib.lib.msp_p2p_buffer_bytes(self.world, self.max_floats))
        if nbytes <= 0:
            raise ValueError(f"PeerAllReduce: unsupported world size {self.world}")
        # Every collective below is reached by every rank whatever failed locally, so that the ranks fail TOGETHER
        # (a rank that raised early would leave the others blocked in the next collective).
        self._local, self._opened, err = None, [], None
        with torch.cuda.device(self.device):
            handle = None
            try:
                local = C.c_void_p()
                hbuf = C.create_string_buffer(64)
                _lib.call("msp_p2p_alloc", nbytes, C.byref(local), hbuf)
                self._local, handle = local.value, bytes(hbuf.raw)
            except Exception as e:   # noqa: BLE001
                err = e
            handles = [handle]
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, handle, group=group)
            self._ptrs = (C.c_void_p * self.world)()
            if err is None and all(h is not None for h in handles):
                try:
                    for r in range(self.world):
                        if r == self.rank:
                            self._ptrs[r] = self._local
                        else:
                            peer = C.c_void_p()
                            _lib.call("msp_p2p_open", C.create_string_buffer(handles[r], 64), C.byref(peer))
                            self._ptrs[r] = peer.value
                            self._opened.append(peer.value)
                except Exception as e:   # noqa: BLE001
                    err = e
            elif err is None:
                err = RuntimeError("a peer rank could not allocate its communication buffer")
            self.seq = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)   # msp_p2p_stats_exchange's block counter
            torch.cuda.synchronize()
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.device)
            if self.world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # also: every buffer is mapped everywhere
            if int(ok.item()) == 0:
                self.close()
                raise RuntimeError(f"PeerAllReduce: CUDA IPC set-up failed on at least one rank ({err})")

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() > self.max_floats or t.numel() == 0:
            raise ValueError("PeerAllReduce: needs a non-empty contiguous fp32 tensor of at most "
                             f"{self.max_floats} elements (got {t.dtype}, {t.numel()})")
        _lib.call("msp_p2p_allreduce_sum_f32", t.data_ptr(), t.numel(), self.rank, self.world, self.max_floats,
                  self._ptrs, self.seq.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        return t

    def stats_exchange(self, ws: torch.Tensor, rows: int, c: int, local_out=None, global_out=None, reset: bool = False,
                       finalize=None, local_halves=None, local_add: bool = False):
        """SyncBN exchange in one launch (msp_p2p_stats_exchange): `ws` = this rank's [rows, 2, C] per-CTA rows (or the
        [2, C] sums with rows = 1) -> rank-local sums (`local_out` [2, C], or `local_halves` = two [C] tensors that are
        written / with `local_add` added to: param.grad) -> sums over the ranks (`global_out`), optionally
        `finalize` = (global count, eps, momentum, mi [2, C], running_mean, running_var): mean / invstd / running stats."""
        if ws.dtype != torch.float32 or not ws.is_contiguous() or 2 * c > self.max_floats:
            raise ValueError(f"PeerAllReduce.stats_exchange: needs contiguous fp32 rows with 2C <= {self.max_floats}")
        count, eps, mom, mean, invstd, rm, rv = 0.0, 0.0, 0.0, None, None, None, None
        if finalize is not None:
            count, eps, mom, mi, rm, rv = finalize
            mean, invstd = mi[0].data_ptr(), mi[1].data_ptr()
        la = lb = None
        if local_out is not None:
            la, lb = local_out[0].data_ptr(), local_out[1].data_ptr()
        elif local_halves is not None:
            for t in local_halves:
                if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != c):
                    raise ValueError("PeerAllReduce.stats_exchange: local halves must be contiguous fp32 [C]")
            la = None if local_halves[0] is None else local_halves[0].data_ptr()
            lb = None if local_halves[1] is None else local_halves[1].data_ptr()
        _lib.call("msp_p2p_stats_exchange", ws.data_ptr(), int(rows), int(c), la, lb, int(local_add),
                  None if global_out is None else global_out.data_ptr(),
                  int(reset), float(count), float(eps), float(mom), mean, invstd,
                  None if rm is None else rm.data_ptr(), None if rv is None else rv.data_ptr(),
                  self.rank, self.world, self.max_floats, self._ptrs, self.seq.data_ptr(), self.ticket.data_ptr(),
                  torch.cuda.current_stream(self.device).cuda_stream)

    def close(self) -> None:
        torch.cuda.synchronize()
        for p in self._opened:
            try:
                _lib.call("msp_p2p_close", p)
            except Exception:   # noqa: BLE001 - best effort on the failure path
                pass
        self._opened = []
        if self._local:
            _lib.call("msp_p2p_free", self._local)
            self._local = None


_PEER = {}          # process group (None = default) -> PeerAllReduce


def enable_peer_allreduce(group=None, max_floats: int = 8192, strict: bool = False) -> Optional[PeerAllReduce]:
    """Route the small exchanges of `group` (SyncBN sums) through the peer-memory kernel.  Collective.

    The set-up needs CUDA IPC between the ranks' processes (one node, peer access).  Unless `strict`, a failure on ANY rank
    makes EVERY rank keep the NCCL path (the decision is all-reduced, so the ranks cannot disagree) and returns None."""
    key = group if group is not None else "default"
    if key in _PEER:
        return _PEER[key]
    try:
        par = PeerAllReduce(group, max_floats)       # raises on EVERY rank if the set-up failed on any
    except Exception as e:   # noqa: BLE001
        if strict:
            raise
        import sys
        print(f"[msp_b200] peer-memory all-reduce unavailable ({e}); SyncBN statistics stay on NCCL", file=sys.stderr,
              flush=True)
        return None
    _PEER[key] = par
    return par


def peer_allreduce_for(group) -> Optional[PeerAllReduce]:
    if not _PEER:
        return None
    par = _PEER.get(group if group is not None else "default")
    if par is None and group is not None and dist.is_initialized() and group is dist.group.WORLD:
        par = _PEER.get("default")
    return par


def allreduce_small_sum_(t: torch.Tensor, group) -> None:
    """Sum `t` over the ranks of `group` in place: peer-memory kernel when enabled and applicable, NCCL/gloo otherwise."""
    if group is None or not dist.is_initialized() or dist.get_world_size(group) <= 1:
        return
    par = peer_allreduce_for(group)
    if par is not None and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and 0 < t.numel() <= par.max_floats:
        par.allreduce_sum_(t)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
