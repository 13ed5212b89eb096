"""Host-side feeding of the hot path.

The reference moves every batch with a synchronous pageable `.to(device)` inside the step (train_model.py:60).
`BatchPrefetcher` keeps two sets of device buffers and a copy stream: the pinned host batch i+1 travels over PCIe while
the step on batch i runs; the consumer stream only waits on the copy's event."""
from __future__ import annotations

from typing import Sequence

import torch


class BatchPrefetcher:
    def __init__(self, example: Sequence[torch.Tensor], device):
        self.device = torch.device(device)
        self.bufs = [[torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in example] for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]      # recorded on the consumer stream once the buffer's step has been enqueued
        self._put = 0                 # next buffer to fill
        self._get = 0                 # next buffer to hand out
        self._last = None
        self._outstanding = 0         # batches copied (or being copied) but not handed out yet

    def put(self, *host: torch.Tensor) -> None:
        """Enqueue the H2D copy of the next batch (pinned memory for a truly asynchronous copy)."""
        k = self._put
        if self._outstanding >= 2:
            raise RuntimeError("BatchPrefetcher: both buffers hold batches that were not handed out yet (call get())")
        if self._last == k:
            # the buffer about to be refilled is the one the consumer was handed LAST: whatever has been enqueued on
            # the consumer stream so far may still read it, so the copy waits for that point (get() records the
            # event itself for the usual get-then-put order; put, put, get, step, put lands here)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.free[k] = ev
            self._last = None
        if self.free[k] is not None:
            self.copy_stream.wait_event(self.free[k])
        with torch.cuda.stream(self.copy_stream):
            for d, h in zip(self.bufs[k], host):
                d.copy_(h, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self._put ^= 1
        self._outstanding += 1

    def release(self) -> None:
        """The consumer has ENQUEUED its last read of the buffer handed out by the latest get() (e.g. graphs.GraphedStep
        has copied it into the graph's static inputs): the buffer may be refilled from that point of the current stream
        on.  Without this call the buffer counts as in use until the next get(), i.e. until the whole step that consumed
        it has finished — the refill then starts exactly at a step boundary and competes with the next step's own leading
        copies for the copy engine (measured: the full H2D time, 0.6 ms of a 10 ms U-Net step, landed on the critical
        path; bench.py --graph-timeline ... .e2e.txt)."""
        if self._last is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.free[self._last] = ev
            self._last = None

    def get(self):
        """Device tensors of the oldest pending batch; the current stream waits for its copy."""
        cur = torch.cuda.current_stream(self.device)
        if self._last is not None:   # everything enqueued so far has consumed the previous buffer
            ev = torch.cuda.Event()
            ev.record(cur)
            self.free[self._last] = ev
        if self._outstanding == 0:
            raise RuntimeError("BatchPrefetcher: get() without a pending put()")
        k = self._get
        cur.wait_event(self.ready[k])
        self._last = k
        self._get ^= 1
        self._outstanding -= 1
        return self.bufs[k]


class ScalarReader:
    """Device -> host read of a per-step scalar (the loss) that does not stall the pipeline.

    The reference reads the loss with a blocking `.item()` inside the step (loss/loss.py:85): the GPU then idles while
    the host prepares the next step.  Here the value is copied into pinned host memory on the consumer stream and
    handed out one step later (`depth` steps of slack), when its copy has long finished; `drain()` returns what is
    still pending.  Every step's value is still read back — only the wait is deferred."""

    def __init__(self, depth: int = 1, dtype=torch.float32):
        self.depth = depth
        self.slots = [torch.empty((), dtype=dtype).pin_memory() for _ in range(depth + 1)]
        self.events = [torch.cuda.Event() for _ in range(depth + 1)]
        self.pending = []          # slot indices in submission order
        self._next = 0

    def push(self, value: torch.Tensor):
        """Enqueue the D2H copy of `value` (0-dim device tensor); returns the oldest value once more than `depth`
        reads are pending, else None."""
        k = self._next
        self._next = (k + 1) % len(self.slots)
        self.slots[k].copy_(value.detach().reshape(()), non_blocking=True)
        self.events[k].record()
        self.pending.append(k)
        return self._pop() if len(self.pending) > self.depth else None

    def _pop(self) -> float:
        k = self.pending.pop(0)
        self.events[k].synchronize()
        return float(self.slots[k])

    def drain(self):
        return [self._pop() for _ in range(len(self.pending))]
