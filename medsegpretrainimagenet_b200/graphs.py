"""CUDA-graph execution of a whole training / evaluation step.

The reference's step (train_model.py:51-120) is ~1000 small launches driven from Python; on a B200 the kernels of a
ResNet-50 step take ~25 ms while enqueueing them from the host takes about as long.  `GraphedStep` captures ONE call of a
user step function — forward, loss, backward, gradient reduction, metrics, clipping, optimizer — into a CUDA graph on
static buffers and replays it per batch: the host cost of a step becomes one graph launch.  (CUDA streams and graphs
instead of a tracing compiler: nothing is re-compiled, the captured work is exactly the eager kernels.)

Rules for the step function (the same as for any CUDA-graph capture): fixed shapes, no host synchronisation inside
(`.item()`, `.cpu()`; return device tensors and read them after the replay), optimizers created with `capturable=True`
where they keep a step counter (Adam/AdamW).  DropPath masks — drawn on the CPU generator by the reference
(classification/models.py:320-323) — are handled by the converter: inside a capture they live in static device buffers
that `GraphedStep` refills from the CPU generator, in the reference's order, before every replay."""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from . import _lib
from . import converter as _cv


class GraphedStep:
    def __init__(self, step_fn: Callable, example_inputs: Sequence[torch.Tensor], models: Sequence[torch.nn.Module] = (),
                 warmup: int = 3):
        """step_fn(*inputs) -> tensor or tuple of tensors; `models`: converted models whose DropPath masks must be
        refreshed per replay (may be empty)."""
        self.step_fn = step_fn
        self.after_copy = None          # optional callable run after the inputs were copied into the static buffers
        self.static_in = [torch.empty_like(t) for t in example_inputs]
        for s, t in zip(self.static_in, example_inputs):
            s.copy_(t)
        self.ctxs = [c for c in (_cv.context_of(m) for m in models) if c is not None]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):       # lazy initialisation (optimizer state, function attributes, ...)
                step_fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for c in self.ctxs:
            c.begin_static_droppath()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_out = step_fn(*self.static_in)
        self.launches_per_replay = _lib.launch_count() - n0   # kernels of libmsp_b200.so inside one replay
        for c in self.ctxs:
            c.end_static_droppath()

    def __call__(self, *inputs: torch.Tensor):
        for s, t in zip(self.static_in, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        if self.after_copy is not None:
            self.after_copy()           # e.g. host.BatchPrefetcher.release: the input buffers may be refilled from here on
        for c in self.ctxs:
            c.refresh_droppath()
        self.graph.replay()
        return self.static_out
