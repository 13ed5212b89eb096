"""Multi-tensor optimizer step and gradient-norm clipping on the B200 path.

The reference's step ends with `torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm, norm_type)`
(train_model.py:93-98) and `optimizer.step()` (train_model.py:107; `optim/optimizer.py:41-48` wraps any class named in the
YAML, e.g. `torch.optim.SGD`).  `SGD` and `AdamW` below take torch's constructor arguments, keep torch's `param_groups` /
`state` layout (`momentum_buffer`; `step`, `exp_avg`, `exp_avg_sq`) — so schedulers, `state_dict()` and the reference's
`Optimizer` wrapper work unchanged when the YAML names `medsegpretrainimagenet_b200.optim.SGD` — and update all parameters
with a handful of kernels of `csrc/msp_optim.cu` (32 tensors per launch).  fp32 CUDA parameters only; anything else raises.

`lr`, `momentum`, ... are read from `param_groups` at every `step()`; inside a CUDA-graph capture they are baked into the
graph (re-capture after a scheduler change, or keep the optimizer outside the graph)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch

from . import _lib

_CHUNK = 32


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr_array(tensors: List[torch.Tensor]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _numel_array(tensors: List[torch.Tensor]):
    arr = (C.c_longlong * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.numel()
    return arr


def _check(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError(f"medsegpretrainimagenet_b200.optim: {what} must be a contiguous fp32 CUDA tensor "
                           f"(got {t.dtype}, {t.device}, contiguous={t.is_contiguous()}); there is no fallback")


def _chunks(n: int):
    for lo in range(0, n, _CHUNK):
        yield lo, min(n, lo + _CHUNK)


@torch.no_grad()
def grad_sqnorm(grads: List[torch.Tensor]) -> torch.Tensor:
    """Device double holding sum(g^2) over all tensors (one launch per 32 tensors)."""
    dev = grads[0].device
    sq = torch.empty(1, dtype=torch.float64, device=dev)
    for lo, hi in _chunks(len(grads)):
        part = grads[lo:hi]
        _lib.call("msp_optim_sqnorm", len(part), _ptr_array(part), _numel_array(part), sq.data_ptr(), int(lo == 0),
                  _stream(dev))
    return sq


@torch.no_grad()
def clip_grad_norm_(parameters: Iterable[torch.Tensor], max_norm: float, norm_type: float = 2.0,
                    error_if_nonfinite: bool = False, foreach=None) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_ for the 2-norm: returns the total norm (0-dim device tensor, no host sync) and scales
    the gradients in place by min(1, max_norm / (norm + 1e-6)).  `max_norm = inf` only measures (train_model.py:93-98 logs
    the value as 'gradient_magnitude')."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if float(norm_type) != 2.0:
        raise RuntimeError("medsegpretrainimagenet_b200.optim.clip_grad_norm_: only norm_type 2 is implemented")
    if not grads:
        return torch.zeros(())
    for g in grads:
        _check(g, "every gradient")
    dev = grads[0].device
    sq = grad_sqnorm(grads)
    total = torch.empty((), dtype=torch.float32, device=dev)
    _lib.call("msp_optim_norm", sq.data_ptr(), total.data_ptr(), _stream(dev))
    if error_if_nonfinite and not bool(torch.isfinite(total)):
        raise RuntimeError("The total norm for gradients is non-finite, so it cannot be clipped")
    if max_norm != float("inf"):
        for lo, hi in _chunks(len(grads)):
            part = grads[lo:hi]
            _lib.call("msp_optim_clip", len(part), _ptr_array(part), _numel_array(part), sq.data_ptr(), float(max_norm),
                      _stream(dev))
    return total


class SGD(torch.optim.Optimizer):
    """torch.optim.SGD(params, lr, momentum, dampening, weight_decay, nesterov) on the multi-tensor kernel."""

    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, dampening: float = 0.0,
                 weight_decay: float = 0.0, nesterov: bool = False):
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("invalid SGD hyper-parameter")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            mom = float(group["momentum"])
            fresh, seasoned = [], []            # parameters whose momentum buffer is created by this step / exists
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p, "every parameter")
                _check(p.grad, "every gradient")
                st = self.state[p]
                if mom != 0 and "momentum_buffer" not in st:
                    st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    fresh.append(p)
                else:
                    seasoned.append(p)
            for plist, first in ((fresh, 1), (seasoned, 0)):
                for lo, hi in _chunks(len(plist)):
                    part = plist[lo:hi]
                    bufs = _ptr_array([self.state[p]["momentum_buffer"] for p in part]) if mom != 0 else None
                    _lib.call("msp_optim_sgd", len(part), _ptr_array(part), _ptr_array([p.grad for p in part]), bufs,
                              _numel_array(part), float(group["lr"]), mom, float(group["dampening"]),
                              float(group["weight_decay"]), int(bool(group["nesterov"])), first, _stream(part[0].device))
        return loss


class AdamW(torch.optim.Optimizer):
    """torch.optim.AdamW(params, lr, betas, eps, weight_decay) (no amsgrad / maximize) on the multi-tensor kernel."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 amsgrad: bool = False):
        if amsgrad:
            raise ValueError("medsegpretrainimagenet_b200.optim.AdamW: amsgrad is not implemented")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False))
        self._host_steps = {}

    def load_state_dict(self, state_dict):
        """torch.optim.AdamW checkpoints load as they are: torch leaves a non-capturable optimizer's `step` on the
        device it was saved on (the CPU for stock torch.optim.AdamW or a map_location='cpu' checkpoint), but the
        kernel reads the counter on the GPU — so every `step` is moved to an fp32 scalar on its parameter's device,
        and the host-side mirror of the step counts (one read per parameter, at load time only) is rebuilt."""
        super().load_state_dict(state_dict)
        self._host_steps = {}
        shared = {}
        for group in self.param_groups:
            if group.get("amsgrad", False):
                raise ValueError("medsegpretrainimagenet_b200.optim.AdamW: amsgrad state cannot be loaded")
            for p in group["params"]:
                st = self.state.get(p)
                if not st or "step" not in st:
                    continue
                step = st["step"]
                host = float(step.item()) if isinstance(step, torch.Tensor) else float(step)
                key = (p.device, host)
                if key not in shared:       # one counter per distinct age (see step())
                    shared[key] = torch.full((), host, dtype=torch.float32, device=p.device)
                st["step"] = shared[key]
                self._host_steps[id(p)] = int(round(host))
                for k in ("exp_avg", "exp_avg_sq"):
                    if k in st and (st[k].device != p.device or st[k].dtype != torch.float32 or not st[k].is_contiguous()):
                        st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        fresh_counters = {}
        for group in self.param_groups:
            plist = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p, "every parameter")
                _check(p.grad, "every gradient")
                st = self.state[p]
                if "exp_avg" not in st:
                    # parameters that start in the same step() share ONE step counter tensor (torch's state layout —
                    # every state[p]['step'] is a 0-dim fp32 tensor — with one increment per step instead of one per
                    # parameter)
                    key = (p.device, "fresh")
                    if key not in fresh_counters:
                        fresh_counters[key] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["step"] = fresh_counters[key]
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                plist.append(p)
            if not plist:
                continue
            self._split_partial_counters(plist)
            counters = list({self.state[p]["step"].data_ptr(): self.state[p]["step"] for p in plist}.values())
            for lo, hi in _chunks(len(counters)):
                _lib.call("msp_optim_add_scalar", hi - lo, _ptr_array(counters[lo:hi]), 1.0, _stream(plist[0].device))
            b1, b2 = group["betas"]
            for same_age in self._group_by_step(plist):
                for lo, hi in _chunks(len(same_age)):
                    part = same_age[lo:hi]
                    _lib.call("msp_optim_adamw", len(part), _ptr_array(part), _ptr_array([p.grad for p in part]),
                              _ptr_array([self.state[p]["exp_avg"] for p in part]),
                              _ptr_array([self.state[p]["exp_avg_sq"] for p in part]), _numel_array(part),
                              float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                              self.state[part[0]]["step"].data_ptr(), _stream(part[0].device))
        return loss

    def _split_partial_counters(self, plist) -> None:
        """A shared counter may only advance if ALL its parameters step: when some of them have no gradient this time,
        the ones that do step move to a private copy of the counter (rare: frozen / unused branches)."""
        stepping = {id(p) for p in plist}
        owners = {}
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if st and "step" in st:
                    owners.setdefault(st["step"].data_ptr(), []).append(p)
        for ptr, ps in owners.items():
            movers = [p for p in ps if id(p) in stepping]
            if movers and len(movers) != len(ps):
                fresh = self.state[movers[0]]["step"].clone()
                for p in movers:
                    self.state[p]["step"] = fresh

    def _group_by_step(self, plist):
        """Lists of parameters with the same step count (normally one list): a launch reads ONE device counter.  The
        count is tracked on the host as well, so no device read is needed to group."""
        groups = {}
        for p in plist:
            t = self._host_steps.get(id(p), 0) + 1
            self._host_steps[id(p)] = t
            groups.setdefault(t, []).append(p)
        return list(groups.values())
