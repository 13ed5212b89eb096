"""Host logic of optim.SGD / AdamW / clip_grad_norm_ with the C-ABI calls recorded instead of executed (no GPU): 32
tensors per launch, momentum buffers created by a step are launched with first_step = 1 and apart from the older ones,
parameters without a gradient are skipped, AdamW groups parameters by step count and keeps torch's state keys, and the
classes refuse anything but contiguous fp32 CUDA tensors (no fallback)."""
import pytest
import torch

from medsegpretrainimagenet_b200 import optim


def _patched(monkeypatch, allow_cpu=True):
    calls = []

    def record(name, *a):
        if name == "msp_optim_add_scalar":      # the step-counter bump, emulated on the CPU tensors' memory
            import ctypes
            for i in range(a[0]):
                ctypes.c_float.from_address(a[1][i]).value += a[2]
            return
        calls.append((name, a))
    monkeypatch.setattr(optim._lib, "call", record)
    monkeypatch.setattr(optim, "_stream", lambda dev: 0)
    if allow_cpu:
        monkeypatch.setattr(optim, "_check", lambda t, what: None)
    return calls


def _params(n):
    ps = [torch.nn.Parameter(torch.randn(3 + i)) for i in range(n)]
    for p in ps:
        p.grad = torch.randn_like(p)
    return ps


def test_sgd_chunks_and_first_step_groups(monkeypatch):
    calls = _patched(monkeypatch)
    ps = _params(70)
    ps[5].grad = None
    opt = optim.SGD(ps, lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt.step()
    assert [c[0] for c in calls] == ["msp_optim_sgd"] * 3                     # 69 tensors -> 32 + 32 + 5
    assert [c[1][0] for c in calls] == [32, 32, 5]
    assert all(c[1][-2] == 1 for c in calls)                                  # first_step: buffers initialised to the gradient
    assert "momentum_buffer" in opt.state[ps[0]] and ps[5] not in opt.state or "momentum_buffer" not in opt.state[ps[5]]
    calls.clear()
    ps[5].grad = torch.randn_like(ps[5])                                      # joins one step late
    opt.param_groups[0]["lr"] = 0.05                                          # as a scheduler would
    opt.step()
    firsts = [(c[1][0], c[1][-2]) for c in calls]
    assert firsts == [(1, 1), (32, 0), (32, 0), (5, 0)]
    assert all(abs(c[1][5] - 0.05) < 1e-12 for c in calls)                    # lr read from param_groups at every step
    with pytest.raises(ValueError):
        optim.SGD(ps, lr=0.1, nesterov=True)                                  # torch's own argument check


def test_adamw_state_layout_and_step_groups(monkeypatch):
    calls = _patched(monkeypatch)
    ps = _params(40)
    ps[0].grad = None
    opt = optim.AdamW(ps, lr=0.004, betas=(0.9, 0.999), weight_decay=0.05)
    opt.step()
    assert [c[1][0] for c in calls] == [32, 7]
    assert set(opt.state[ps[1]].keys()) == {"step", "exp_avg", "exp_avg_sq"}     # torch.optim.AdamW's keys
    assert float(opt.state[ps[1]]["step"]) == 1.0
    calls.clear()
    ps[0].grad = torch.randn_like(ps[0])
    opt.step()
    assert sorted(c[1][0] for c in calls) == [1, 7, 32]                          # the late parameter has its own step count
    assert float(opt.state[ps[0]]["step"]) == 1.0 and float(opt.state[ps[1]]["step"]) == 2.0
    # one shared counter per parameter age, not one per parameter
    assert len({opt.state[p]["step"].data_ptr() for p in ps}) == 2
    # a parameter that sits a step out leaves the shared counter: the others move to a private copy
    calls.clear()
    ps[7].grad = None
    opt.step()
    assert float(opt.state[ps[7]]["step"]) == 2.0 and float(opt.state[ps[1]]["step"]) == 3.0
    assert float(opt.state[ps[0]]["step"]) == 2.0
    sd = opt.state_dict()
    assert len(sd["state"]) == 40 and sd["param_groups"][0]["betas"] == (0.9, 0.999)
    with pytest.raises(ValueError):
        optim.AdamW(ps, amsgrad=True)


def test_clip_grad_norm_calls_and_refusals(monkeypatch):
    calls = _patched(monkeypatch)
    ps = _params(33)
    ps[3].grad = None
    optim.clip_grad_norm_(ps, float("inf"))
    assert [c[0] for c in calls] == ["msp_optim_sqnorm", "msp_optim_norm"]       # 32 gradients: measure only
    assert calls[0][1][4] == 1                                                   # the first launch clears the accumulator
    calls.clear()
    optim.clip_grad_norm_(ps, 1.0)
    assert [c[0] for c in calls] == ["msp_optim_sqnorm", "msp_optim_norm", "msp_optim_clip"]
    with pytest.raises(RuntimeError):
        optim.clip_grad_norm_(ps, 1.0, norm_type=1.0)
    assert optim.clip_grad_norm_([torch.nn.Parameter(torch.zeros(3))], 1.0).item() == 0.0   # no gradients at all


def test_cpu_tensors_are_refused(monkeypatch):
    _patched(monkeypatch, allow_cpu=False)
    ps = _params(2)
    with pytest.raises(RuntimeError, match="no fallback"):
        optim.SGD(ps, lr=0.1).step()
    with pytest.raises(RuntimeError, match="no fallback"):
        optim.clip_grad_norm_(ps, 1.0)
