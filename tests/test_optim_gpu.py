"""Multi-tensor optimizer kernels (csrc/msp_optim.cu, medsegpretrainimagenet_b200/optim.py) against torch.optim and
torch.nn.utils.clip_grad_norm_ — what the reference's step calls (train_model.py:93-107, optim/optimizer.py:41-48) — on
identical parameters and gradient sequences.  fp32 both sides: the only differences are fused multiply-adds and the order
of the norm's sum (tolerances beside each check)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _params(seed, n_extra=70):
    """Parameter tensors of many sizes, some of them odd-offset views of one flat buffer (GradReducer layout); more
    than 32 of them, so that several launches per step are needed."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 3, 7, 7), (64,), (1,), (7,), (2049,), (256, 64, 1, 1), (512, 512, 3, 3), (1000, 2048), (3, 5, 2)]
    shapes += [(17 + i,) for i in range(n_extra)]
    return [torch.randn(s, generator=g) for s in shapes]


def _clone_as_params(ts, flat_views=False):
    if not flat_views:
        return [torch.nn.Parameter(t.clone().to(DEV)) for t in ts]
    flat = torch.empty(sum(t.numel() for t in ts) + 3, device=DEV)
    out, off = [], 3                                   # odd element offset: nothing is 16-byte aligned
    for t in ts:
        v = flat[off:off + t.numel()].view(t.shape)
        v.copy_(t)
        out.append(torch.nn.Parameter(v))
        off += t.numel()
    return out


def _run(opt_a, pa, opt_b, pb, steps, seed):
    g = torch.Generator().manual_seed(seed)
    for step in range(steps):
        for a, c in zip(pa, pb):
            if step == 0 and a.numel() == 7:
                continue                                # one parameter gets its first gradient a step late
            gr = torch.randn(a.shape, generator=g).to(DEV)
            a.grad, c.grad = gr.clone(), gr.clone()
        opt_a.step()
        opt_b.step()


def _worst(pa, pb):
    return max(((a - c).abs().max() / (c.abs().max() + 1e-12)).item() for a, c in zip(pa, pb))


@pytest.mark.parametrize("nesterov,dampening,flat", [(False, 0.0, False), (True, 0.0, True), (False, 0.1, True)])
def test_sgd_matches_torch(nesterov, dampening, flat):
    from medsegpretrainimagenet_b200 import optim
    ts = _params(1)
    pa, pb = _clone_as_params(ts, flat), _clone_as_params(ts)
    kw = dict(lr=0.05, momentum=0.9, dampening=dampening, weight_decay=1e-4, nesterov=nesterov)
    _run(optim.SGD(pa, **kw), pa, torch.optim.SGD(pb, foreach=False, **kw), pb, 5, 2)
    assert _worst(pa, pb) <= 2e-6
    # plain SGD without momentum / weight decay
    pa, pb = _clone_as_params(ts), _clone_as_params(ts)
    _run(optim.SGD(pa, lr=0.1), pa, torch.optim.SGD(pb, lr=0.1, foreach=False), pb, 2, 3)
    assert _worst(pa, pb) <= 1e-6


def test_adamw_matches_torch_and_keeps_torch_state_layout():
    from medsegpretrainimagenet_b200 import optim
    ts = _params(4)
    pa, pb = _clone_as_params(ts, True), _clone_as_params(ts)
    kw = dict(lr=0.004, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)          # the cfg2 optimizer (simple.yaml:38-59)
    oa, ob = optim.AdamW(pa, **kw), torch.optim.AdamW(pb, foreach=False, **kw)
    _run(oa, pa, ob, pb, 6, 5)
    assert _worst(pa, pb) <= 5e-6
    sa, sb = oa.state[pa[0]], ob.state[pb[0]]
    assert set(sb.keys()) <= set(sa.keys())
    assert float(sa["step"]) == float(sb["step"]) == 6.0
    assert ((sa["exp_avg"] - sb["exp_avg"]).abs().max() / sb["exp_avg"].abs().max()).item() <= 5e-6
    assert ((sa["exp_avg_sq"] - sb["exp_avg_sq"]).abs().max() / sb["exp_avg_sq"].abs().max()).item() <= 5e-6
    # a learning-rate scheduler drives param_groups as with any torch optimizer
    sched = torch.optim.lr_scheduler.StepLR(oa, step_size=1, gamma=0.5)
    sched.step()
    assert oa.param_groups[0]["lr"] == pytest.approx(0.002)
    bad = torch.nn.Parameter(torch.zeros(4, dtype=torch.float64, device=DEV))
    bad.grad = torch.ones_like(bad)
    with pytest.raises(RuntimeError):                   # fp32 only, and no silent fallback
        optim.AdamW([bad]).step()


@pytest.mark.parametrize("max_norm", [float("inf"), 1.0, 1e6])
def test_clip_grad_norm_matches_torch(max_norm):
    from medsegpretrainimagenet_b200 import optim
    ts = _params(6)
    pa, pb = _clone_as_params(ts, True), _clone_as_params(ts)
    g = torch.Generator().manual_seed(7)
    for a, c in zip(pa, pb):
        gr = torch.randn(a.shape, generator=g).to(DEV)
        a.grad, c.grad = gr.clone(), gr.clone()
    pa[2].grad = None
    pb[2].grad = None                                   # parameters without a gradient are skipped
    na = optim.clip_grad_norm_(pa, max_norm)
    nb = torch.nn.utils.clip_grad_norm_(pb, max_norm, foreach=False)
    assert na.is_cuda and na.dim() == 0
    assert abs(na.item() - nb.item()) <= 1e-6 * nb.item()
    worst = max(((a.grad - c.grad).abs().max() / c.grad.abs().max()).item() for a, c in zip(pa, pb) if c.grad is not None)
    assert worst <= 1e-6
    with pytest.raises(RuntimeError):
        optim.clip_grad_norm_(pa, 1.0, norm_type=1.0)
