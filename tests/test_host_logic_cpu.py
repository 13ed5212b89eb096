"""Host-side logic that needs no GPU: the bench models' state-dict layout against the oracle's, the sequential
pretraining model and its checkpoint hand-off, the AdamW checkpoint loader, the derived-metric formulas and the metric
DAG driver against the reference's own classes (when the checkout is present) and the prefetcher's ordering guard."""
import itertools

import numpy as np
import pytest
import torch

from oracle import ref_models
from oracle import reference_harness as H
from medsegpretrainimagenet_b200 import metrics as M
from medsegpretrainimagenet_b200 import models, optim


@pytest.mark.parametrize("mine, theirs", [
    (lambda: models.resnet50_classifier(), lambda: ref_models.resnet50_classifier()),
    (lambda: models.resnet50_attention_unet(out_ch=4, final_activation="softmax"),
     lambda: ref_models.resnet50_attention_unet(out_ch=4, final_activation="softmax")),
    (lambda: models.resnet18_attention_unet(), lambda: ref_models.resnet18_attention_unet()),
    (lambda: models.basic_unet(out_ch=5), lambda: ref_models.basic_unet(out_ch=5)),
])
def test_bench_models_have_the_oracles_state_dict_layout(mine, theirs):
    """The objects bench.py times (`b200.models.*`) are the networks the oracle (pinned to the reference) defines: same
    keys in the same order, same shapes, same parameter count."""
    a, b = mine().state_dict(), theirs().state_dict()
    assert list(a.keys()) == list(b.keys())
    assert [tuple(v.shape) for v in a.values()] == [tuple(v.shape) for v in b.values()]


def test_sequential_pretraining_model_and_encoder_checkpoint_handoff(tmp_path):
    ffm = models.resnet50_pretraining_model()
    cls = models.resnet50_classifier()
    ka, kb = list(ffm.state_dict().keys()), list(cls.state_dict().keys())
    # [DeepResNet, pool, flatten, Linear] == DeepResNet(head=True) (classification/models.py:71-77) up to key names
    assert [k.replace("layers.0.", "").replace("layers.3.", "classifier.2.") for k in ka] == kb
    assert sum(p.numel() for p in ffm.parameters()) == 22_780_456
    path = tmp_path / "pretrained.pt"
    torch.save(ffm.state_dict(), path)
    unet = models.resnet50_attention_unet()
    before = unet.encoder.stem[0].weight.detach().clone()
    missing, unexpected = models.load_encoder_checkpoint(unet, str(path), strict=True)
    assert missing == [] and unexpected == []
    assert torch.equal(unet.encoder.stem[0].weight, ffm.layers[0].model.stem[0].weight)
    assert not torch.equal(unet.encoder.stem[0].weight, before)
    with pytest.raises(RuntimeError, match="CUDA"):
        ffm(torch.rand(1, 3, 32, 32))                          # no CPU fallback


def test_unsupported_heads_raise_unsupported_module():
    from medsegpretrainimagenet_b200 import converter
    enc = models.DeepResNet(bias=False)
    m = models.FeedForwardModel([enc, torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(2048, 10)])
    y = torch.empty((2, 7, 7, 2048))
    with pytest.raises(converter.UnsupportedModule, match="out_features"):
        converter.run_head(converter.ExecContext(), list(m.layers)[1:], y)


def test_adamw_loads_a_stock_torch_checkpoint(monkeypatch):
    """torch.optim.AdamW saves `step` as a CPU tensor; the kernel reads it on the parameter's device."""
    ps = [torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(3, 2))]
    ref = torch.optim.AdamW(ps, lr=1e-3, weight_decay=0.05)
    for n in range(3):
        for p in ps:
            p.grad = torch.randn_like(p)
        if n == 2:
            ps[1].grad = None                                   # parameters with different step counts
        ref.step()
    mine = optim.AdamW(ps, lr=1e-3, weight_decay=0.05)
    mine.load_state_dict(ref.state_dict())
    steps = [mine.state[p]["step"] for p in ps]
    assert all(s.dtype == torch.float32 and s.dim() == 0 and s.device == p.device for s, p in zip(steps, ps))
    assert [float(s) for s in steps] == [3.0, 2.0]
    assert [mine._host_steps[id(p)] for p in ps] == [3, 2]
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
    monkeypatch.setattr(optim, "_check", lambda t, what: None)
    for p in ps:
        p.grad = torch.randn_like(p)
    mine.step()
    assert [c[1][0] for c in calls] == [1, 1]                   # two launches: the parameters are one step apart
    assert [float(mine.state[p]["step"]) for p in ps] == [4.0, 3.0]


# ------------------------------------------------------------------------------------------------
# derived metrics and the metric DAG
# ------------------------------------------------------------------------------------------------
_COUNTS = [(5, 7, 2, 1), (0, 9, 0, 0), (0, 0, 0, 4), (3, 0, 0, 0), (0, 0, 5, 5), (10, 0, 3, 0), (1, 1, 1, 1)]


@pytest.mark.skipif(not H.available(), reason="/root/reference not present")
def test_derived_binary_metrics_match_the_reference_classes():
    H.setup()
    import metrics.metrics as ref
    for name in ("Accuracy", "BalancedAccuracy", "Sensitivity", "Specificity", "Precision", "DiceIndex", "JaccardIndex",
                 "MCC"):
        r, m = getattr(ref, name)(threshold=0.5), getattr(M, name)(threshold=0.5)
        assert r.name == m.name
        for tp, tn, fp, fn in _COUNTS:
            if name == "Accuracy" and tp + tn + fp + fn == 0:
                continue
            pv = {"true_positives": torch.tensor(tp), "true_negatives": torch.tensor(tn),
                  "false_positives": torch.tensor(fp), "false_negatives": torch.tensor(fn)}
            assert r.calculate_batch(pv) == m.calculate_batch(pv) == {}
            assert r.evaluate_batch(pv) == m.evaluate_batch(pv), (name, tp, tn, fp, fn)
            assert r.evaluate_epoch(pv) == m.evaluate_epoch(pv)
        assert r.evaluate_epoch(pv) == m.evaluate_epoch(pv)       # num_batches == 0 -> neutral value


@pytest.mark.skipif(not H.available(), reason="/root/reference not present")
def test_multiclass_means_match_the_reference_classes():
    H.setup()
    import metrics.multiclass_metrics as ref
    rng = np.random.default_rng(0)
    cfg = {"metrics/calculation/include_background_in_averages": False, "metrics/calculation/number_of_classes": 5,
           "metrics/calculation/log_classwise_dice_idcs": False, "metrics/calculation/log_classwise_jaccard_idcs": False}

    class CD(dict):
        def get(self, k, d=None):
            return dict.get(self, k, d)
    for inc in (False, True):
        cfg["metrics/calculation/include_background_in_averages"] = inc
        for rcls, mcls in ((ref.DiceIndex, M.MeanDiceIndex), (ref.JaccardIndex, M.MeanJaccardIndex)):
            r, m = rcls(_config_dict=CD(cfg)), mcls(_config_dict=cfg)
            for trial in range(4):
                cm = rng.integers(0, 50, (5, 5)).astype(float)
                if trial == 1:
                    cm[2, :] = 0; cm[:, 2] = 0                     # an absent class is left out of the mean
                if trial == 2:
                    cm[:] = 0
                pv = {"confusion_matrix": cm}
                a, b = r.evaluate_batch(pv), m.evaluate_batch(pv)
                assert a.keys() == b.keys() and all(abs(float(a[k]) - b[k]) < 1e-12 for k in a), (inc, trial)
    ra, ma = ref.Accuracy(), M.MultiClassAccuracy()
    for _ in range(3):
        cm = rng.integers(0, 9, (4, 4)).astype(float)
        assert ra.evaluate_batch({"confusion_matrix": cm}) == ma.evaluate_batch({"confusion_matrix": cm})
    assert ra.evaluate_epoch() == ma.evaluate_epoch()


def test_metric_dag_runs_parents_first_and_prefixes_values():
    """metrics/metric_wrapper.py:247-287 restated: one shared ConfusionMatrix parent per threshold, children see its
    counts with the `_threshold_x` suffix stripped, only int / float values are kept, the loss comes last."""
    calls = []

    class FakeCM:
        def __init__(self, threshold=0.5, **kw):
            self.threshold = threshold

        def _v(self, tag):
            calls.append((tag, self.threshold))
            t = self.threshold
            return {f"true_positives_threshold_{t}": torch.tensor(6), f"true_negatives_threshold_{t}": torch.tensor(3),
                    f"false_positives_threshold_{t}": torch.tensor(2), f"false_negatives_threshold_{t}": torch.tensor(1)}

        def calculate_batch(self, **kw):
            return self._v("calc")

        def evaluate_batch(self, **kw):
            return self._v("evalb")

        def evaluate_epoch(self, **kw):
            return self._v("evale")

    class Dice(M.DiceIndex):
        PARENT_METRIC = FakeCM

    class Mcc(M.MCC):
        PARENT_METRIC = FakeCM

    loss = lambda batch, *a, **k: {"dice_loss": 0.25}
    mc = M.MetricsCalculator([Dice, Mcc], loss=loss, thresholds=(0.5, 0.7))
    assert list(mc.metrics) == ["fake_c_m_threshold_0.5", "dice_index_threshold_0.5", "fake_c_m_threshold_0.7",
                                "dice_index_threshold_0.7", "mcc_threshold_0.5", "mcc_threshold_0.7"]
    out = mc.calculate_batch({"prediction": None, "mask": None}, train=True)
    assert out == {"dice_loss": 0.25}                           # tensors are not logged, derived metrics accumulate
    assert calls == [("calc", 0.5), ("calc", 0.7)]              # each parent evaluated once for all its children
    out = mc.evaluate_batch({})
    assert out["metrics/dice_index_threshold_0.5"] == (2 * 6 + 1) / (2 * 6 + 2 + 1 + 1)
    assert out["metrics/mcc_threshold_0.7"] == pytest.approx((6 * 3 - 2 * 1) / np.sqrt(7 * 8 * 5 * 4))
    assert out["dice_loss"] == 0.25 and set(out) == {"metrics/dice_index_threshold_0.5", "metrics/dice_index_threshold_0.7",
                                                     "metrics/mcc_threshold_0.5", "metrics/mcc_threshold_0.7", "dice_loss"}

    class Broken(M.MCC):
        PARENT_METRIC = FakeCM

        def evaluate_batch(self, parent_value, **kw):
            raise ValueError("kernel failure")
    with pytest.raises(ValueError, match="kernel failure"):     # not swallowed
        M.MetricsCalculator([Broken]).evaluate_batch({})


def test_one_vs_rest_counts_from_the_matrix():
    cm = np.arange(16).reshape(4, 4)
    for idx in range(4):
        tp, tn, fp, fn = M.binary_counts_of_class(cm, idx)
        assert tp == cm[idx, idx] and fn == cm[idx].sum() - tp and fp == cm[:, idx].sum() - tp
        assert tp + tn + fp + fn == cm.sum()


def test_symmetric_chunks_are_closed_under_the_negative_permutation():
    """Streaming cfg5's unpooled levels: every chunk, read as a local array, must pair its rows exactly as the global
    permutation [1, 0, n-1, ..., 2] (robustness/eval.py:22-23) does, and the chunks must cover every row."""
    from medsegpretrainimagenet_b200.robustness import negative_permutation, symmetric_chunks
    for n in (2, 3, 5, 10, 11, 50, 51, 1000, 1001):
        for rows in (6, 8, 16, 33):
            perm, seen = negative_permutation(n), set()
            for chunk in symmetric_chunks(n, rows):
                idx = [i for lo, hi in chunk for i in range(lo, hi)]
                local = negative_permutation(len(idx))
                assert all(idx[local[j]] == perm[i] for j, i in enumerate(idx)), (n, rows, chunk)
                assert len(idx) <= max(rows, 4) + 2
                seen |= set(idx)
            assert seen == set(range(n))


def test_gradient_sink_eligibility_and_branch_streams_off_the_gpu():
    """Host rules of two round-2 paths.  functional._grad_sink_ok: a kernel may ADD a parameter's gradient into `p.grad`
    itself only for a leaf with an existing fp32, contiguous, same-shape gradient (the reducer's bucket views, accumulated
    micro-batches) — otherwise the gradient goes back through autograd.  ExecContext.branch_streams: no side streams on the
    CPU, when switched off, or while bench.py's per-kernel instrumentation needs stream order."""
    import torch
    from medsegpretrainimagenet_b200 import converter, ops
    from medsegpretrainimagenet_b200.functional import _grad_sink_ok
    p = torch.nn.Parameter(torch.zeros(8))
    assert _grad_sink_ok(None)
    assert not _grad_sink_ok(p)                       # no gradient yet: autograd has to adopt one
    p.grad = torch.zeros(8)
    assert _grad_sink_ok(p)
    p.grad = torch.zeros(16)[::2]
    assert not _grad_sink_ok(p)                       # strided view
    q = torch.nn.Parameter(torch.zeros(8, dtype=torch.float64))
    q.grad = torch.zeros(8, dtype=torch.float64)
    assert not _grad_sink_ok(q)
    assert not _grad_sink_ok((p * 2.0))               # not a leaf
    ctx = converter.ExecContext()
    assert ctx.branch_streams("cpu", 1000) is None
    tl = []
    ops.set_conv_timeline(tl)
    try:
        assert ops.conv_timeline_active() and ops.wgrad_stream() is None
    finally:
        ops.set_conv_timeline(None)
    assert not ops.conv_timeline_active()
