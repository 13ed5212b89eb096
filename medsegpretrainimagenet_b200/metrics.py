"""Metric counters of the hot path, backed by the single-pass kernels of csrc/msp_metrics.cu, behind the
reference's Metric contract (`calculate_batch / evaluate_batch / evaluate_epoch` returning dicts;
metrics/metric_wrapper.py:247-322).

The classes can be named in a YAML (`medsegpretrainimagenet_b200.metrics.ConfusionMatrix`) and the
`*_calculate_batch` functions are what `patch.install()` binds onto the reference's own classes, whose
derived metrics (Dice / Jaccard / MCC / accuracy ..., metrics/metrics.py:126-302) keep consuming the
counts unchanged.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops


def _cfg(cd, path, default=None):
    """Look `a/b/c` up in a reference ConfigDict (slash paths) or a plain nested dict."""
    if cd is None:
        return default
    try:
        return cd[path]
    except Exception:
        pass
    cur = cd
    for part in path.split("/"):
        try:
            cur = cur[part]
        except Exception:
            return default
    return cur


def _f32c(t):
    t = t.detach().contiguous()
    return t if t.dtype == torch.float32 else t.float()


# ------------------------------------------------------------------------------------------------
# counting primitives
# ------------------------------------------------------------------------------------------------
def binary_confusion_counts(prediction, target, threshold=0.5, per_channel=False, nan_multiplicity=1):
    """metrics/metrics.py:61-82 in one pass.  -> dict of int64 tensors (0-dim, or (C,) when per_channel):
    TP, TN (NaN targets already subtracted), FP, FN, class_counts."""
    if not prediction.is_cuda:
        raise RuntimeError("binary_confusion_counts: CUDA tensors only")
    pred = _f32c(prediction)
    tgt = target.detach().reshape(pred.shape).contiguous()
    if tgt.dtype not in (torch.float32, torch.int64):
        tgt = tgt.float() if tgt.is_floating_point() else tgt.long()
    out = ops.confusion_binary(pred, tgt, threshold, per_channel)   # [..., 6]
    tp, tn, fp, fn, pos, nan = out.unbind(-1)
    return {"TP": tp, "TN": tn - nan * int(nan_multiplicity), "FP": fp, "FN": fn, "class_counts": pos}


def multiclass_confusion_matrix(prediction, target, num_classes=None):
    """metrics/multiclass_metrics.py:90-99 without the device -> host -> sklearn round trip.
    -> int64 CUDA tensor (C, C), rows = truth."""
    if not prediction.is_cuda:
        raise RuntimeError("multiclass_confusion_matrix: CUDA tensors only")
    pred = _f32c(prediction)
    c = pred.shape[1]
    if num_classes is not None and num_classes != c:
        raise ValueError(f"prediction has {c} classes, metric was configured for {num_classes}")
    onehot = tuple(target.shape) == tuple(pred.shape)
    tgt = _f32c(target) if onehot else target.detach().reshape(pred.shape[0], -1).long().contiguous()
    return ops.confusion_multiclass(pred, tgt, onehot)


def topk_correct(prediction, target, k=5):
    """metrics/multiclass_metrics.py:424-437 -> (int64 CUDA tensor [1] of hits, number of positions)."""
    pred = _f32c(prediction)
    if tuple(target.shape) == tuple(pred.shape):
        target = target.argmax(dim=1)
    lab = target.detach().reshape(pred.shape[0], -1).long().contiguous()
    return ops.topk_hits(pred, lab, k), lab.numel()


# ------------------------------------------------------------------------------------------------
# scalar metrics on integer counts (metrics/metrics.py:170-302) — host arithmetic on Python ints
# ------------------------------------------------------------------------------------------------
def accuracy(tp, fp, tn, fn):
    return (tp + tn) / (tp + fp + tn + fn)


def balanced_accuracy(tp, tn, fp, fn, neutral=0):
    p, n = tp + fn, fp + tn
    if p == 0 and n == 0:
        return neutral
    if p == 0:
        return tn / n
    if n == 0:
        return tp / p
    return (tp / p + tn / n) / 2


def tversky_index(tp, fp, fn, w_tp=1, w_fp=1, w_fn=1, eps=1, neutral=1):
    if tp + fp + fn == 0:
        return neutral
    return (w_tp * tp + eps) / (w_tp * tp + w_fp * fp + w_fn * fn + eps)


def dice_index(tp, fp, fn, eps=1):
    return tversky_index(tp, fp, fn, 2, 1, 1, eps)


def jaccard_index(tp, fp, fn, eps=1):
    return tversky_index(tp, fp, fn, 1, 1, 1, eps)


def mcc(tp, fp, fn, tn, neutral=0):
    denom_sq = (tp + fn) * (tp + fp) * (tn + fp) * (tn + fn)
    if denom_sq == 0:
        return neutral
    return (tp * tn - fp * fn) / math.sqrt(denom_sq)


# ------------------------------------------------------------------------------------------------
# Metric-contract classes
# ------------------------------------------------------------------------------------------------
def confusion_calculate_batch(self, prediction, mask=None, label=None, cumulate=True, *args, **kwargs):
    """Drop-in body for ConfusionMatrix.calculate_batch (metrics/metrics.py:61-95)."""
    y = mask if mask is not None else label
    c = binary_confusion_counts(prediction, y, self.threshold, per_channel=bool(self.idx_start),
                                nan_multiplicity=self.nan_multiplicity)
    tp, tn, fp, fn = c["TP"], c["TN"], c["FP"], c["FN"]
    self.class_counts = self.class_counts + c["class_counts"]
    if cumulate:
        self.TP, self.TN, self.FP, self.FN = self.TP + tp, self.TN + tn, self.FP + fp, self.FN + fn
    if self.accumulate:
        self.acc_TP, self.acc_TN = self.acc_TP + tp, self.acc_TN + tn
        self.acc_FP, self.acc_FN = self.acc_FP + fp, self.acc_FN + fn
    t = self.threshold
    return {f"true_positives_threshold_{t}": tp, f"false_positives_threshold_{t}": fp,
            f"true_negatives_threshold_{t}": tn, f"false_negatives_threshold_{t}": fn}


class ConfusionMatrix:
    """metrics/metrics.py:29-124."""

    PARAMS = dict(multilabel=False, ignore_nans=True)

    def __init__(self, _config_dict=None, threshold=0.5, accumulate=True, *args, **kwargs):
        self.threshold = threshold
        self.multilabel = bool(_cfg(_config_dict, "metrics/calculation/multilabel", False))
        self.idx_start = int(self.multilabel)
        self.nan_multiplicity = int(_cfg(_config_dict, "metrics/calculation/ignore_nans", True))
        self.accumulate = accumulate
        self.TP = self.TN = self.FP = self.FN = 0
        self.acc_TP = self.acc_TN = self.acc_FP = self.acc_FN = 0
        self.class_counts = 0

    calculate_batch = confusion_calculate_batch

    def _pack(self, tp, tn, fp, fn):
        t = self.threshold
        return {f"true_positives_threshold_{t}": tp, f"false_positives_threshold_{t}": fp,
                f"true_negatives_threshold_{t}": tn, f"false_negatives_threshold_{t}": fn}

    def evaluate_batch(self, flush=True, *args, **kwargs):
        out = self._pack(self.acc_TP, self.acc_TN, self.acc_FP, self.acc_FN)
        if flush:
            self.acc_TP = self.acc_TN = self.acc_FP = self.acc_FN = 0
        return out

    def evaluate_epoch(self, flush=True, *args, **kwargs):
        out = self._pack(self.TP, self.TN, self.FP, self.FN)
        out[f"class_counts_threshold_{self.threshold}"] = self.class_counts
        if flush:
            self.TP = self.TN = self.FP = self.FN = 0
            self.class_counts = 0
        return out


def multiclass_calculate_batch(self, prediction, mask=None, label=None, cumulate=True, *args, **kwargs):
    """Drop-in body for MultiClassConfusionMatrix.calculate_batch (multiclass_metrics.py:90-107): the
    C x C histogram is counted on the device; only C*C int64 values cross to the host, where the
    reference keeps its float64 accumulators."""
    y = mask if mask is not None else label
    cm_dev = multiclass_confusion_matrix(prediction, y, len(self.range))
    cm = cm_dev.cpu().numpy()
    row = cm.sum(axis=1)
    self.class_counts = [c + int(r) for c, r in zip(self.class_counts, row)]
    if cumulate:
        self.cm += cm
    if self.accumulate:
        self.acc_cm += cm
    return {"confusion_matrix": cm}


class MultiClassConfusionMatrix:
    """metrics/multiclass_metrics.py:11-124 (without the plotting side)."""

    PARAMS = {"number_of_classes": 1000, "log_confusion_matrix": False}

    def __init__(self, accumulate=True, _config_dict=None, number_of_classes=None, *args, **kwargs):
        n = number_of_classes or _cfg(_config_dict, "metrics/calculation/number_of_classes", 1000)
        self.init_cm = lambda: np.zeros((n, n))
        self.cm = self.init_cm()
        self.range = list(range(n))
        self.accumulate = accumulate
        if accumulate:
            self.acc_cm = self.init_cm()
        self.class_counts = [0] * n

    calculate_batch = multiclass_calculate_batch

    def evaluate_batch(self, flush=True, *args, **kwargs):
        cm = self.acc_cm
        if flush:
            self.acc_cm = self.init_cm()
        return {"confusion_matrix": cm}

    def evaluate_epoch(self, flush=True, *args, **kwargs):
        cm, counts = self.cm, self.class_counts
        if flush:
            self.cm = self.init_cm()
            self.class_counts = [0 for _ in counts]
        return {"confusion_matrix": cm, "class_counts": counts}


def top5_calculate_batch(self, prediction, mask=None, label=None, cumulate=True, *args, **kwargs):
    """Drop-in body for Top5Accuracy.calculate_batch (multiclass_metrics.py:424-446)."""
    y = mask if mask is not None else label
    hits, num = topk_correct(prediction, y, self.n)
    hits = int(hits.item())
    if cumulate:
        self.num_correct_preds += hits
        self.num_records += num
    if self.accumulate:
        self.num_correct_preds_in_batch += hits
        self.num_records_in_batch += num
    return {self.name: hits / num}


class Top5Accuracy:
    """metrics/multiclass_metrics.py:410-458."""

    def __init__(self, accumulate=True, *args, **kwargs):
        self.name, self.n, self.accumulate = "top_5_accuracy", 5, accumulate
        self.num_records = self.num_correct_preds = 0
        self.num_records_in_batch = self.num_correct_preds_in_batch = 0

    calculate_batch = top5_calculate_batch

    def evaluate_batch(self, flush=True, *args, **kwargs):
        n, k = self.num_records_in_batch, self.num_correct_preds_in_batch
        if flush:
            self.num_records_in_batch = self.num_correct_preds_in_batch = 0
        return {self.name: k / n}

    def evaluate_epoch(self, flush=True, *args, **kwargs):
        n, k = self.num_records, self.num_correct_preds
        if flush:
            self.num_records = self.num_correct_preds = 0
        return {self.name: k / n}


# ------------------------------------------------------------------------------------------------
# derived metrics (Metric contract with PARENT_METRIC) and the calculator that drives the metric DAG
# ------------------------------------------------------------------------------------------------
def _host_counts(parent_value):
    """Four count tensors -> Python ints with ONE device -> host copy (the reference calls `.item()` on each of them in
    every derived metric, metrics/metrics.py:152,166: 4 host syncs per metric per step)."""
    keys = list(parent_value.keys())
    vals = [parent_value[k] for k in keys]
    if all(isinstance(v, torch.Tensor) and v.is_cuda and v.dim() == 0 for v in vals):
        host = torch.stack(vals).cpu().tolist()
        return dict(zip(keys, (int(v) for v in host)))
    return {k: (v.item() if hasattr(v, "item") else v) for k, v in parent_value.items()}


class DerivedConfusionMatrixMetric:
    """metrics/metrics.py:126-167: a scalar formula on the parent ConfusionMatrix's counts.  `calculate_batch` is silent
    while accumulating, `evaluate_batch` / `evaluate_epoch` evaluate the formula on the accumulated counts; the string
    'invalid' from the formula maps to the metric's neutral value."""

    PARENT_METRIC = ConfusionMatrix
    FORMULA = None          # staticmethod(tp, tn, fp, fn) -> float or 'invalid'
    NAME = None
    NEUTRAL = 0

    def __init__(self, accumulate=True, threshold=0.5, _config_dict=None, *args, **kwargs):
        self.name = f"{self.NAME}_threshold_{threshold}"
        self.neutral = self.NEUTRAL
        self.accumulate = accumulate
        self.num_batches = 0

    def _value(self, parent_value):
        c = _host_counts(parent_value)
        v = type(self).FORMULA(c["true_positives"], c["true_negatives"], c["false_positives"], c["false_negatives"])
        return self.neutral if isinstance(v, str) else v

    def calculate_batch(self, parent_value, calculate=False, *args, **kwargs):
        if self.accumulate and not calculate:
            return {}
        self.num_batches += 1
        return {self.name: self._value(parent_value)}

    def evaluate_batch(self, parent_value, *args, **kwargs):
        return self.calculate_batch(parent_value, calculate=True)

    def evaluate_epoch(self, parent_value, flush=True, *args, **kwargs):
        if self.num_batches == 0:
            return {self.name: self.neutral}
        if flush:
            self.num_batches = 0
        return {self.name: self._value(parent_value)}


def _derived(name, formula, neutral=0, doc=""):
    return type(name, (DerivedConfusionMatrixMetric,),
                {"NAME": "".join("_" + ch.lower() if ch.isupper() and i else ch.lower() for i, ch in enumerate(name))
                 if name != "MCC" else "mcc", "FORMULA": staticmethod(formula), "NEUTRAL": neutral, "__doc__": doc})


def _safe(fn):
    def g(tp, tn, fp, fn_):
        try:
            return fn(tp, tn, fp, fn_)
        except ZeroDivisionError:
            return "invalid"
    return g


def _bal_acc(tp, tn, fp, fn):
    p, n = tp + fn, fp + tn
    if p == 0:
        return tn / n
    if n == 0:
        return tp / p
    return (tp / p + tn / n) / 2


def _tversky(w_tp, w_fp, w_fn, eps=1):
    def f(tp, tn, fp, fn):
        if tp + fp + fn == 0:
            return "invalid"
        return (w_tp * tp + eps) / (w_tp * tp + w_fp * fp + w_fn * fn + eps)
    return f


def _mcc(tp, tn, fp, fn):
    d = (tp + fn) * (tp + fp) * (tn + fp) * (tn + fn)
    return "invalid" if d == 0 else (tp * tn - fp * fn) / math.sqrt(d)


Accuracy = _derived("Accuracy", lambda tp, tn, fp, fn: (tp + tn) / (tp + fp + tn + fn), doc="metrics/metrics.py:170-179")
BalancedAccuracy = _derived("BalancedAccuracy", _safe(_bal_acc), doc="metrics/metrics.py:181-199")
Sensitivity = _derived("Sensitivity", lambda tp, tn, fp, fn: "invalid" if tp + fn == 0 else tp / (tp + fn),
                       doc="metrics/metrics.py:201-211")
Specificity = _derived("Specificity", lambda tp, tn, fp, fn: "invalid" if tn + fp == 0 else tn / (tn + fp),
                       doc="metrics/metrics.py:213-224")
Precision = _derived("Precision", _safe(lambda tp, tn, fp, fn: tp / (tp + fp)), doc="metrics/metrics.py:226-234")
DiceIndex = _derived("DiceIndex", _tversky(2, 1, 1), neutral=1, doc="metrics/metrics.py:262-272 (eps = 1)")
JaccardIndex = _derived("JaccardIndex", _tversky(1, 1, 1), neutral=1, doc="metrics/metrics.py:274-284 (eps = 1)")
MCC = _derived("MCC", _mcc, doc="metrics/metrics.py:286-302")


def binary_counts_of_class(cm: np.ndarray, idx: int):
    """metrics/multiclass_metrics.py:191-203: one-vs-rest counts of class `idx` from the C x C matrix (rows = truth)."""
    tp = cm[idx, idx]
    fn = cm[idx, :].sum() - tp
    fp = cm[:, idx].sum() - tp
    return tp, cm.sum() - tp - fn - fp, fp, fn


class AverageBinaryCMMetric:
    """metrics/multiclass_metrics.py:156-246 without the class-wise logging side: the mean over the classes PRESENT in
    the matrix (row or column sum > 0) of a binary formula on their one-vs-rest counts."""

    PARENT_METRIC = MultiClassConfusionMatrix
    NAME, BINARY = None, None

    def __init__(self, _config_dict=None, include_background_in_averages=None, number_of_classes=None, *args, **kwargs):
        inc = include_background_in_averages
        if inc is None:
            inc = _cfg(_config_dict, "metrics/calculation/include_background_in_averages", False)
        self.start = int(not inc)
        self.num_classes = number_of_classes or _cfg(_config_dict, "metrics/calculation/number_of_classes", 1000)
        self.name = self.NAME
        self.neutral = self.BINARY.NEUTRAL

    def _mean(self, parent_value):
        cm = np.asarray(parent_value["confusion_matrix"])
        vals = []
        for idx in range(self.start, self.num_classes):
            if cm[idx, :].sum() + cm[:, idx].sum() > 0:
                v = self.BINARY.FORMULA(*binary_counts_of_class(cm, idx))
                vals.append(self.neutral if isinstance(v, str) else v)
        return {self.name: self.neutral if not vals else float(np.mean(vals))}

    def calculate_batch(self, parent_value, *args, **kwargs):
        return {}           # binary metrics accumulate (metrics/metrics.py:145-146): nothing per fragment

    def evaluate_batch(self, parent_value, *args, **kwargs):
        return self._mean(parent_value)

    def evaluate_epoch(self, parent_value, *args, **kwargs):
        return self._mean(parent_value)


class MeanDiceIndex(AverageBinaryCMMetric):
    """metrics/multiclass_metrics.py:248-262 (`mean_dice_index`)."""
    NAME, BINARY = "mean_dice_index", DiceIndex


class MeanJaccardIndex(AverageBinaryCMMetric):
    """metrics/multiclass_metrics.py:264-278 (`mean_jaccard_index`)."""
    NAME, BINARY = "mean_jaccard_index", JaccardIndex


class MultiClassAccuracy:
    """metrics/multiclass_metrics.py:292-316: trace / total per batch, mean of the batch values per epoch."""
    PARENT_METRIC = MultiClassConfusionMatrix

    def __init__(self, accumulate=True, *args, **kwargs):
        self.name, self.accumulate, self.num_batches, self.value = "accuracy", accumulate, 0, 0

    def calculate_batch(self, *args, **kwargs):
        return None

    def evaluate_batch(self, parent_value, *args, **kwargs):
        cm = np.asarray(parent_value["confusion_matrix"])
        value = float(np.diagonal(cm).sum() / np.sum(cm))
        self.value += value
        self.num_batches += 1
        return {self.name: value}

    def evaluate_epoch(self, flush=True, *args, **kwargs):
        value = self.value / self.num_batches
        if flush:
            self.value, self.num_batches = 0, 0
        return {self.name: value}


class MetricsCalculator:
    """Host-side mirror of metrics/metric_wrapper.py:122-322 for use WITHOUT the reference checkout: builds the metric
    DAG (every metric's PARENT_METRIC is created once per threshold and evaluated first, its dict handed to the child as
    `parent_value` with the `_threshold_x` suffix stripped from the keys, :254-262), calls `calculate_batch` /
    `evaluate_batch` / `evaluate_epoch` on all of them and then on the loss, keeps only int / float values and
    prefixes them with 'metrics/' (:281).  Unlike the reference (:268-272) a failing metric RAISES: on this path a
    swallowed exception would hide a kernel failure.

        mc = MetricsCalculator([metrics.DiceIndex, metrics.MCC], loss=loss_wrapper, thresholds=(0.5,), config=cfg)
        logs = mc.calculate_batch(batch, train=True);  logs = mc.evaluate_batch(batch);  logs = mc.evaluate_epoch()
    """

    def __init__(self, metric_classes, loss=None, thresholds=(0.5,), config=None, **kwargs):
        import inspect
        self.metrics = {}
        for cls in metric_classes:
            for thr in (thresholds if self._needs_threshold(cls, inspect) else (None,)):
                self._create(cls, thr, config, kwargs, inspect)
        self.loss = loss if loss is not None else (lambda *a, **k: {})
        self.loss_name = getattr(loss, "name", "loss")

    @staticmethod
    def _needs_threshold(cls, inspect):
        if "threshold" in inspect.signature(cls).parameters:
            return True
        parent = getattr(cls, "PARENT_METRIC", None)
        return parent is not None and MetricsCalculator._needs_threshold(parent, inspect)

    def _create(self, cls, thr, config, kwargs, inspect):
        kw = dict(kwargs, _config_dict=config)
        if self._needs_threshold(cls, inspect):
            kw["threshold"] = thr
        obj = cls(**kw)
        name = getattr(obj, "name", None) or "".join("_" + c.lower() if c.isupper() and i else c.lower()
                                                      for i, c in enumerate(cls.__name__))
        if "threshold" not in name and self._needs_threshold(cls, inspect):
            name = f"{name}_threshold_{thr}"
        entry = {"calculator": obj}
        parent = getattr(cls, "PARENT_METRIC", None)
        if parent is not None:
            entry["parent"] = self._create(parent, thr, config, kwargs, inspect)
        if name not in self.metrics:           # one shared parent per (class, threshold)
            self.metrics[name] = entry
        return name

    def _run(self, batch, func, *args, **kwargs):
        import re
        done = {}

        def calc(name):
            if name in done:
                return done[name]
            entry = self.metrics[name]
            parent = entry.get("parent")
            f = getattr(entry["calculator"], func)
            if parent:
                pv = calc(parent)
                if "threshold" in name and pv is not None:
                    pv = {re.match("(.*)_threshold_.*", k).group(1): v for k, v in pv.items()}
                value = f(parent_value=pv, *args, **kwargs, **batch)
            else:
                value = f(*args, **kwargs, **batch)
            done[name] = value
            return value

        for name in list(self.metrics):
            calc(name)
        values = {}
        for value in done.values():
            if value is not None:
                values.update(value)
        values = {"metrics/" + k: v for k, v in values.items() if isinstance(v, (int, float))}
        values.update(getattr(self.loss, func, self.loss)(batch, *args, **kwargs) or {})
        return values

    def calculate_batch(self, batch, *args, **kwargs):
        return self._run(batch, "calculate_batch", *args, **kwargs)

    def evaluate_batch(self, batch, *args, **kwargs):
        return self._run(batch, "evaluate_batch", *args, **kwargs)

    def evaluate_epoch(self, *args, **kwargs):
        return self._run({}, "evaluate_epoch", *args, **kwargs)
