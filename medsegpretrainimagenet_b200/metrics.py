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
