"""ORACLE (test infrastructure): numpy / integer restatement of the reference's metric counters and the
scalar metrics derived from them."""
from __future__ import annotations

import math

import numpy as np
import torch


def confusion_counts(prediction, target, threshold=0.5, multilabel=False, ignore_nans=True):
    """metrics/metrics.py:61-95.  Returns int64 numpy arrays TP, TN, FP, FN, class_counts — scalars when
    `multilabel` is False (flatten from dim 0), one entry per channel otherwise (flatten from dim 1 after
    moving the channel axis first).  Positives: target == 1; predicted positives: prediction >= threshold;
    NaN targets count as negatives and are then subtracted from TN when `ignore_nans`."""
    p = prediction.detach().cpu().numpy()
    t = target.detach().cpu().numpy().reshape(p.shape)
    p, t = np.moveaxis(p, 0, 1), np.moveaxis(t, 0, 1)
    if multilabel:
        p, t = p.reshape(p.shape[0], -1), t.reshape(t.shape[0], -1)
    else:
        p, t = p.reshape(-1), t.reshape(-1)
    yp = t == 1
    hp = p >= threshold
    nan = np.isnan(t).sum(-1).astype(np.int64) * int(ignore_nans) if t.dtype.kind == "f" else 0
    tp = (yp & hp).sum(-1).astype(np.int64)
    tn = (~yp & ~hp).sum(-1).astype(np.int64) - nan
    fp = (~yp & hp).sum(-1).astype(np.int64)
    fn = (yp & ~hp).sum(-1).astype(np.int64)
    return tp, tn, fp, fn, yp.sum(-1).astype(np.int64)


def multiclass_confusion_matrix(prediction, target, num_classes):
    """metrics/multiclass_metrics.py:90-107: argmax over dim 1 (first maximal index, like torch), target
    arg-maxed too when it is one-hot shaped, then sklearn.metrics.confusion_matrix(y, y_hat,
    labels=range(C)) == a C x C histogram (rows = truth) ignoring labels outside the range."""
    y = target
    if y.shape == prediction.shape:
        y = y.argmax(dim=1)
    y = y.detach().cpu().flatten().numpy().astype(np.int64)
    y_hat = prediction.argmax(dim=1).detach().cpu().flatten().numpy().astype(np.int64)
    ok = (y >= 0) & (y < num_classes)
    cm = np.zeros((num_classes, num_classes), dtype=np.int64)
    np.add.at(cm, (y[ok], y_hat[ok]), 1)
    return cm


def topk_hits(prediction, target, k=5):
    """metrics/multiclass_metrics.py:424-446: number of positions whose label is among the k largest
    scores along dim 1.  Ties are resolved towards the lower class index (the documented behaviour of the
    kernel; torch.topk leaves tie order unspecified)."""
    y = target
    if y.shape == prediction.shape:
        y = y.argmax(dim=1, keepdim=True)
    p = prediction.detach().cpu().double()
    y = y.detach().cpu().long().reshape(p.shape[0], 1, *p.shape[2:])
    sl = torch.gather(p, 1, y)
    idx = torch.arange(p.shape[1]).reshape(1, -1, *([1] * (p.dim() - 2)))
    rank = ((p > sl) | ((p == sl) & (idx < y))).sum(dim=1)
    return int((rank < k).sum().item())


# ---- scalar metrics on integer counts (metrics/metrics.py:170-302) -------------------------------
def accuracy(tp, fp, tn, fn):
    return (tp + tn) / (tp + fp + tn + fn)


def balanced_accuracy(tp, tn, fp, fn):
    p, n = tp + fn, fp + tn
    try:
        if p == 0:
            return tn / n
        if n == 0:
            return tp / p
    except ZeroDivisionError:
        return "invalid"
    return (tp / p + tn / n) / 2


def tversky(tp, fp, fn, w_tp=1, w_fp=1, w_fn=1, eps=1):
    if tp + fp + fn == 0:
        return "invalid"   # neutral value 1 (metrics.py:252-258)
    return (w_tp * tp + eps) / (w_tp * tp + w_fp * fp + w_fn * fn + eps)


def dice_index(tp, fp, fn, eps=1):
    return tversky(tp, fp, fn, 2, 1, 1, eps)


def jaccard_index(tp, fp, fn, eps=1):
    return tversky(tp, fp, fn, 1, 1, 1, eps)


def mcc(tp, fp, fn, tn):
    denom_sq = (tp + fn) * (tp + fp) * (tn + fp) * (tn + fn)
    if denom_sq == 0:
        return "invalid"   # neutral value 0 (metrics.py:297-300)
    return (tp * tn - fp * fn) / math.sqrt(denom_sq)


def binary_from_multiclass(cm, idx):
    """metrics/multiclass_metrics.py:191-203."""
    tp = cm[idx, idx]
    fn = cm[idx, :].sum() - tp
    fp = cm[:, idx].sum() - tp
    return dict(tp=tp, fp=fp, fn=fn, tn=cm.sum() - tp - fn - fp)


def mean_over_present_classes(cm, fn_metric, include_background=False, neutral=1):
    """metrics/multiclass_metrics.py:205-217: mean of a binary metric over the classes that occur in
    the truth or the prediction, optionally skipping class 0."""
    vals = []
    for idx in range(0 if include_background else 1, cm.shape[0]):
        if cm[idx, :].sum() + cm[:, idx].sum() > 0:
            b = binary_from_multiclass(cm, idx)
            v = fn_metric(b["tp"], b["fp"], b["fn"])
            vals.append(neutral if v == "invalid" else v)
    return neutral if not vals else float(np.mean(vals))
