"""ORACLE (test infrastructure): fp32 restatement of the reference's robustness distances / scorer."""
from __future__ import annotations

import torch


def l2_distance(x, y):
    """robustness/distance.py:3-4."""
    return torch.mean((x - y) ** 2, dim=1)


def inv_pearson(x, y):
    """robustness/distance.py:6-7: 1 - corrcoef(x_i, y_i)[0, 1] row by row."""
    out = []
    for a, b in zip(x, y):
        out.append(torch.corrcoef(torch.stack([a.flatten(), b.flatten()]))[0, 1])
    return 1 - torch.stack(out)


def cosine_distance(x, y):
    """robustness/distance.py:9-10."""
    return 1 - torch.sum(x * y, dim=1) / torch.sqrt(torch.sum(x ** 2, dim=1) * torch.sum(y ** 2, dim=1))


def negative_permutation(n: int):
    """robustness/eval.py:22-23: reverse, then rotate by two -> [1, 0, n-1, n-2, ..., 2]."""
    rev = list(range(n - 1, -1, -1))
    return [rev[-2], rev[-1], *rev[:-2]]


def robustness_scores(preds0, preds1, distance_fn=cosine_distance, margin=0.5):
    """robustness/eval.py:16-28: max(0, d(q, k_pos) - d(q, k_neg) + margin)."""
    q, k1 = preds0.flatten(1), preds1.flatten(1)
    k0 = k1[negative_permutation(len(q))]
    return torch.clamp_min(distance_fn(q, k1) - distance_fn(q, k0) + margin, 0)


def pooled(features):
    """robustness/eval.py:51-52: spatial mean of a (N, C, H, W) representation."""
    return torch.mean(features.flatten(2), dim=2)
