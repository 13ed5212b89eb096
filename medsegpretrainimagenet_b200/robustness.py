"""Encoder-robustness distances and triplet scorer (robustness/distance.py:3-10,
robustness/eval.py:7-54) on the single-pass row-pair kernel of csrc/msp_robust.cu.

The reference evaluates each distance separately, once for the positives and once for the explicitly
materialised negatives `k0 = k1[perm]`; here ONE launch reads every row of q and k once and yields all
three distances for both pairings, so `Robustness` accepts the distance by name (or by the functions
below, which carry a `kind` tag) and never builds `k0`.
"""
from __future__ import annotations

from typing import Callable, Sequence, Union

import torch

from . import ops

_KINDS = {"cosine": 0, "l2": 1, "pearson": 2}


def negative_permutation(n: int):
    """robustness/eval.py:22-23 -> [1, 0, n-1, n-2, ..., 2] (an involution)."""
    return [1, 0] + list(range(n - 1, 1, -1)) if n >= 2 else list(range(n))


def _f32c(t):
    t = t.detach().contiguous()
    return t if t.dtype == torch.float32 else t.float()


def all_distances(preds0: torch.Tensor, preds1: torch.Tensor, pool: bool = False) -> torch.Tensor:
    """-> fp32 (6, N): cos(q,k+), cos(q,k-), l2(q,k+), l2(q,k-), 1-pearson(q,k+), 1-pearson(q,k-), with
    k- the fixed negative permutation.  `pool=True`: inputs are (N, C, H, W) maps and the spatial mean
    (robustness/eval.py:51-52) is fused into the same pass."""
    if not preds0.is_cuda:
        raise RuntimeError("robustness distances: CUDA tensors only")
    q, k = _f32c(preds0), _f32c(preds1)
    if q.shape != k.shape:
        raise ValueError("representations must have the same shape")
    if pool:
        if q.dim() < 3:
            raise ValueError("pool=True needs (N, C, *spatial) representations")
        hw = q[0, 0].numel()
        return ops.rowpair_distances(q.reshape(q.shape[0], q.shape[1], hw), k.reshape(q.shape[0], q.shape[1], hw),
                                     pooled_hw=hw if hw > 1 else 0)
    return ops.rowpair_distances(q.flatten(1), k.flatten(1))


def _pairwise(kind):
    def fn(x, y):
        """Distance of row i of x to row i of y (the reference's signature, distance.py:3-10)."""
        # positives only: evaluate with an identity pairing by running the kernel on (x, y)
        return all_distances(x, y)[2 * _KINDS[kind]]
    fn.kind = kind
    fn.__name__ = {"cosine": "cosine_distance", "l2": "l2_loss", "pearson": "inv_pearson_corr"}[kind]
    return fn


cosine_distance = _pairwise("cosine")
l2_loss = _pairwise("l2")
inv_pearson_corr = _pairwise("pearson")


class Robustness:
    """robustness/eval.py:7-28: max(0, d(q, k+) - d(q, k-) + margin) per row."""

    def __init__(self, distance_fn: Union[str, Callable] = cosine_distance, margin: float = 0.5):
        kind = distance_fn if isinstance(distance_fn, str) else getattr(distance_fn, "kind", None)
        if kind is None:
            name = getattr(distance_fn, "__name__", "")
            kind = {"cosine_distance": "cosine", "l2_loss": "l2", "inv_pearson_corr": "pearson"}.get(name)
        if kind not in _KINDS:
            raise ValueError(f"unknown distance {distance_fn!r}")
        self.kind, self.margin = kind, margin

    def __call__(self, preds0: torch.Tensor, preds1: torch.Tensor, pool: bool = False) -> torch.Tensor:
        d = all_distances(preds0, preds1, pool=pool)
        m = torch.tensor([self.margin], dtype=torch.float32, device=d.device)
        return ops.triplet_hinge(d, m)[0, _KINDS[self.kind]]


def robustness_table(preds0, preds1, margins: Sequence[float] = (0.0, 0.25, 0.5, 0.75, 1.0), pool=False):
    """All three distances x all margins from one pass over the data -> fp32 (len(margins), 3, N) in the
    order (cosine, l2, pearson) — the (level, pooled) slice of results/robustness_scores.csv."""
    d = all_distances(preds0, preds1, pool=pool)
    m = torch.tensor(list(margins), dtype=torch.float32, device=d.device)
    return ops.triplet_hinge(d, m)


@torch.no_grad()
def predict_w_model(model, imgs: torch.Tensor, batch_size: int = 32, device="cuda:0", level: int = -2,
                    pool: bool = True, *args, **kwargs) -> torch.Tensor:
    """robustness/eval.py:30-54 with its missing `torch.cat` restored: batched encoder
    forward with `return_skip_vals=True`, the requested level's representation for every image.  Pooling
    is left to the distance kernel (pass `pool=` to Robustness / robustness_table) unless requested here,
    in which case the reference's (N, C) tensor is returned."""
    model = model.to(device)
    # `model.layers[0]` is a `Model` wrapper whose forward drops every argument but x (model/model.py:65-70), so the
    # reference's `model(x, return_skip_vals=True)` never reaches the encoder through it: call the wrapped network
    from .converter import _unwrap
    net = _unwrap(model)
    outs = []
    for i in range(0, len(imgs), batch_size):
        y, inner = net(imgs[i:i + batch_size].to(device), return_skip_vals=True)
        levels = list(inner) + [y]
        outs.append(levels[level])
    pred = torch.cat(outs)
    if pool:
        return torch.mean(pred.flatten(2), dim=2)
    return pred


def eval_encoder(model, imgs: torch.Tensor, scorer: Robustness, level: int, pool: bool, *args, **kwargs):
    """robustness/eval.py:56-69: the encoder is the first layer of the sequential pretraining model
    (`model.model.layers[0]`), two independent ColorJitter(0.1, 0.05, 0.1, 0.05) views of the image batch go through
    it, the requested level's representations are scored.  The augmentation runs on the device
    (transforms.ColorJitter, torchvision's parameter draw on the CPU generator) and the spatial mean of `pool=True`
    is fused into the distance kernel instead of being materialised."""
    from .transforms import ColorJitter
    inner = getattr(model, "model", model)
    encoder = inner.layers[0]
    encoder.eval()
    device = kwargs.pop("device", "cuda:0")
    aug = ColorJitter(brightness=0.1, contrast=0.05, hue=0.05, saturation=0.1)
    imgs = imgs.to(device)
    imgs0, imgs1 = aug(imgs), aug(imgs)
    preds0 = predict_w_model(encoder, imgs0, level=level, pool=False, device=device, *args, **kwargs)
    preds1 = predict_w_model(encoder, imgs1, level=level, pool=False, device=device, *args, **kwargs)
    if isinstance(scorer, Robustness):
        return scorer(preds0, preds1, pool=pool)
    # a reference-constructed scorer (robustness/eval.py:7-28, re-routed by patch.install)
    if pool:
        preds0, preds1 = torch.mean(preds0.flatten(2), dim=2), torch.mean(preds1.flatten(2), dim=2)
    return scorer(preds0, preds1)


def symmetric_chunks(n: int, rows: int):
    """Row sets closed under the negative permutation perm = [1, 0, n-1, n-2, ..., 2] (robustness/eval.py:22-23), each
    a list of ascending global index ranges whose concatenation, taken as a LOCAL array, has the same permutation
    structure: local rows 0, 1 are the global pair (0, 1) and local row j >= 2 pairs with local row m + 1 - j (m = the
    chunk's length), exactly as global row i >= 2 pairs with n + 1 - i.  A chunk is therefore scored by the same kernel
    as the whole array; rows 0 and 1 ride along in every chunk (their scores are taken from the first one).
    Used to stream representations that do not fit in HBM at once: cfg5's unpooled levels 1-3 are 160 GB per tensor."""
    if n <= max(4, rows):
        return [[(0, n)]]
    chunks, lo, hi = [], 2, n            # rows [lo, hi) of the tail are still unassigned; lo + (hi - 1) == n + 1
    half = max(1, (rows - 2) // 2)
    while lo < hi:
        take = min(half, (hi - lo + 1) // 2)
        a_lo, a_hi = lo, lo + take                      # ascending block from the front ...
        b_lo, b_hi = n + 2 - a_hi, n + 2 - a_lo         # ... and its partners n + 1 - i, also ascending
        if b_lo <= a_hi:                                # the two blocks meet: one contiguous, self-paired range
            chunks.append([(0, 2), (a_lo, hi)])
            break
        chunks.append([(0, 2), (a_lo, a_hi), (b_lo, b_hi)])
        lo, hi = a_hi, b_lo
    return chunks


def robustness_table_streamed(rep_fn, n: int, rows_per_chunk: int, margins: Sequence[float] = (0.0, 0.25, 0.5, 0.75, 1.0),
                              pool: bool = False) -> torch.Tensor:
    """`robustness_table` for representations too large to hold at once: `rep_fn(lo, hi) -> (reps0, reps1)` produces the
    two views' representations of global rows [lo, hi) (e.g. an encoder forward over those images); the rows are
    visited in permutation-closed chunks (`symmetric_chunks`), each scored by the same single-pass kernel.
    -> fp32 (len(margins), 3, n), identical to the unstreamed result."""
    out = None
    for chunk in symmetric_chunks(n, rows_per_chunk):
        parts = [rep_fn(lo, hi) for lo, hi in chunk]
        r0 = torch.cat([p[0] for p in parts])
        r1 = torch.cat([p[1] for p in parts])
        t = robustness_table(r0, r1, margins, pool=pool)            # (M, 3, rows of this chunk)
        if out is None:
            out = torch.empty((t.shape[0], 3, n), dtype=torch.float32, device=t.device)
        pos = 0
        for lo, hi in chunk:            # (rows 0 and 1 are rewritten by every chunk with the same values)
            out[:, :, lo:hi] = t[:, :, pos:pos + (hi - lo)]
            pos += hi - lo
    return out
