"""ORACLE (test infrastructure): fp32 restatement of the reference's loss criteria."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def dice_loss(prediction, mask, batchwise=True, include_background=True, eps=1e-5, apply_softmax=False):
    """segmentation/losses/losses.py:34-58.  prediction (N, C, *spatial) fp32 probabilities, mask
    (N, 1, *spatial) integer labels.  Per class: (2*sum(y*p) + eps) / (sum(y) + sum(p^2) + eps) over
    (N, *spatial) if batchwise else (*spatial); loss = 1 - mean."""
    if apply_softmax:
        prediction = torch.softmax(prediction, dim=1)
    n_classes = prediction.shape[1]
    start = 0 if include_background else 1
    if n_classes == 1:
        if include_background:
            prediction = torch.cat([1 - prediction, prediction], dim=1)  # :46-49
            n_classes = 2
        else:
            start, mask = 0, 1 - mask                                    # :50-52
    axes = tuple(range(0 if batchwise else 1, prediction.dim() - 1))
    mask = mask.reshape(-1, *prediction.shape[2:])
    terms = []
    for c in range(start, n_classes):
        p, y = prediction[:, c], (mask == c)
        inter = torch.sum(y * p, dim=axes, keepdim=True)
        terms.append((2 * inter + eps) / (torch.sum(y, dim=axes, keepdim=True)
                                          + torch.sum(p ** 2, dim=axes, keepdim=True) + eps))
    return 1 - torch.cat(terms).mean()


def bce_loss_plain(prediction, label):
    """classification/losses.py:4-11 (reduction 'mean'; logs are NOT clamped)."""
    return -torch.mean(label * torch.log(prediction) + (1 - label) * torch.log(1 - prediction))


def bce_loss_torch(prediction, label):
    """torch.nn.BCELoss — the framework default loss (utils/default_dict.py:10)."""
    return F.binary_cross_entropy(prediction, label)


def ce_with_softmax(logits, label, label_smoothing=0.0):
    """classification/losses.py:24-25."""
    return F.cross_entropy(logits, label.squeeze(1).long(), label_smoothing=label_smoothing)


def ce_without_softmax(prediction, label, label_smoothing=0.0):
    """classification/losses.py:27-40: log(p) with NaN -> 0 and a floor of -100, one-hot target clamped
    to [s/C, 1 - s/C] (not PyTorch's smoothing formula), summed over classes, mean over pixels."""
    c = prediction.shape[1]
    logp = torch.log(prediction).reshape(prediction.shape[0], c, -1).nan_to_num().clamp(-100)
    target = F.one_hot(label.flatten(1).long(), num_classes=c).moveaxis(-1, 1)
    if label_smoothing:
        target = torch.clamp(target, label_smoothing / c, 1 - label_smoothing / c)
    return (-torch.sum(logp * target, dim=1)).mean()
