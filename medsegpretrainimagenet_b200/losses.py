"""Loss criteria with the reference's constructor signatures and call contract
(`criterion(prediction, target) -> 0-dim tensor` carrying an autograd graph, wrapped by loss/loss.py:69-95
which calls `.item()` and then `.backward()`), backed by the fused kernels of csrc/msp_loss.cu.

Usable from YAML as `medsegpretrainimagenet_b200.losses.DiceLoss` etc., or swapped in for the
reference's criteria by `patch.install()`.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from . import ops


def _world(group) -> int:
    if group is not None and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensors only (the B200 path has no CPU fallback)")


class _Dice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, mask, two_class, label_offset, batchwise, class_start, eps, group):
        sums = ops.dice_sums(prob, mask, two_class, label_offset, batchwise)
        world = _world(group)
        if batchwise and world > 1:
            # a ratio of GLOBAL sums: 3*C doubles over NCCL
            dist.all_reduce(sums, group=group)
        loss, coef = ops.dice_finalize(sums, class_start, eps)
        ctx.cfg = (two_class, label_offset, batchwise, world if batchwise else 1)
        ctx.save_for_backward(prob, mask, coef)
        return loss

    @staticmethod
    def backward(ctx, g):
        two_class, label_offset, batchwise, world = ctx.cfg
        prob, mask, coef = ctx.saved_tensors
        # gradient averaging over ranks follows; the global-ratio gradient is a SUM over ranks
        dprob = ops.dice_bwd(prob, mask, two_class, label_offset, batchwise, coef, gscale=float(world),
                             gscale_dev=g.contiguous())
        return dprob, None, None, None, None, None, None, None


class DiceLoss(nn.Module):
    """segmentation/losses/losses.py:11-58.  `prediction` (N, C, *spatial) fp32 probabilities, `mask`
    (N, 1, *spatial) integer labels."""

    def __init__(self, batchwise=True, include_background=True, smoothing_term=1e-5, apply_softmax=False,
                 group=None, *args, **kwargs):
        super().__init__()
        if apply_softmax:
            raise NotImplementedError("DiceLoss(apply_softmax=True): apply the softmax as the model's final "
                                      "activation (fused into the head kernel) instead")
        self.eps = smoothing_term
        self.batchwise = bool(batchwise)
        self.include_background = include_background
        self.group = group

    def forward(self, prediction, mask, *args, **kwargs):
        _need_cuda(prediction, "DiceLoss")
        prediction = prediction.contiguous()
        if prediction.dtype != torch.float32:
            prediction = prediction.float()
        n_classes = prediction.shape[1]
        two_class, label_offset, class_start = False, 0, int(not self.include_background)
        if n_classes == 1:
            if self.include_background:
                two_class = True                      # [1 - p, p]            (losses.py:46-49)
            else:
                label_offset, class_start = 1, 0      # p against (mask == 1)  (losses.py:50-52)
        mask = mask.reshape(prediction.shape[0], -1)
        if mask.dtype != torch.int64:
            mask = mask.long()
        return _Dice.apply(prediction, mask.contiguous(), two_class, label_offset, self.batchwise,
                           class_start, float(self.eps), self.group)


class _SumLoss(torch.autograd.Function):
    """Shared shape of the mean-reduced losses: forward = one reduction kernel, backward = the same
    kernel emitting the gradient scaled by autograd's incoming (device-resident) factor."""

    @staticmethod
    def forward(ctx, pred, target, kind, arg, scale):
        if kind == "ce_prob":
            ls, _ = ops.ce_prob(pred, target, arg)
        elif kind == "bce":
            ls, _ = ops.bce(pred, target, arg)
        elif kind == "softmax_ce_soft":
            ls, _ = ops.softmax_ce_soft(pred, target, arg)
        elif kind == "softmax_ce_spatial":
            ls, _ = ops.softmax_ce_spatial(pred, target, arg)
        else:
            ls, _ = ops.softmax_ce(pred, target, arg)
        ctx.cfg = (kind, arg, scale)
        ctx.save_for_backward(pred, target)
        return ops.sum_to_mean(ls, scale)

    @staticmethod
    def backward(ctx, g):
        kind, arg, scale = ctx.cfg
        pred, target = ctx.saved_tensors
        g = g.contiguous()
        if kind == "ce_prob":
            _, d = ops.ce_prob(pred, target, arg, gscale=scale, gscale_dev=g, want_loss=False, want_grad=True)
        elif kind == "bce":
            _, d = ops.bce(pred, target, arg, gscale=scale, gscale_dev=g, want_loss=False, want_grad=True)
        elif kind == "softmax_ce_soft":
            _, d = ops.softmax_ce_soft(pred, target, arg, gscale=scale, gscale_dev=g, want_loss=False, want_grad=True)
        elif kind == "softmax_ce_spatial":
            _, d = ops.softmax_ce_spatial(pred, target, arg, gscale=scale, gscale_dev=g, want_loss=False,
                                          want_grad=True)
        else:
            _, d = ops.softmax_ce(pred, target, arg, gscale=scale, gscale_dev=g, want_loss=False,
                                  want_grad=True)
        return d, None, None, None, None


def _f32c(t):
    t = t.contiguous()
    return t if t.dtype == torch.float32 else t.float()


class CrossEntropyLoss(nn.Module):
    """classification/losses.py:13-40: `apply_softmax=True` -> F.cross_entropy(logits (N, C),
    label (N, 1), label_smoothing); `apply_softmax=False` -> pixel-wise CE on probabilities with the
    log floor of -100 and the clamped one-hot target."""

    def __init__(self, label_smoothing=0.0, apply_softmax=True, *args, **kwargs):
        super().__init__()
        if label_smoothing >= 0.5:
            raise ValueError("Label smoothing value should be <0.5")
        self.smooth = float(label_smoothing)
        self.apply_softmax = apply_softmax

    def forward(self, prediction, label, *args, **kwargs):
        _need_cuda(prediction, "CrossEntropyLoss")
        prediction = _f32c(prediction)
        if self.apply_softmax:
            return softmax_cross_entropy(prediction, label, self.smooth, squeeze_label=True)
        lab = label.flatten(1).long().contiguous()
        n_pix = prediction.shape[0] * prediction[0, 0].numel()
        return _SumLoss.apply(prediction, lab, "ce_prob", self.smooth, 1.0 / n_pix)


def softmax_cross_entropy(prediction, label, smooth: float, squeeze_label: bool = False):
    """F.cross_entropy(prediction, label, label_smoothing=smooth) with mean reduction, for the three target forms torch
    accepts: (N, C) logits with class indices (N,) [(N, 1) when `squeeze_label`: classification/losses.py:25], (N, C)
    logits with class PROBABILITIES (N, C) (Mixup / CutMix: config/pretraining/resnet50/advanced.yaml:17,48), and
    spatial logits (N, C, *spatial) with class indices (N, *spatial)."""
    prediction = _f32c(prediction)
    if prediction.dim() == 2 and label.is_floating_point() and tuple(label.shape) == tuple(prediction.shape):
        return _SumLoss.apply(prediction, _f32c(label), "softmax_ce_soft", smooth, 1.0 / prediction.shape[0])
    if label.is_floating_point() and tuple(label.shape) == tuple(prediction.shape):
        raise NotImplementedError("softmax cross entropy with probability targets is implemented for (N, C) logits")
    if prediction.dim() == 2:
        lab = (label.squeeze(1) if (squeeze_label and label.dim() == 2) else label).long().contiguous()
        if lab.dim() != 1 or lab.shape[0] != prediction.shape[0]:
            raise ValueError(f"cross entropy: target shape {tuple(label.shape)} does not match logits {tuple(prediction.shape)}")
        return _SumLoss.apply(prediction, lab, "softmax_ce", smooth, 1.0 / prediction.shape[0])
    n = prediction.shape[0]
    n_pix = n * prediction[0, 0].numel()
    lab = label
    if lab.dim() == prediction.dim() and lab.shape[1] == 1:
        lab = lab.squeeze(1)
    if lab.numel() != n_pix:
        raise ValueError(f"cross entropy: target shape {tuple(label.shape)} does not match logits {tuple(prediction.shape)}")
    lab = lab.reshape(n, -1).long().contiguous()
    return _SumLoss.apply(prediction, lab, "softmax_ce_spatial", smooth, 1.0 / n_pix)


class TorchCrossEntropyLoss(nn.Module):
    """torch.nn.CrossEntropyLoss(weight=None, ignore_index=-100, reduction='mean', label_smoothing=0.0) on the fused
    kernels — the criterion config/pretraining/resnet50/advanced.yaml:48 names by its torch class path (`patch.install()`
    redirects that path here).  Class weights, a non-default ignore_index and reductions other than 'mean' raise."""

    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction="mean",
                 label_smoothing=0.0):
        super().__init__()
        if weight is not None or size_average is not None or reduce is not None or ignore_index != -100 \
                or reduction != "mean":
            raise NotImplementedError("CrossEntropyLoss on the B200 path: weight / ignore_index / reduction other than "
                                      "'mean' are not implemented")
        self.label_smoothing = float(label_smoothing)

    def forward(self, input, target):
        _need_cuda(input, "CrossEntropyLoss")
        return softmax_cross_entropy(input, target, self.label_smoothing)


class TorchBCELoss(nn.Module):
    """torch.nn.BCELoss(weight=None, reduction='mean') — the framework's default loss (utils/default_dict.py:10)."""

    def __init__(self, weight=None, size_average=None, reduce=None, reduction="mean"):
        super().__init__()
        if weight is not None or size_average is not None or reduce is not None:
            raise NotImplementedError("BCELoss on the B200 path: element weights are not implemented")
        self._msp = BCELoss(reduction, torch_semantics=True)

    def forward(self, input, target):
        return self._msp(input, target)


class BCELoss(nn.Module):
    """classification/losses.py:4-11 (`torch_semantics=False`: plain logs) or torch.nn.BCELoss
    (`torch_semantics=True`, the framework default loss, utils/default_dict.py:10)."""

    def __init__(self, reduction="mean", torch_semantics=False, *args, **kwargs):
        super().__init__()
        if reduction not in ("mean", "sum"):
            raise NotImplementedError(f"BCELoss reduction {reduction!r}")
        self.reduction = reduction
        self.clamp = int(bool(torch_semantics))

    def forward(self, prediction, label, *args, **kwargs):
        _need_cuda(prediction, "BCELoss")
        prediction = _f32c(prediction)
        label = _f32c(label.reshape(prediction.shape))
        scale = 1.0 / prediction.numel() if self.reduction == "mean" else 1.0
        return _SumLoss.apply(prediction, label, "bce", self.clamp, scale)


class Loss:
    """Host-side mirror of the reference's `Loss` wrapper (loss/loss.py:8-114) for use WITHOUT the checkout (with it,
    `patch.install()` leaves the reference's own wrapper in place and only re-routes the criteria): picks
    `batch['prediction']` and `batch[label_type]`, divides by the accumulation scale, reads the value with `.item()`
    (:85, the step's one host read), runs `.backward()` in training (:86-87) and keeps the running sums the epoch logs
    are made of."""

    def __init__(self, criterion, name=None, label_type="mask", accumulate=True):
        import re
        self.calculator = criterion
        cls = type(criterion).__name__
        snake = re.sub("([a-z0-9])([A-Z])", r"\1_\2", re.sub("(.)([A-Z][a-z]+)", r"\1_\2", cls)).lower()
        self.name = name or getattr(criterion, "name", snake)
        self.label_type = label_type
        self.accumulate = accumulate
        self.value, self.num_batches = 0, 0
        self.acc_value, self.num_batch_fragments = 0, 0
        self.train = True

    def calculate_batch(self, batch, cumulate=True, train=True, average=True, accumulation_scale=1, last=False):
        self.train = train
        loss = self.calculator(batch["prediction"], batch[self.label_type])
        if average:
            loss = loss / accumulation_scale
        value = loss.item()
        if train and not last:
            loss.backward()
        if cumulate:
            if self.accumulate:
                self.acc_value += value
                self.num_batch_fragments += 1
            else:
                self.value += value
                self.num_batches += 1
        return {self.name: value}

    def evaluate_batch(self, *args, cumulate=True, flush=True, **kwargs):
        value = self.acc_value if self.accumulate else self.value
        if flush:
            self.num_batch_fragments, self.acc_value = 0, 0
        if cumulate:
            self.value += value
            self.num_batches += 1
        return {self.name: value}

    def evaluate_epoch(self, *args, flush=True, average=True, **kwargs):
        value = self.value
        if average and self.num_batches > 0:
            value = value / self.num_batches
        if flush:
            self.value, self.num_batches = 0, 0
        return {self.name: value}
