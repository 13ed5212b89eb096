"""Hook the B200 path into an UNCHANGED checkout of the reference (SURVEY.md §8b "how to hook in").

    import medsegpretrainimagenet_b200.patch as patch
    patch.install()            # after `sys.path` contains <reference>/src
    # ... then run src/experiment.py (e.g. through runpy); YAML configs stay as they are.

What is replaced, and nothing else:
  * model.Model.__init__           -> after construction, `convert(self.model)` when the wrapped module is
                                      a U-Net / DeepResNet (sub-block wrappers are left alone)
  * segmentation.losses.losses.DiceLoss, classification.losses.{CrossEntropyLoss, BCELoss}
                                   -> forward() routed to the fused loss kernels
  * metrics.metrics.ConfusionMatrix.calculate_batch,
    metrics.multiclass_metrics.{MultiClassConfusionMatrix, Top5Accuracy}.calculate_batch
                                   -> single-pass counter kernels
  * robustness.distance.{l2_loss, inv_pearson_corr, cosine_distance}, robustness.eval.Robustness.__call__
Also applies the two import shims the reference needs on this Python: the float
arguments of `random.randint` at run_experiment.py:35.
"""
from __future__ import annotations

import importlib
import random

from . import converter as _convert
from . import losses as _losses
from . import metrics as _metrics
from . import robustness as _robust

_installed = False


def _shim_randint():
    orig = random.randint
    if getattr(orig, "_msp_shim", False):
        return
    def randint(a, b):
        return orig(int(a), int(b))
    randint._msp_shim = True
    random.randint = randint


def install(group=None, convert_models: bool = True) -> None:
    global _installed
    if _installed:
        return
    _shim_randint()
    model_mod = importlib.import_module("model")
    seg_losses = importlib.import_module("segmentation.losses.losses")
    cls_losses = importlib.import_module("classification.losses")
    met = importlib.import_module("metrics.metrics")
    mmet = importlib.import_module("metrics.multiclass_metrics")
    rdist = importlib.import_module("robustness.distance")
    reval = importlib.import_module("robustness.eval")

    if convert_models:
        orig_init = model_mod.Model.__init__

        def init(self, *args, **kwargs):
            orig_init(self, *args, **kwargs)
            inner = getattr(self, "model", None)
            if inner is not None and type(inner).__name__ in ("UNet", "DeepResNet"):
                _convert.convert(self, group=group)

        model_mod.Model.__init__ = init

    def dice_forward(self, prediction, mask, *args, **kwargs):
        crit = _losses.DiceLoss(batchwise=not self.axes_start, include_background=self.include_background,
                                smoothing_term=self.eps, apply_softmax=self.softmax, group=group)
        return crit(prediction, mask)

    seg_losses.DiceLoss.forward = dice_forward

    def ce_init(self, label_smoothing=0.0, apply_softmax=True, *args, **kwargs):
        import torch
        torch.nn.Module.__init__(self)
        self._msp = _losses.CrossEntropyLoss(label_smoothing, apply_softmax)
        self.forward = self._msp.forward

    cls_losses.CrossEntropyLoss.__init__ = ce_init

    def bce_forward(self, prediction, label):
        return _losses.BCELoss("mean")(prediction, label)

    cls_losses.BCELoss.forward = bce_forward

    met.ConfusionMatrix.calculate_batch = _metrics.confusion_calculate_batch
    mmet.MultiClassConfusionMatrix.calculate_batch = _metrics.multiclass_calculate_batch
    mmet.Top5Accuracy.calculate_batch = _metrics.top5_calculate_batch

    rdist.l2_loss = _robust.l2_loss
    rdist.inv_pearson_corr = _robust.inv_pearson_corr
    rdist.cosine_distance = _robust.cosine_distance
    reval.cosine_distance = _robust.cosine_distance

    def robustness_call(self, preds0, preds1):
        return _robust.Robustness(self.distance_fn, self.margin)(preds0, preds1)

    reval.Robustness.__call__ = robustness_call
    reval.predict_w_model = _robust.predict_w_model
    _installed = True
