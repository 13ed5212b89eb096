"""Hook the B200 path into an UNCHANGED checkout of the reference (SURVEY.md §8b "how to hook in").

    import medsegpretrainimagenet_b200.patch as patch
    patch.install()            # after `sys.path` contains <reference>/src
    # ... then run src/experiment.py (e.g. through runpy); YAML configs stay as they are.

What is replaced, and nothing else (`uninstall()` puts every original back):
  * model.Model.__init__           -> after construction, `convert(self.model)` when the wrapped module is
                                      a U-Net / DeepResNet / sequential FeedForwardModel (sub-block wrappers are left alone)
  * model.FeedForwardModel         -> the SEQUENTIAL compound model the pretraining YAMLs need (`layers:` key,
                                      `.layers` ModuleList, `layers.0.` checkpoint prefix; the shipped class takes
                                      `threads` and runs them in parallel, model/model.py:313-333)
  * segmentation.losses.losses.DiceLoss, classification.losses.{CrossEntropyLoss, BCELoss}
                                   -> forward() routed to the fused loss kernels; options the kernels do not
                                      implement raise instead of being dropped
  * utils.get_class_constr         -> the class paths `torch.nn.CrossEntropyLoss` / `torch.nn.BCELoss` (advanced.yaml:48,
                                      utils/default_dict.py:10) resolve to fused-kernel classes with torch's signature
  * metrics.metrics.ConfusionMatrix.calculate_batch,
    metrics.multiclass_metrics.{MultiClassConfusionMatrix, Top5Accuracy}.calculate_batch
                                   -> single-pass counter kernels
  * robustness.distance.{l2_loss, inv_pearson_corr, cosine_distance}, robustness.eval.{Robustness.__call__,
    predict_w_model, eval_encoder}
Also applies the import shim the reference needs on this Python: the float
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
_undo = []          # (object, attribute, original value) in installation order

_CONVERTIBLE = ("UNet", "DeepResNet", "FeedForwardModel")


def _set(obj, name, value):
    _undo.append((obj, name, getattr(obj, name)))
    setattr(obj, name, value)


def _shim_randint():
    orig = random.randint
    if getattr(orig, "_msp_shim", False):
        return

    def randint(a, b):
        return orig(int(a), int(b))
    randint._msp_shim = True
    _set(random, "randint", randint)


def _reduction_name(loss_mod, fn) -> str:
    """classification/losses.py:8: `self.reduce = Loss.REDUCTION_METHODS[reduction]` — map the stored callable back to
    its name (loss/loss.py:22-24)."""
    for name, f in loss_mod.Loss.REDUCTION_METHODS.items():
        if f is fn:
            return name
    raise NotImplementedError(f"BCELoss: unknown reduction callable {fn!r}")


def _sequential_feed_forward(model_mod):
    """model.FeedForwardModel as the pretraining YAMLs use it: `layers:` applied in sequence."""
    utils = importlib.import_module("utils")
    weight_init = importlib.import_module("model.weight_init")

    class FeedForwardModel(model_mod.CompoundModel):
        @staticmethod
        def fill_kwargs(config_dict):
            for layer_dict in config_dict.elements_of("layers"):
                utils.fill_dict(layer_dict)
                if "weight initialisation" in layer_dict:
                    init_name, init_dict = layer_dict["weight initialisation"].item()
                    init_dict.fill_with_defaults(weight_init.inits_dict[init_name]["arguments"])

        def __init__(self, layers=None, *args, **kwargs):
            if layers is None and "threads" in kwargs:     # the shipped keyword (model/model.py:325)
                layers = kwargs.pop("threads")
            super().__init__(layers, *args, **kwargs)      # CompoundModel: self.layers = ModuleList of Model wrappers

        def forward(self, x, *args, **kwargs):             # replaced by convert(); plain torch only if never converted
            for pass_all, layer in zip(self.pass_all_inputs, self.layers):
                x = layer(x, *args, **kwargs) if pass_all else layer(x)
            return x

    FeedForwardModel.__module__ = model_mod.__name__
    return FeedForwardModel


def install(group=None, convert_models: bool = True) -> None:
    global _installed
    if _installed:
        return
    _shim_randint()
    # run_experiment.py:65 switches torch.use_deterministic_algorithms on for the downstream YAMLs; torch then FILLS every
    # torch.empty() buffer with NaNs (one extra kernel per allocation).  Every buffer this path allocates is fully written
    # by its producing kernel, so the fill is pure overhead here.
    import torch
    _set(torch.utils.deterministic, "fill_uninitialized_memory", False)
    model_mod = importlib.import_module("model")
    model_impl = importlib.import_module("model.model")
    loss_mod = importlib.import_module("loss")
    seg_losses = importlib.import_module("segmentation.losses.losses")
    cls_losses = importlib.import_module("classification.losses")
    met = importlib.import_module("metrics.metrics")
    mmet = importlib.import_module("metrics.multiclass_metrics")
    rdist = importlib.import_module("robustness.distance")
    reval = importlib.import_module("robustness.eval")

    ffm = _sequential_feed_forward(model_impl)
    _set(model_impl, "FeedForwardModel", ffm)
    if hasattr(model_mod, "FeedForwardModel"):
        _set(model_mod, "FeedForwardModel", ffm)

    if convert_models:
        orig_init = model_impl.Model.__init__

        def init(self, *args, **kwargs):
            orig_init(self, *args, **kwargs)
            inner = getattr(self, "model", None)
            if inner is not None and type(inner).__name__ in _CONVERTIBLE:
                _convert.convert(self, group=group)

        _set(model_impl.Model, "__init__", init)

    def dice_forward(self, prediction, mask, *args, **kwargs):
        crit = _losses.DiceLoss(batchwise=not self.axes_start, include_background=self.include_background,
                                smoothing_term=self.eps, apply_softmax=self.softmax, group=group)
        return crit(prediction, mask)

    _set(seg_losses.DiceLoss, "forward", dice_forward)

    def ce_init(self, label_smoothing=0.0, apply_softmax=True, *args, **kwargs):
        # the reference forwards *args / **kwargs to torch.nn.CrossEntropyLoss (weight, ignore_index, reduction, ...;
        # classification/losses.py:18): none of them is implemented by the fused kernel, so refuse rather than drop
        extra = {k: v for k, v in kwargs.items()
                 if not (k == "reduction" and v == "mean") and not (k == "ignore_index" and v == -100)
                 and not (k in ("weight", "size_average", "reduce") and v is None)}
        if args or extra:
            raise NotImplementedError(f"CrossEntropyLoss on the B200 path: unsupported arguments {args} {extra}")
        import torch
        torch.nn.Module.__init__(self)
        self.smooth, self.log_clamp = label_smoothing, -100
        self._msp = _losses.CrossEntropyLoss(label_smoothing, apply_softmax)
        self.forward = self._msp.forward

    _set(cls_losses.CrossEntropyLoss, "__init__", ce_init)

    def bce_forward(self, prediction, label):
        return _losses.BCELoss(_reduction_name(loss_mod, self.reduce))(prediction, label)   # 'none' raises there

    _set(cls_losses.BCELoss, "forward", bce_forward)

    _set(met.ConfusionMatrix, "calculate_batch", _metrics.confusion_calculate_batch)
    _set(mmet.MultiClassConfusionMatrix, "calculate_batch", _metrics.multiclass_calculate_batch)
    _set(mmet.Top5Accuracy, "calculate_batch", _metrics.top5_calculate_batch)

    _set(rdist, "l2_loss", _robust.l2_loss)
    _set(rdist, "inv_pearson_corr", _robust.inv_pearson_corr)
    _set(rdist, "cosine_distance", _robust.cosine_distance)
    _set(reval, "cosine_distance", _robust.cosine_distance)

    # criteria named by their TORCH class path in a YAML (`torch.nn.CrossEntropyLoss`: pretraining/resnet50/advanced.yaml:48;
    # `torch.nn.BCELoss`: the framework default, utils/default_dict.py:10) resolve to the fused-kernel classes with
    # torch's constructor signature; every other class path resolves as before
    import torch
    utils_impl = importlib.import_module("utils._utils")
    utils_pkg = importlib.import_module("utils")
    redirect = {torch.nn.CrossEntropyLoss: _losses.TorchCrossEntropyLoss, torch.nn.BCELoss: _losses.TorchBCELoss}
    orig_get = utils_impl.get_class_constr

    def get_class_constr(class_path):
        cls = orig_get(class_path)
        try:
            return redirect.get(cls, cls)
        except TypeError:               # unhashable
            return cls

    _set(utils_impl, "get_class_constr", get_class_constr)
    if getattr(utils_pkg, "get_class_constr", None) is orig_get:
        _set(utils_pkg, "get_class_constr", get_class_constr)

    def robustness_call(self, preds0, preds1):
        return _robust.Robustness(self.distance_fn, self.margin)(preds0, preds1)

    _set(reval.Robustness, "__call__", robustness_call)
    _set(reval, "predict_w_model", _robust.predict_w_model)
    _set(reval, "eval_encoder", _robust.eval_encoder)
    _installed = True


def uninstall() -> None:
    """Restore everything `install()` replaced (tests that also need the unmodified reference as their oracle)."""
    global _installed
    while _undo:
        obj, name, orig = _undo.pop()
        setattr(obj, name, orig)
    _installed = False
