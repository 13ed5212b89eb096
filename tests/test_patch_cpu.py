"""`patch.install()` against the REAL reference checkout (build container only; skipped on the GPU box, where
tests/test_predict_loop_gpu.py drives the same call pattern over reference-structured objects): the north star's
"src/experiment.py runs unchanged from its YAML configs" route.  Without a GPU nothing can execute — what is checked is
the wiring: Model gets converted, the criteria / counters / distances are re-bound, the sequential FeedForwardModel makes
the pretraining YAML buildable, unsupported options raise instead of being dropped, and uninstall() restores
the reference exactly."""
import pytest
import torch

from oracle import reference_harness as H

pytestmark = pytest.mark.skipif(not H.available(), reason="/root/reference not present")


@pytest.fixture()
def patched():
    H.setup()
    import medsegpretrainimagenet_b200.patch as patch
    patch.install()
    try:
        yield patch
    finally:
        patch.uninstall()


def test_install_converts_models_and_binds_losses_and_metrics(patched):
    import medsegpretrainimagenet_b200 as b200
    from medsegpretrainimagenet_b200 import metrics as M
    cd = H.load_config("downstream/covidqu/unet.yaml")
    model = H.build_model(cd, seed=0)
    assert b200.is_converted(model), "model.Model.__init__ must route the wrapped U-Net through convert()"
    # sub-block wrappers stay plain reference modules (only the top-level network owns an execution context)
    assert not hasattr(model.model.encoder.first_block, "_msp_ctx")
    keys = list(model.state_dict().keys())
    assert keys[0] == "encoder.first_block.weight" and not any(".model." in k for k in keys)
    # there is no CPU fallback behind the converted forward
    with pytest.raises(RuntimeError, match="CUDA"):
        model(x=torch.rand(1, 3, 32, 32))
    loss_fn = H.build_loss(cd)
    import segmentation.losses.losses as seg_losses
    assert type(loss_fn.calculator) is seg_losses.DiceLoss            # the reference's own class ...
    assert seg_losses.DiceLoss.forward.__name__ == "dice_forward"     # ... whose forward is the fused kernel path
    with pytest.raises(RuntimeError, match="CUDA"):
        loss_fn.calculator(torch.rand(2, 1, 8, 8), torch.zeros(2, 1, 8, 8, dtype=torch.long))
    mc = H.build_metrics(cd, loss_fn)
    import metrics.metrics as met
    import metrics.multiclass_metrics as mmet
    assert met.ConfusionMatrix.calculate_batch is M.confusion_calculate_batch
    assert mmet.MultiClassConfusionMatrix.calculate_batch is M.multiclass_calculate_batch
    assert mmet.Top5Accuracy.calculate_batch is M.top5_calculate_batch
    parents = [v["calculator"] for k, v in mc.metrics.items() if k.startswith("confusion_matrix")]
    assert parents and all(type(p) is met.ConfusionMatrix for p in parents)
    import robustness.distance as rdist
    import robustness.eval as reval
    from medsegpretrainimagenet_b200 import robustness as R
    assert rdist.cosine_distance is R.cosine_distance and reval.predict_w_model is R.predict_w_model
    assert reval.eval_encoder is R.eval_encoder


def test_resnet50_attention_unet_yaml_builds_and_converts(patched):
    import medsegpretrainimagenet_b200 as b200
    cd = H.load_config("downstream/acdc/resnet50_attention_unet.yaml")
    model = H.build_model(cd, seed=0)
    assert b200.is_converted(model)
    assert b200.is_converted(model.model.encoder)      # the DeepResNet's own wrapper is converted too (eval_encoder)
    assert sum(p.numel() for p in model.parameters()) == 55_668_321      # SURVEY: 55.67 M (1 output channel; bench cfg3 with 4 classes: + 3 * 17)


def test_pretraining_yaml_builds_through_the_sequential_feed_forward_model(patched):
    """config/pretraining/resnet50/simple.yaml names `model.FeedForwardModel: {layers: [...]}`; the shipped class takes
    `threads` (model/model.py:325) -> TypeError without the patch."""
    import medsegpretrainimagenet_b200 as b200
    cd = H.load_config("pretraining/resnet50/simple.yaml")
    model = H.build_model(cd, seed=0)
    inner = model.model
    assert type(inner).__name__ == "FeedForwardModel" and len(inner.layers) == 4
    assert b200.is_converted(model)
    keys = list(model.state_dict().keys())
    assert keys[0] == "layers.0.stem.0.weight" and keys[-2:] == ["layers.3.weight", "layers.3.bias"]
    assert sum(p.numel() for p in model.parameters()) == 22_780_456       # SURVEY: ResNet-50 classifier, 22.78 M
    # checkpoint hand-off: the `layers.0.` entries initialise a U-Net's encoder (unet_models.py:555-588)
    from medsegpretrainimagenet_b200 import models
    enc = models.encoder_state_dict(model.state_dict())
    assert "stem.0.weight" in enc and len(enc) == len(keys) - 2
    # the loss of that YAML is the reference class with the fused forward
    loss_fn = H.build_loss(cd)
    assert hasattr(loss_fn.calculator, "_msp") and loss_fn.calculator._msp.smooth == pytest.approx(0.1)


def test_unsupported_loss_options_raise_instead_of_being_dropped(patched):
    import classification.losses as cls_losses
    with pytest.raises(NotImplementedError):
        cls_losses.CrossEntropyLoss(label_smoothing=0.1, ignore_index=3)
    with pytest.raises(NotImplementedError):
        cls_losses.CrossEntropyLoss(weight=torch.ones(4))
    cls_losses.CrossEntropyLoss(label_smoothing=0.1, reduction="mean")        # the default spelled out is fine
    bce = cls_losses.BCELoss(reduction="none")
    with pytest.raises(NotImplementedError, match="none"):
        bce(torch.rand(2, 3), torch.rand(2, 3))
    assert patched._reduction_name(__import__("loss"), cls_losses.BCELoss("sum").reduce) == "sum"


def test_uninstall_restores_the_reference():
    H.setup()
    import medsegpretrainimagenet_b200.patch as patch
    import metrics.metrics as met
    import model.model as model_impl
    import segmentation.losses.losses as seg_losses
    before = (met.ConfusionMatrix.calculate_batch, model_impl.Model.__init__, seg_losses.DiceLoss.forward,
              model_impl.FeedForwardModel)
    patch.install()
    assert met.ConfusionMatrix.calculate_batch is not before[0]
    patch.uninstall()
    after = (met.ConfusionMatrix.calculate_batch, model_impl.Model.__init__, seg_losses.DiceLoss.forward,
             model_impl.FeedForwardModel)
    assert before == after
    # and the unmodified reference still computes on the CPU
    crit = seg_losses.DiceLoss()
    assert torch.isfinite(crit(torch.rand(2, 1, 8, 8), torch.zeros(2, 1, 8, 8, dtype=torch.long)))


def test_advanced_pretraining_yaml_resolves_torch_criterion_to_the_fused_class(patched):
    """config/pretraining/resnet50/advanced.yaml:48 names `torch.nn.CrossEntropyLoss` (soft Mixup / CutMix labels); the
    class path resolves to the fused-kernel class with torch's constructor, other torch classes resolve unchanged."""
    import utils
    from medsegpretrainimagenet_b200 import losses
    assert utils.get_class_constr("torch.nn.CrossEntropyLoss") is losses.TorchCrossEntropyLoss
    assert utils.get_class_constr("torch.nn.BCELoss") is losses.TorchBCELoss
    assert utils.get_class_constr("torch.nn.Conv2d") is torch.nn.Conv2d
    cd = H.load_config("pretraining/resnet50/advanced.yaml")
    loss_fn = H.build_loss(cd)
    assert isinstance(loss_fn.calculator, losses.TorchCrossEntropyLoss)
    assert loss_fn.calculator.label_smoothing == pytest.approx(0.1) and loss_fn.label_type == "label"
    with pytest.raises(NotImplementedError):
        losses.TorchCrossEntropyLoss(ignore_index=0)
    model = H.build_model(cd, seed=0)          # stochastic_depth_rate 0.1 sequential model
    import medsegpretrainimagenet_b200 as b200
    assert b200.is_converted(model) and type(model.model).__name__ == "FeedForwardModel"
