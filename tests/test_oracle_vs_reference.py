"""Pins the oracle (oracle/ref_*.py) against the REAL reference imported from /root/reference (only
present in the build container; skipped elsewhere — the committed goldens in tests/golden/ carry the same
comparison to the GPU box, see test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import ref_losses, ref_metrics, ref_models, ref_robustness
from oracle import reference_harness as H

pytestmark = pytest.mark.skipif(not H.available(), reason="/root/reference not present")


def _seeded(seed):
    torch.manual_seed(seed)
    return torch.Generator().manual_seed(seed)


def _compare_models(ref, mine, x, train=True, seed=3):
    sd = ref.state_dict()
    assert list(sd.keys()) == list(mine.state_dict().keys()), "state_dict keys / order differ"
    ref_models.load_flat_state_dict(mine, sd)
    for m in (ref, mine):
        m.train(train)
    torch.manual_seed(seed)   # DropPath draws from the global CPU generator (models.py:320-323)
    y_ref = ref(x)
    torch.manual_seed(seed)
    y = mine(x)
    assert y.shape == y_ref.shape
    assert torch.allclose(y, y_ref, rtol=1e-5, atol=1e-6), (y - y_ref).abs().max()
    if train:
        g = torch.randn_like(y_ref)
        y_ref.backward(g)
        y.backward(g)
        gr = {k: p.grad for k, p in ref.model.named_parameters()}
        names = [k.replace(".model.", ".") for k in gr]
        gm = dict(mine.named_parameters())
        gm = {k.replace(".model.", "."): p.grad for k, p in gm.items()}
        for k_ref, k in zip(gr, names):
            assert torch.allclose(gm[k], gr[k_ref], rtol=1e-4, atol=1e-6), k
        # BatchNorm running statistics moved identically
        for (k1, b1), (k2, b2) in zip(ref.state_dict().items(), mine.state_dict().items()):
            assert torch.allclose(b1.float(), b2.float(), rtol=1e-5, atol=1e-6), k1


@pytest.mark.parametrize("train", [True, False])
def test_resnet50_attention_unet(train):
    cd = H.load_config("downstream/acdc/resnet50_attention_unet.yaml")
    ref = H.build_model(cd, seed=0)
    mine = ref_models.resnet50_attention_unet(out_ch=1, final_activation="sigmoid")
    x = torch.rand((2, 3, 64, 64), generator=_seeded(1))
    _compare_models(ref, mine, x, train=train)


def test_resnet50_attention_unet_4class_softmax():
    cd = H.load_config("downstream/acdc/resnet50_attention_unet.yaml",
                       overrides={"model/segmentation.models.UNet/architecture/out_channel_size": 4,
                                  "model/segmentation.models.UNet/architecture/activation_function/final": "softmax"})
    ref = H.build_model(cd, seed=0)
    assert ref.model.decoder.final_block.model.out_channels == 4
    mine = ref_models.resnet50_attention_unet(out_ch=4, final_activation="softmax")
    x = torch.rand((1, 3, 64, 64), generator=_seeded(2))
    _compare_models(ref, mine, x)


def test_basic_unet():
    cd = H.load_config("downstream/covidqu/unet.yaml")
    ref = H.build_model(cd, seed=0)
    in_ch = ref.model.encoder.first_block.model.in_channels
    mine = ref_models.basic_unet(out_ch=1, final_activation="sigmoid", in_channels=in_ch)
    x = torch.rand((1, in_ch, 32, 32), generator=_seeded(4))
    _compare_models(ref, mine, x)


def test_resnet18_encoder_and_classifier_head():
    H.setup()
    from classification import models as ref_cls
    torch.manual_seed(0)
    ref = ref_cls.DeepResNet(bottleneck=False, channel_sizes=(64, 128, 256, 512), widths=(2, 2, 2, 2),
                             in_channels=1, bias=False, head=True, output_size=10, stochastic_depth_rate=0.2)
    mine = ref_models.DeepResNet(bottleneck=False, channel_sizes=(64, 128, 256, 512), widths=(2, 2, 2, 2),
                                 in_channels=1, bias=False, head=True, output_size=10, stochastic_depth_rate=0.2)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    x = torch.rand((3, 1, 64, 64), generator=_seeded(5))
    for train in (True, False):
        ref.train(train), mine.train(train)
        torch.manual_seed(9)
        a, sa = ref(x, return_skip_vals=True)
        torch.manual_seed(9)
        b, sb = mine(x, return_skip_vals=True)
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
        assert len(sa) == len(sb) == 4
        for u, v in zip(sa, sb):
            assert torch.allclose(u, v, rtol=1e-5, atol=1e-6)


def test_kaiming_init_matches_reference_scheme():
    """model/model.py:136-198 with kaiming_normal_: same tensors touched, same RNG consumption."""
    cd = H.load_config("downstream/covidqu/unet.yaml")
    ref = H.build_model(cd, seed=0, init=False)
    in_ch = ref.model.encoder.first_block.model.in_channels
    mine = ref_models.basic_unet(in_channels=in_ch)
    ref_models.load_flat_state_dict(mine, ref.state_dict())
    torch.manual_seed(123)
    ref.init_weight(cd["model"].value())
    torch.manual_seed(123)
    ref_models.kaiming_init_(mine)
    for (k, a), (_, b) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert torch.equal(a, b), k


# ------------------------------------------------------------------------------------------------
def test_losses():
    H.setup()
    from segmentation.losses.losses import DiceLoss
    from classification.losses import BCELoss, CrossEntropyLoss
    g = _seeded(7)
    for c, act in ((1, "sigmoid"), (4, "softmax"), (5, "sigmoid")):
        logits = torch.randn((3, c, 20, 24), generator=g)
        p = torch.sigmoid(logits) if act == "sigmoid" else torch.softmax(logits, 1)
        mask = torch.randint(0, max(c, 2), (3, 1, 20, 24), generator=g)
        for batchwise in (True, False):
            for bg in (True, False):
                if c == 1 and not bg:
                    continue  # the reference mutates `classes_start` on that branch (losses.py:51)
                a = DiceLoss(batchwise=batchwise, include_background=bg)(p.clone().requires_grad_(True), mask)
                b = ref_losses.dice_loss(p, mask, batchwise=batchwise, include_background=bg)
                assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), (c, batchwise, bg)
    p1 = torch.sigmoid(torch.randn((3, 1, 20, 24), generator=g))
    m1 = torch.randint(0, 2, (3, 1, 20, 24), generator=g)
    a = DiceLoss(include_background=False)(p1, m1)
    b = ref_losses.dice_loss(p1, m1, include_background=False)
    assert torch.allclose(a, b)
    # CE with and without softmax, with label smoothing
    logits = torch.randn((6, 10), generator=g)
    lab = torch.randint(0, 10, (6, 1), generator=g)
    for s in (0.0, 0.1):
        assert torch.allclose(CrossEntropyLoss(label_smoothing=s)(logits, lab),
                              ref_losses.ce_with_softmax(logits, lab, s))
    p4 = torch.softmax(torch.randn((2, 4, 9, 9), generator=g), 1)
    p4[0, 1, 0, 0] = 0.0   # log(0) -> floor of -100
    lab4 = torch.randint(0, 4, (2, 1, 9, 9), generator=g)
    lab4[0, 0, 0, 0] = 1
    for s in (0.0, 0.2):
        assert torch.allclose(CrossEntropyLoss(label_smoothing=s, apply_softmax=False)(p4, lab4),
                              ref_losses.ce_without_softmax(p4, lab4, s))
    t = torch.randint(0, 2, (3, 1, 20, 24), generator=g).float()
    assert torch.allclose(BCELoss()(p1, t), ref_losses.bce_loss_plain(p1, t))


def test_confusion_counters():
    H.setup()
    from metrics.metrics import ConfusionMatrix
    from utils.config_dict import ConfigDict
    g = _seeded(11)
    for multilabel, c in ((False, 1), (True, 5)):
        cd = ConfigDict({"metrics": {"calculation": {"multilabel": multilabel, "ignore_nans": True}}})
        cm = ConfusionMatrix(cd, threshold=0.5)
        pred = torch.rand((3, c, 16, 18), generator=g)
        pred[0, 0, 0, :4] = 0.5   # exactly on the threshold: `>=`
        tgt = torch.randint(0, 2, (3, c, 16, 18), generator=g).float()
        tgt[1, 0, 2, :3] = float("nan")
        out = cm.calculate_batch(pred, mask=tgt)
        tp, tn, fp, fn, cc = ref_metrics.confusion_counts(pred, tgt, 0.5, multilabel)
        for name, mine in (("true_positives", tp), ("true_negatives", tn), ("false_positives", fp),
                           ("false_negatives", fn)):
            assert np.array_equal(out[f"{name}_threshold_0.5"].numpy(), mine), name
        assert np.array_equal(np.asarray(cm.class_counts), cc)


def test_multiclass_confusion_and_top5():
    H.setup()
    from metrics.multiclass_metrics import MultiClassConfusionMatrix, Top5Accuracy
    from utils.config_dict import ConfigDict
    g = _seeded(13)
    cd = ConfigDict({"metrics": {"calculation": {"number_of_classes": 4}}})
    cm = MultiClassConfusionMatrix(_config_dict=cd)
    pred = torch.softmax(torch.randn((2, 4, 12, 12), generator=g), 1)
    pred[0, :, 0, 0] = 0.25   # four-way tie: argmax takes the first index
    tgt = torch.randint(0, 4, (2, 1, 12, 12), generator=g)
    out = cm.calculate_batch(pred, mask=tgt)["confusion_matrix"]
    assert np.array_equal(out, ref_metrics.multiclass_confusion_matrix(pred, tgt, 4))
    onehot = torch.nn.functional.one_hot(tgt.squeeze(1), 4).permute(0, 3, 1, 2).float()
    out2 = MultiClassConfusionMatrix(_config_dict=cd).calculate_batch(pred, mask=onehot)["confusion_matrix"]
    assert np.array_equal(out2, out)
    logits = torch.randn((64, 50), generator=g)
    lab = torch.randint(0, 50, (64, 1), generator=g)
    t5 = Top5Accuracy()
    frac = t5.calculate_batch(logits, label=lab)["top_5_accuracy"]
    assert round(frac * 64) == ref_metrics.topk_hits(logits, lab, 5)


def test_derived_metrics():
    H.setup()
    from metrics import metrics as rm
    rng = np.random.default_rng(0)
    cases = [tuple(int(v) for v in rng.integers(0, 50, 4)) for _ in range(40)]
    cases += [(0, 0, 0, 0), (0, 5, 0, 0), (3, 0, 0, 0), (0, 0, 7, 0), (0, 0, 0, 9), (4, 4, 0, 0)]
    for tp, tn, fp, fn in cases:
        pv = dict(true_positives=torch.tensor(tp), true_negatives=torch.tensor(tn),
                  false_positives=torch.tensor(fp), false_negatives=torch.tensor(fn))

        def ref_val(metric):
            return list(metric.evaluate_batch(pv).values())[0]

        def norm(v, neutral):
            return neutral if v == "invalid" else v

        assert ref_val(rm.DiceIndex()) == norm(ref_metrics.dice_index(tp, fp, fn), 1)
        assert ref_val(rm.JaccardIndex()) == norm(ref_metrics.jaccard_index(tp, fp, fn), 1)
        assert ref_val(rm.MCC()) == norm(ref_metrics.mcc(tp, fp, fn, tn), 0)
        assert ref_val(rm.BalancedAccuracy()) == norm(ref_metrics.balanced_accuracy(tp, tn, fp, fn), 0)
        if tp + tn + fp + fn:
            assert ref_val(rm.Accuracy()) == ref_metrics.accuracy(tp, fp, tn, fn)


def test_robustness():
    H.setup()
    from robustness import distance as rd
    from robustness.eval import Robustness
    g = _seeded(17)
    for n in (2, 3, 8, 9):
        q = torch.relu(torch.randn((n, 6, 5, 5), generator=g))
        k = torch.relu(q + 0.1 * torch.randn((n, 6, 5, 5), generator=g))
        for ref_fn, my_fn in ((rd.cosine_distance, ref_robustness.cosine_distance),
                              (rd.l2_loss, ref_robustness.l2_distance),
                              (rd.inv_pearson_corr, ref_robustness.inv_pearson)):
            assert torch.allclose(ref_fn(q.flatten(1), k.flatten(1)), my_fn(q.flatten(1), k.flatten(1)))
            for margin in (0.0, 0.5):
                a = Robustness(ref_fn, margin)(q, k)
                b = ref_robustness.robustness_scores(q, k, my_fn, margin)
                assert torch.allclose(a, b), (n, margin)
    assert ref_robustness.negative_permutation(6) == [1, 0, 5, 4, 3, 2]
