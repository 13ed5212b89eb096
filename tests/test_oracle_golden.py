"""The oracle (oracle/ref_*.py) against the committed golden fixtures tests/golden/*.pt, which hold the outputs
of the REAL reference on seeded inputs (minted by tools/make_golden.py in the build container).  Runs anywhere
(CPU, no /root/reference needed): this is what pins the oracle on the GPU box."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import ref_losses, ref_metrics, ref_models, ref_robustness
from oracle.seeded_weights import fill_state_

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def oracle_loss(case, pred):
    k, t = case["kind"], case["target"]
    if k == "dice":
        return ref_losses.dice_loss(pred, t, batchwise=case["batchwise"], include_background=case["include_background"])
    if k == "ce_softmax":
        return ref_losses.ce_with_softmax(pred, t, case["smoothing"])
    if k == "ce_prob":
        return ref_losses.ce_without_softmax(pred, t, case["smoothing"])
    if k == "bce_plain":
        return ref_losses.bce_loss_plain(pred, t)
    if k == "bce_torch":
        return ref_losses.bce_loss_torch(pred, t)
    raise KeyError(k)


def test_losses_match_reference_goldens():
    cases = load("losses")
    assert len(cases) >= 15
    for case in cases:
        p = case["pred"].clone().requires_grad_(True)
        loss = oracle_loss(case, p)
        loss.backward()
        assert torch.allclose(loss, case["loss"], rtol=1e-6, atol=1e-7), case["kind"]
        assert torch.allclose(p.grad, case["grad"], rtol=1e-5, atol=1e-8), case["kind"]


def test_metrics_match_reference_goldens():
    g = load("metrics")
    for case in g["binary"]:
        tp, tn, fp, fn, cc = ref_metrics.confusion_counts(case["pred"], case["target"], case["threshold"], case["multilabel"])
        for name, mine in (("true_positives", tp), ("true_negatives", tn), ("false_positives", fp), ("false_negatives", fn),
                           ("class_counts", cc)):
            assert np.array_equal(np.asarray(mine), case[name].numpy()), name      # bit-exact integers
    for case in g["multiclass"]:
        cm = ref_metrics.multiclass_confusion_matrix(case["pred"], case["target"], case["classes"])
        assert np.array_equal(np.asarray(cm).astype(np.int64), case["confusion_matrix"].numpy())
    for case in g["top5"]:
        assert ref_metrics.topk_hits(case["pred"], case["target"], 5) == case["hits"]
    neutral = dict(dice=1, jaccard=1, mcc=0, balanced_accuracy=0)
    for row in g["derived"]:
        tp, tn, fp, fn = row["tp"], row["tn"], row["fp"], row["fn"]
        mine = dict(dice=ref_metrics.dice_index(tp, fp, fn), jaccard=ref_metrics.jaccard_index(tp, fp, fn),
                    mcc=ref_metrics.mcc(tp, fp, fn, tn), balanced_accuracy=ref_metrics.balanced_accuracy(tp, tn, fp, fn))
        for k, v in mine.items():
            v = neutral[k] if v == "invalid" else v
            assert float(v) == row[k], (k, tp, tn, fp, fn)
        if "accuracy" in row:
            assert float(ref_metrics.accuracy(tp, fp, tn, fn)) == row["accuracy"]


def test_robustness_matches_reference_goldens():
    fns = dict(cosine=ref_robustness.cosine_distance, l2=ref_robustness.l2_distance, pearson=ref_robustness.inv_pearson)
    for case in load("robustness"):
        q, k = case["q"], case["k"]
        for name, fn in fns.items():
            assert torch.allclose(fn(q.flatten(1), k.flatten(1)), case["dist"][name], rtol=1e-6, atol=1e-7), name
        for (name, margin), ref in case["scores"].items():
            assert torch.allclose(ref_robustness.robustness_scores(q, k, fns[name], margin), ref, rtol=1e-6, atol=1e-7)


def _digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def oracle_model(name):
    if name == "r50_attention_unet_binary":
        return fill_state_(ref_models.resnet50_attention_unet(out_ch=1, final_activation="sigmoid"), 100)
    if name == "r50_attention_unet_4class":
        return fill_state_(ref_models.resnet50_attention_unet(out_ch=4, final_activation="softmax"), 100)
    if name == "basic_unet_binary":
        return None
    if name == "deepresnet18_head10":
        return fill_state_(ref_models.DeepResNet(bottleneck=False, channel_sizes=(64, 128, 256, 512), widths=(2, 2, 2, 2),
                                                 in_channels=1, bias=False, head=True, output_size=10,
                                                 stochastic_depth_rate=0.2), 101)
    raise KeyError(name)


@pytest.mark.parametrize("idx", range(4))
def test_models_match_reference_goldens(idx):
    case = load("models")[idx]
    m = oracle_model(case["name"])
    if m is None:
        m = fill_state_(ref_models.basic_unet(out_ch=1, final_activation="sigmoid", in_channels=case["x"].shape[1]), 100)
    assert _digest(m.state_dict()) == case["state_digest"], "seeded weights differ from the reference's"
    x = case["x"]
    if case["name"] == "deepresnet18_head10":
        m.train()
        torch.manual_seed(9)
        y, skips = m(x, return_skip_vals=True)
        assert torch.allclose(y, case["y_train"], rtol=1e-4, atol=1e-5)
        assert np.allclose([float(s.mean()) for s in skips], case["skip_means_train"], rtol=1e-4)
        m.eval()
        with torch.no_grad():
            y, skips = m(x, return_skip_vals=True)
        assert torch.allclose(y, case["y_eval"], rtol=1e-4, atol=1e-5)
        assert np.allclose([float(s.mean()) for s in skips], case["skip_means_eval"], rtol=1e-4)
        return
    m.train()
    torch.manual_seed(3)
    y = m(x)
    assert torch.allclose(y, case["y_train"], rtol=1e-4, atol=1e-5), (y - case["y_train"]).abs().max()
    loss = ref_losses.dice_loss(y, case["mask"])
    assert abs(float(loss) - case["loss"]) <= 1e-5 * abs(case["loss"])
    loss.backward()
    for k, p in m.named_parameters():
        k = k.replace(".model.", ".")
        if k in case["grad_norms"]:
            ref = case["grad_norms"][k]
            assert abs(float(p.grad.norm()) - ref) <= 1e-3 * ref + 1e-7, k
    m.eval()
    with torch.no_grad():
        assert torch.allclose(m(x), case["y_eval"], rtol=1e-4, atol=1e-5)
