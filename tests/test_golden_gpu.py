"""The CUDA path against the committed golden fixtures (outputs of the REAL reference, tools/make_golden.py),
without the oracle in between.  Integer results bit-exact; fp32 kernels (losses, distances) within the stated
tolerance; bf16 networks within the bf16-storage tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle.seeded_weights import fill_state_
from oracle import ref_models

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def _b():
    import medsegpretrainimagenet_b200 as b
    return b


def test_losses_vs_reference_goldens():
    b = _b()
    for case in load("losses"):
        k = case["kind"]
        crit = {"dice": lambda: b.losses.DiceLoss(batchwise=case["batchwise"], include_background=case["include_background"]),
                "ce_softmax": lambda: b.losses.CrossEntropyLoss(case["smoothing"]),
                "ce_prob": lambda: b.losses.CrossEntropyLoss(case["smoothing"], apply_softmax=False),
                "bce_plain": lambda: b.losses.BCELoss(torch_semantics=False),
                "bce_torch": lambda: b.losses.BCELoss(torch_semantics=True)}[k]()
        p = case["pred"].to(DEV).requires_grad_(True)
        loss = crit(p, case["target"].to(DEV))
        loss.backward()
        ref = float(case["loss"])
        assert abs(loss.item() - ref) <= 1e-5 * max(1.0, abs(ref)), k           # fp32 reductions
        scale = case["grad"].abs().max().item() + 1e-30
        assert (p.grad.cpu() - case["grad"]).abs().max().item() <= 1e-4 * scale, k


def test_metrics_vs_reference_goldens_bit_exact():
    b = _b()
    g = load("metrics")
    for case in g["binary"]:
        got = b.metrics.binary_confusion_counts(case["pred"].to(DEV), case["target"].to(DEV), case["threshold"],
                                                per_channel=case["multilabel"])
        for mine, name in (("TP", "true_positives"), ("TN", "true_negatives"), ("FP", "false_positives"),
                           ("FN", "false_negatives"), ("class_counts", "class_counts")):
            assert np.array_equal(got[mine].cpu().numpy(), case[name].numpy()), name
    for case in g["multiclass"]:
        cm = b.metrics.multiclass_confusion_matrix(case["pred"].to(DEV), case["target"].to(DEV))
        assert np.array_equal(cm.cpu().numpy(), case["confusion_matrix"].numpy())
    for case in g["top5"]:
        hits, num = b.metrics.topk_correct(case["pred"].to(DEV), case["target"].to(DEV), 5)
        assert int(hits.item()) == case["hits"] and num == case["target"].numel()
    for row in g["derived"]:
        tp, tn, fp, fn = row["tp"], row["tn"], row["fp"], row["fn"]
        neutral = dict(dice=1, jaccard=1, mcc=0, balanced_accuracy=0)
        mine = dict(dice=b.metrics.dice_index(tp, fp, fn), jaccard=b.metrics.jaccard_index(tp, fp, fn),
                    mcc=b.metrics.mcc(tp, fp, fn, tn), balanced_accuracy=b.metrics.balanced_accuracy(tp, tn, fp, fn))
        for k, v in mine.items():
            v = neutral[k] if v == "invalid" else v
            assert float(v) == row[k], (k, tp, tn, fp, fn)


def test_robustness_vs_reference_goldens():
    b = _b()
    cols = dict(cosine=0, l2=2, pearson=4)
    for case in load("robustness"):
        q, k = case["q"].to(DEV), case["k"].to(DEV)
        d = b.robustness.all_distances(q.flatten(1), k.flatten(1)).cpu()
        for name, col in cols.items():
            ref = case["dist"][name]
            # BASELINE.json: rel <= 1e-4 in fp32 (plus the 1 - O(1) cancellation floor of the reference's own fp32)
            assert ((d[col] - ref).abs() <= 1e-4 * ref.abs() + 2e-6).all(), name
        for (name, margin), ref in case["scores"].items():
            got = b.robustness.Robustness(name, margin)(q, k).cpu()
            assert (got - ref).abs().max().item() <= 1e-4, (name, margin)


def _rms(a, r):
    return ((a - r).double().pow(2).mean().sqrt() / (r.double().pow(2).mean().sqrt() + 1e-30)).item()


@pytest.mark.parametrize("idx", range(3))
def test_unets_vs_reference_goldens(idx):
    """Converted U-Nets against the reference's own outputs on the same (seeded) weights and input, in the order the
    fixture was minted: one train-mode forward (updates the running statistics), Dice loss, then an eval-mode forward.
    Loss within 1 % (BASELINE.json).  Predictions: bf16 storage noise compounds over the 50-80 layers of these randomly
    weighted networks, so the yardstick is the fp32 oracle with that storage emulated (oracle/bf16_emulation.py) run
    here on the same weights — the converted model may be at most 1.5x as far from the reference (+1e-2)."""
    import copy
    from oracle import bf16_emulation
    b = _b()
    case = load("models")[idx]
    name = case["name"]
    if name == "basic_unet_binary":
        m = ref_models.basic_unet(out_ch=1, final_activation="sigmoid", in_channels=case["x"].shape[1])
    else:
        m = ref_models.resnet50_attention_unet(out_ch=4 if "4class" in name else 1,
                                               final_activation="softmax" if "4class" in name else "sigmoid")
    fill_state_(m, 100)
    emu = copy.deepcopy(m)
    bf16_emulation.emulate_bf16_storage(bf16_emulation.round_weights_(emu))
    gpu = b.convert(m.to(DEV))
    outs = {}
    for k, mod, x in (("emu", emu, case["x"]), ("gpu", gpu, case["x"].to(DEV))):
        mod.train()
        torch.manual_seed(3)
        yt = mod(x)
        if k == "gpu":
            loss = b.losses.DiceLoss()(yt, case["mask"].to(DEV))
        mod.eval()
        with torch.no_grad():
            outs[k] = (yt.detach().cpu(), mod(x).cpu())
    assert abs(loss.item() - case["loss"]) <= 1e-2 * abs(case["loss"]), f"{name}: loss {loss.item()} vs {case['loss']}"
    for which, gold in ((0, case["y_train"]), (1, case["y_eval"])):
        ctl, got = _rms(outs["emu"][which], gold), _rms(outs["gpu"][which], gold)
        print(f"{name} {'train' if which == 0 else 'eval'}: rms rel vs reference golden {got:.4f} (emulated-storage oracle {ctl:.4f})")
        assert got <= 1.5 * ctl + 1e-2, f"{name}: prediction rms rel {got:.4f} vs control {ctl:.4f}"
