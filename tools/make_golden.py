"""Mint the golden fixtures tests/golden/*.pt from the REAL reference (imported from /root/reference/src by
oracle/reference_harness.py; this script only runs in the build container).  Every fixture holds seeded
inputs and the outputs of the reference's own classes on them:

  losses.pt      DiceLoss / CrossEntropyLoss / BCELoss values and input gradients
                 (segmentation/losses/losses.py:34-58, classification/losses.py:4-40)
  metrics.pt     ConfusionMatrix, MultiClassConfusionMatrix, Top5Accuracy counts and the derived
                 Dice / Jaccard / MCC / balanced accuracy / accuracy values (metrics/metrics.py, multiclass_metrics.py)
  robustness.pt  the three distances and Robustness scores (robustness/distance.py, robustness/eval.py:16-28)
  models.pt      forward output, loss and parameter-gradient norms of the reference's U-Nets / DeepResNet built
                 from its YAML configs with seeded initialisation (no weights stored: oracle/seeded_weights.py
                 reproduces them on both sides; a digest of the state_dict pins that)

    python tools/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_harness as H  # noqa: E402
from oracle.seeded_weights import fill_state_  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def gen(seed):
    torch.manual_seed(seed)
    return torch.Generator().manual_seed(seed)


def state_digest(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def losses():
    from segmentation.losses.losses import DiceLoss
    from classification.losses import BCELoss, CrossEntropyLoss
    g = gen(7)
    cases = []
    for c, act in ((1, "sigmoid"), (4, "softmax"), (5, "sigmoid")):
        logits = torch.randn((3, c, 20, 24), generator=g)
        p = torch.sigmoid(logits) if act == "sigmoid" else torch.softmax(logits, 1)
        mask = torch.randint(0, max(c, 2), (3, 1, 20, 24), generator=g)
        for batchwise in (True, False):
            for bg in (True, False):
                if c == 1 and not bg:
                    continue  # the reference mutates `classes_start` on that branch (losses.py:51)
                pp = p.clone().requires_grad_(True)
                loss = DiceLoss(batchwise=batchwise, include_background=bg)(pp, mask)
                loss.backward()
                cases.append(dict(kind="dice", pred=p, target=mask, batchwise=batchwise, include_background=bg,
                                  loss=loss.detach(), grad=pp.grad.clone()))
    logits = torch.randn((6, 10), generator=g)
    lab = torch.randint(0, 10, (6, 1), generator=g)
    for s in (0.0, 0.1):
        ll = logits.clone().requires_grad_(True)
        loss = CrossEntropyLoss(label_smoothing=s)(ll, lab)
        loss.backward()
        cases.append(dict(kind="ce_softmax", pred=logits, target=lab, smoothing=s, loss=loss.detach(),
                          grad=ll.grad.clone()))
    p4 = torch.softmax(torch.randn((2, 4, 9, 9), generator=g), 1)
    lab4 = torch.randint(0, 4, (2, 1, 9, 9), generator=g)
    for s in (0.0, 0.2):
        pp = p4.clone().requires_grad_(True)
        loss = CrossEntropyLoss(label_smoothing=s, apply_softmax=False)(pp, lab4)
        loss.backward()
        cases.append(dict(kind="ce_prob", pred=p4, target=lab4, smoothing=s, loss=loss.detach(),
                          grad=pp.grad.clone()))
    p1 = torch.sigmoid(torch.randn((3, 2, 10, 12), generator=g))
    t = torch.randint(0, 2, (3, 2, 10, 12), generator=g).float()
    pp = p1.clone().requires_grad_(True)
    loss = BCELoss()(pp, t)
    loss.backward()
    cases.append(dict(kind="bce_plain", pred=p1, target=t, loss=loss.detach(), grad=pp.grad.clone()))
    pp = p1.clone().requires_grad_(True)
    loss = torch.nn.BCELoss()(pp, t)
    loss.backward()
    cases.append(dict(kind="bce_torch", pred=p1, target=t, loss=loss.detach(), grad=pp.grad.clone()))
    return cases


def metrics():
    from metrics.metrics import ConfusionMatrix
    from metrics import metrics as rm
    from metrics.multiclass_metrics import MultiClassConfusionMatrix, Top5Accuracy
    from utils.config_dict import ConfigDict
    g = gen(11)
    out = dict(binary=[], multiclass=[], top5=[], derived=[])
    for multilabel, shape, thr in ((False, (3, 1, 16, 18), 0.5), (True, (3, 5, 16, 18), 0.5), (False, (2, 1, 33, 7), 0.3),
                                   (True, (1, 3, 5, 5), 0.7)):
        cd = ConfigDict({"metrics": {"calculation": {"multilabel": multilabel, "ignore_nans": True}}})
        cm = ConfusionMatrix(cd, threshold=thr)
        pred = torch.rand(shape, generator=g)
        pred.view(-1)[::5] = thr  # exactly on the threshold: `>=`
        tgt = torch.randint(0, 2, shape, generator=g).float()
        tgt.view(-1)[3::17] = float("nan")
        res = cm.calculate_batch(pred, mask=tgt)
        out["binary"].append(dict(
            pred=pred, target=tgt, threshold=thr, multilabel=multilabel,
            **{k: torch.as_tensor(np.asarray(res[f"{k}_threshold_{thr}"])).to(torch.int64)
               for k in ("true_positives", "true_negatives", "false_positives", "false_negatives")},
            class_counts=torch.as_tensor(np.asarray(cm.class_counts)).to(torch.int64)))
    for n, c, hw in ((2, 4, (12, 12)), (3, 2, (5, 7)), (2, 20, (9, 9))):
        cd = ConfigDict({"metrics": {"calculation": {"number_of_classes": c}}})
        pred = torch.randn((n, c, *hw), generator=g)
        pred[0] = torch.round(pred[0] * 2) / 2  # exact ties: argmax takes the first index
        tgt = torch.randint(0, c, (n, 1, *hw), generator=g)
        cmx = MultiClassConfusionMatrix(_config_dict=cd).calculate_batch(pred, mask=tgt)["confusion_matrix"]
        out["multiclass"].append(dict(pred=pred, target=tgt, classes=c,
                                      confusion_matrix=torch.as_tensor(np.asarray(cmx)).to(torch.int64)))
    logits = torch.randn((64, 50), generator=g)
    logits[:8] = torch.round(logits[:8])
    lab = torch.randint(0, 50, (64, 1), generator=g)
    frac = Top5Accuracy().calculate_batch(logits, label=lab)["top_5_accuracy"]
    out["top5"].append(dict(pred=logits, target=lab, hits=int(round(float(frac) * 64))))
    rng = np.random.default_rng(0)
    cases = [tuple(int(v) for v in rng.integers(0, 50, 4)) for _ in range(24)]
    cases += [(0, 0, 0, 0), (0, 5, 0, 0), (3, 0, 0, 0), (0, 0, 7, 0), (0, 0, 0, 9), (4, 4, 0, 0)]
    for tp, tn, fp, fn in cases:
        pv = dict(true_positives=torch.tensor(tp), true_negatives=torch.tensor(tn),
                  false_positives=torch.tensor(fp), false_negatives=torch.tensor(fn))
        row = dict(tp=tp, tn=tn, fp=fp, fn=fn)
        for name, cls in (("dice", rm.DiceIndex), ("jaccard", rm.JaccardIndex), ("mcc", rm.MCC),
                          ("balanced_accuracy", rm.BalancedAccuracy)):
            v = list(cls().evaluate_batch(pv).values())[0]
            row[name] = float(v)
        if tp + tn + fp + fn:
            row["accuracy"] = float(list(rm.Accuracy().evaluate_batch(pv).values())[0])
        out["derived"].append(row)
    return out


def robustness():
    from robustness import distance as rd
    from robustness.eval import Robustness
    g = gen(17)
    cases = []
    for n, shape in ((2, (6, 5, 5)), (3, (64,)), (8, (16, 3, 3)), (9, (500,))):
        q = torch.relu(torch.randn((n, *shape), generator=g))
        k = torch.relu(q + 0.1 * torch.randn((n, *shape), generator=g))
        case = dict(q=q, k=k, dist={}, scores={})
        for name, fn in (("cosine", rd.cosine_distance), ("l2", rd.l2_loss), ("pearson", rd.inv_pearson_corr)):
            case["dist"][name] = fn(q.flatten(1), k.flatten(1))
            for margin in (0.0, 0.25, 0.5):
                case["scores"][(name, margin)] = Robustness(fn, margin)(q, k)
        cases.append(case)
    return cases


def models():
    from segmentation.losses.losses import DiceLoss
    out = []
    specs = [
        ("r50_attention_unet_binary", "downstream/acdc/resnet50_attention_unet.yaml", {}, (2, 3, 64, 64), 2),
        ("r50_attention_unet_4class", "downstream/acdc/resnet50_attention_unet.yaml",
         {"model/segmentation.models.UNet/architecture/out_channel_size": 4,
          "model/segmentation.models.UNet/architecture/activation_function/final": "softmax"}, (1, 3, 64, 64), 4),
        ("basic_unet_binary", "downstream/covidqu/unet.yaml", {}, None, 2),
    ]
    for name, yaml_path, overrides, xshape, classes in specs:
        cd = H.load_config(yaml_path, overrides=overrides)
        ref = H.build_model(cd, seed=0)
        fill_state_(ref, seed=100)      # identical weights on both sides, see oracle/seeded_weights.py
        if xshape is None:
            xshape = (1, ref.model.encoder.first_block.model.in_channels, 32, 32)
        g = gen(21)
        x = torch.rand(xshape, generator=g)
        mask = torch.randint(0, classes, (xshape[0], 1, *xshape[2:]), generator=g)
        digest = state_digest(ref.state_dict())
        ref.train()
        torch.manual_seed(3)
        y = ref(x)
        loss = DiceLoss()(y, mask)
        loss.backward()
        gn = {k.replace(".model.", "."): float(p.grad.norm()) for k, p in ref.model.named_parameters()
              if p.grad is not None}
        ref.eval()
        with torch.no_grad():
            y_eval = ref(x)
        out.append(dict(name=name, yaml=yaml_path, x=x, mask=mask, state_digest=digest, y_train=y.detach(),
                        loss=float(loss), grad_norms=gn, y_eval=y_eval))
    # the DeepResNet encoder family directly (classification/models.py:9-103): ResNet-18 style, skips + head
    from classification import models as ref_cls
    torch.manual_seed(0)
    enc = ref_cls.DeepResNet(bottleneck=False, channel_sizes=(64, 128, 256, 512), widths=(2, 2, 2, 2),
                             in_channels=1, bias=False, head=True, output_size=10, stochastic_depth_rate=0.2)
    fill_state_(enc, seed=101)
    enc_digest = state_digest(enc.state_dict())
    x = torch.rand((3, 1, 64, 64), generator=gen(5))
    enc.train()
    torch.manual_seed(9)
    yt, st = enc(x, return_skip_vals=True)
    enc.eval()
    with torch.no_grad():
        ye, se = enc(x, return_skip_vals=True)
    out.append(dict(name="deepresnet18_head10", x=x, state_digest=enc_digest,
                    y_train=yt.detach(), skip_means_train=[float(s.mean()) for s in st], y_eval=ye,
                    skip_means_eval=[float(s.mean()) for s in se]))
    return out


def main():
    assert H.available(), "the reference is not mounted at /root/reference"
    H.setup()
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("losses", losses), ("metrics", metrics), ("robustness", robustness), ("models", models)):
        obj = fn()
        path = os.path.join(OUT, name + ".pt")
        torch.save(obj, path)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
