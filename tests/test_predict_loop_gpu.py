"""The reference's per-batch call pattern (train_model.py:51-120) on the B200 path, over reference-structured objects:

    optimizer.zero_grad() -> model(**batch) -> MetricsCalculator.calculate_batch(batch, train=, accumulation_scale=)
      [metric DAG, parents first; then Loss.calculate_batch: criterion / scale -> .item() -> .backward()]
    -> evaluate_batch on the last fragment -> clip_grad_norm_ -> optimizer.step()

run side by side with the oracle (the reference's algorithms in fp32 on the CPU) on the same weights and batches, with
gradient accumulation (two fragments per optimizer step) and an evaluation pass.  With the reference checkout present the
same loop is `train_model.predict` itself after `patch.install()` (tests/test_patch_cpu.py checks that wiring); here
the host-side mirrors (`losses.Loss`, `metrics.MetricsCalculator`) drive the identical sequence of kernel calls."""
import copy

import numpy as np
import pytest
import torch

from oracle import bf16_emulation, ref_losses, ref_metrics, ref_models

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _predict_b200(model, batches, mc, optimizer, accumulation_scale, train, b200):
    """train_model.predict (train_model.py:16-130) with its logging side removed."""
    model.train(train)
    logs = []
    for i, batch in enumerate(batches):
        step = (i + 1) % accumulation_scale == 0 or i == len(batches) - 1
        if train and (i % accumulation_scale == 0):      # gradients of the fragments of one virtual batch accumulate
            optimizer.zero_grad()
        batch = {k: v.to(DEV) for k, v in batch.items()}
        if train:
            batch["prediction"] = model(batch["x"])
        else:
            with torch.no_grad():
                batch["prediction"] = model(batch["x"])
        values = mc.calculate_batch(batch, train=train, accumulation_scale=accumulation_scale, last=False)
        if step:
            values = mc.evaluate_batch(batch, train=train, accumulation_scale=accumulation_scale, last=False)
            if train:
                values["gradient_magnitude"] = b200.optim.clip_grad_norm_(list(model.parameters()), float("inf"), 2.0).item()
                optimizer.step()
            logs.append(values)
    return logs, mc.evaluate_epoch()


def _predict_oracle(model, batches, optimizer, accumulation_scale, train, loss_fn):
    model.train(train)
    logs, tot = [], np.zeros(4, dtype=np.int64)
    acc_counts, acc_loss = np.zeros(4, dtype=np.int64), 0.0
    epoch_losses = []
    for i, batch in enumerate(batches):
        step = (i + 1) % accumulation_scale == 0 or i == len(batches) - 1
        if train and (i % accumulation_scale == 0):
            optimizer.zero_grad()
        with torch.set_grad_enabled(train):
            pred = model(batch["x"])
        tp, tn, fp, fn, _ = ref_metrics.confusion_counts(pred.detach(), batch["mask"], 0.5)
        acc_counts += np.array([int(tp), int(tn), int(fp), int(fn)])
        loss = loss_fn(pred, batch["mask"]) / accumulation_scale
        acc_loss += loss.item()
        if train:
            loss.backward()
        if step:
            tp, tn, fp, fn = (int(v) for v in acc_counts)
            values = {"metrics/dice_index_threshold_0.5": ref_metrics.dice_index(tp, fp, fn),
                      "metrics/mcc_threshold_0.5": ref_metrics.mcc(tp, fp, fn, tn),
                      "metrics/balanced_accuracy_threshold_0.5": ref_metrics.balanced_accuracy(tp, tn, fp, fn),
                      "dice_loss": acc_loss, "counts": acc_counts.copy()}
            tot += acc_counts
            epoch_losses.append(acc_loss)
            acc_counts[:] = 0
            acc_loss = 0.0
            if train:
                values["gradient_magnitude"] = torch.nn.utils.clip_grad_norm_(model.parameters(), float("inf"), 2.0).item()
                optimizer.step()
            logs.append(values)
    tp, tn, fp, fn = (int(v) for v in tot)
    return logs, {"metrics/dice_index_threshold_0.5": ref_metrics.dice_index(tp, fp, fn),
                  "dice_loss": float(np.mean(epoch_losses))}


@pytest.mark.parametrize("train", [True, False])
def test_predict_shaped_loop_with_accumulation_matches_the_oracle(train):
    import medsegpretrainimagenet_b200 as b200
    from medsegpretrainimagenet_b200 import losses, metrics as M
    torch.manual_seed(0)
    ref = bf16_emulation.round_weights_(ref_models.kaiming_init_(ref_models.resnet18_attention_unet()))
    gpu = b200.convert(copy.deepcopy(ref).to(DEV))
    g = torch.Generator().manual_seed(3)
    batches = [{"x": torch.rand((4, 1, 64, 64), generator=g).to(torch.bfloat16).float(),
                "mask": (torch.rand((4, 1, 64, 64), generator=g) < 0.3).long()} for _ in range(4)]
    acc = 2
    loss_wrapper = losses.Loss(losses.DiceLoss(), label_type="mask")
    assert loss_wrapper.name == "dice_loss"
    mc = M.MetricsCalculator([M.DiceIndex, M.MCC, M.BalancedAccuracy], loss=loss_wrapper, thresholds=(0.5,))
    opt_g = b200.optim.SGD(list(gpu.parameters()), lr=0.01, momentum=0.9, weight_decay=1e-4)
    opt_r = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    logs_g, epoch_g = _predict_b200(gpu, batches, mc, opt_g, acc, train, b200)
    logs_r, epoch_r = _predict_oracle(ref, batches, opt_r, acc, train, ref_losses.dice_loss)
    assert len(logs_g) == len(logs_r) == 2
    for lg, lr_ in zip(logs_g, logs_r):
        # only int / float values reach the logs (metric_wrapper.py:281); keys as the reference names them
        assert all(isinstance(v, (int, float)) for v in lg.values())
        assert {"metrics/dice_index_threshold_0.5", "metrics/mcc_threshold_0.5",
                "metrics/balanced_accuracy_threshold_0.5", "dice_loss"} <= set(lg)
        assert abs(lg["dice_loss"] - lr_["dice_loss"]) <= 1e-2 * abs(lr_["dice_loss"])          # loss within 1 %
        for k in ("metrics/dice_index_threshold_0.5", "metrics/mcc_threshold_0.5",
                  "metrics/balanced_accuracy_threshold_0.5"):
            assert abs(lg[k] - lr_[k]) <= 2e-2, (k, lg[k], lr_[k])       # bf16 predictions flip a few threshold pixels
        if train:
            assert abs(lg["gradient_magnitude"] - lr_["gradient_magnitude"]) <= 5e-2 * lr_["gradient_magnitude"]
    assert abs(epoch_g["dice_loss"] - epoch_r["dice_loss"]) <= 1e-2 * abs(epoch_r["dice_loss"])
    assert abs(epoch_g["metrics/dice_index_threshold_0.5"] - epoch_r["metrics/dice_index_threshold_0.5"]) <= 2e-2
    if train:
        # two optimizer steps later the weights still agree (the update went through clip + SGD on accumulated gradients)
        for (n, pg), pr in zip(gpu.named_parameters(), ref.parameters()):
            d = (pg.detach().cpu() - pr.detach()).abs().max().item()
            assert d <= 2e-2 * max(pr.detach().abs().max().item(), 1e-3) + 2e-4, (n, d)


def test_metric_counts_in_the_loop_are_bit_exact_on_identical_predictions():
    """Same prediction tensor on both sides -> the accumulated counters and every derived metric are EXACTLY the oracle's."""
    from medsegpretrainimagenet_b200 import metrics as M
    g = torch.Generator().manual_seed(5)
    mc = M.MetricsCalculator([M.DiceIndex, M.JaccardIndex, M.MCC, M.Accuracy, M.Sensitivity, M.Specificity, M.Precision])
    tot = np.zeros(4, dtype=np.int64)
    for _ in range(3):
        pred = torch.rand((2, 1, 96, 96), generator=g)
        mask = (torch.rand((2, 1, 96, 96), generator=g) < 0.4).long()
        mc.calculate_batch({"prediction": pred.to(DEV), "mask": mask.to(DEV)})
        tp, tn, fp, fn, _ = ref_metrics.confusion_counts(pred, mask, 0.5)
        tot += np.array([int(tp), int(tn), int(fp), int(fn)])
    out = mc.evaluate_batch({})
    tp, tn, fp, fn = (int(v) for v in tot)
    assert out["metrics/dice_index_threshold_0.5"] == ref_metrics.dice_index(tp, fp, fn)
    assert out["metrics/jaccard_index_threshold_0.5"] == ref_metrics.jaccard_index(tp, fp, fn)
    assert out["metrics/mcc_threshold_0.5"] == ref_metrics.mcc(tp, fp, fn, tn)
    assert out["metrics/accuracy_threshold_0.5"] == ref_metrics.accuracy(tp, fp, tn, fn)
    assert out["metrics/sensitivity_threshold_0.5"] == tp / (tp + fn)
    assert out["metrics/specificity_threshold_0.5"] == tn / (tn + fp)
    assert out["metrics/precision_threshold_0.5"] == tp / (tp + fp)


def test_multiclass_dag_on_the_device_counters():
    from medsegpretrainimagenet_b200 import metrics as M
    g = torch.Generator().manual_seed(6)
    pred = torch.softmax(torch.randn((3, 4, 40, 40), generator=g), 1)
    mask = torch.randint(0, 4, (3, 1, 40, 40), generator=g)
    mc = M.MetricsCalculator([M.MeanDiceIndex, M.MeanJaccardIndex, M.MultiClassAccuracy], number_of_classes=4,
                             include_background_in_averages=True)
    mc.calculate_batch({"prediction": pred.to(DEV), "mask": mask.to(DEV)})
    out = mc.evaluate_batch({})
    cm = ref_metrics.multiclass_confusion_matrix(pred, mask, 4)
    want = ref_metrics.mean_over_present_classes(cm, lambda tp, fp, fn: ref_metrics.dice_index(tp, fp, fn),
                                                 include_background=True)
    assert out["metrics/mean_dice_index"] == pytest.approx(float(want), abs=1e-12)
    assert out["metrics/accuracy"] == pytest.approx(float(np.diagonal(cm).sum() / cm.sum()), abs=1e-12)
