"""The hot path at BASELINE.json's FULL sizes (cfg2 batch 256, cfg3 batch 24 at 256x256, cfg4 batch 4 at 1024x1024 with 5
channels, cfg5 50 000 rows), where the per-layer oracle comparisons of the other files would take minutes on the CPU:

* integer work (confusion counters, argmax histograms) — bit-exact against the CPU oracle, which is cheap at any size;
* losses — against the CPU oracle at full size (elementwise + one reduction: seconds);
* distances — size-independent properties: scaling both inputs by a power of two (exact in fp32) leaves cosine /
  Pearson unchanged and multiplies the mean-squared distance by its square (rel 1e-6); d(q, q) = 0; the fused negative pair equals
  the positive pair of an explicitly permuted input;
* convolutions / whole network — tile-decomposition independence: a slice of the batch computed inside the full batch
  equals the same images computed alone (whose parity with the oracle the small-size tests establish), zero input gives
  exactly zero, and the fused BatchNorm sums equal the sums of the stored output."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_losses, ref_metrics, ref_models

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _b200():
    import medsegpretrainimagenet_b200 as b
    return b


def _rel(got, ref):
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


def test_cfg4_multilabel_counters_and_bce_full_size():
    """cfg4: prediction (4, 5, 1024, 1024), float multilabel target with NaNs; counters bit-exact, BCE vs the oracle."""
    b = _b200()
    g = torch.Generator().manual_seed(70)
    shape = (4, 5, 1024, 1024)
    pred = torch.rand(shape, generator=g)
    pred.view(-1)[::13] = 0.5                                  # exactly on the threshold
    tgt = (torch.rand(shape, generator=g) < 0.05).float()
    tgt_nan = tgt.clone()
    tgt_nan.view(-1)[5::1001] = float("nan")
    for multilabel in (True, False):
        tp, tn, fp, fn, cc = ref_metrics.confusion_counts(pred, tgt_nan, 0.5, multilabel)
        got = b.metrics.binary_confusion_counts(pred.to(DEV), tgt_nan.to(DEV), 0.5, per_channel=multilabel)
        for name, ref in (("TP", tp), ("TN", tn), ("FP", fp), ("FN", fn), ("class_counts", cc)):
            assert np.array_equal(got[name].cpu().numpy(), ref), (name, multilabel)
    # conservation on clean targets: every element is in exactly one cell
    got = b.metrics.binary_confusion_counts(pred.to(DEV), tgt.to(DEV), 0.5, per_channel=True)
    tot = sum(got[k].cpu().numpy().astype(np.int64) for k in ("TP", "TN", "FP", "FN"))
    assert np.array_equal(tot, np.full(5, 4 * 1024 * 1024, dtype=np.int64))
    p = pred.clamp(1e-4, 1 - 1e-4).requires_grad_(True)
    l_ref = ref_losses.bce_loss_torch(p, tgt)
    l_ref.backward()
    pd = p.detach().to(DEV).requires_grad_(True)
    l = b.losses.BCELoss(torch_semantics=True)(pd, tgt.to(DEV))
    l.backward()
    assert abs(l.item() - l_ref.item()) <= 1e-5 * abs(l_ref.item())
    assert _rel(pd.grad.cpu(), p.grad) <= 1e-4


def test_cfg3_multiclass_confusion_and_dice_full_size():
    """cfg3: (24, 4, 256, 256) softmax prediction, int64 label map: C x C matrix bit-exact, Dice / CE vs the oracle."""
    b = _b200()
    g = torch.Generator().manual_seed(71)
    logits = torch.randn((24, 4, 256, 256), generator=g)
    logits[0] = torch.round(logits[0] * 2) / 2                 # exact ties: first maximum wins
    prob = torch.softmax(logits, 1)
    mask = torch.randint(0, 4, (24, 1, 256, 256), generator=g)
    ref = ref_metrics.multiclass_confusion_matrix(prob, mask.reshape(24, 256, 256), 4)
    got = b.metrics.multiclass_confusion_matrix(prob.to(DEV), mask.to(DEV)).cpu().numpy()
    assert np.array_equal(got, ref)
    assert int(got.sum()) == 24 * 256 * 256
    assert np.array_equal(got.sum(1).astype(np.int64), np.bincount(mask.flatten().numpy(), minlength=4))
    for batchwise in (True, False):
        p = prob.clone().requires_grad_(True)
        l_ref = ref_losses.dice_loss(p, mask, batchwise=batchwise, include_background=True)
        l_ref.backward()
        pd = prob.to(DEV).requires_grad_(True)
        l = b.losses.DiceLoss(batchwise=batchwise, include_background=True)(pd, mask.to(DEV))
        l.backward()
        assert abs(l.item() - l_ref.item()) <= 5e-6 * max(1.0, abs(l_ref.item())), batchwise
        assert _rel(pd.grad.cpu(), p.grad) <= 1e-4, batchwise
    p = prob.clone().requires_grad_(True)
    l_ref = ref_losses.ce_without_softmax(p, mask, 0.0)
    l_ref.backward()
    pd = prob.to(DEV).requires_grad_(True)
    l = b.losses.CrossEntropyLoss(0.0, apply_softmax=False)(pd, mask.to(DEV))
    l.backward()
    assert abs(l.item() - l_ref.item()) <= 1e-5 * abs(l_ref.item())
    assert _rel(pd.grad.cpu(), p.grad) <= 1e-4


def test_cfg5_robustness_properties_50k_rows():
    """cfg5 level-5 pooled size: N = 50 000 rows, D = 2048."""
    b = _b200()
    n, d = 50000, 2048
    g = torch.Generator(device=DEV).manual_seed(72)
    q = torch.relu(torch.randn((n, d), device=DEV, generator=g))
    k = torch.relu(q + 0.1 * torch.randn((n, d), device=DEV, generator=g))
    base = b.robustness.all_distances(q, k)                     # (6, N): cos+, cos-, l2+, l2-, pearson+, pearson-
    assert base.shape == (6, n) and torch.isfinite(base).all()
    # power-of-two scaling is exact in fp32: cosine and Pearson unchanged, mean-squared distance x 16 (rel 1e-6)
    sc = b.robustness.all_distances(q * 4.0, k * 4.0)

    def same(a, c):
        return ((a - c).abs() <= 1e-6 * c.abs() + 1e-9).all().item()

    assert same(sc[0], base[0]) and same(sc[1], base[1])
    assert same(sc[4], base[4]) and same(sc[5], base[5])
    assert same(sc[2], base[2] * 16.0) and same(sc[3], base[3] * 16.0)
    # the fused negative pair is the positive pair of the explicitly permuted input
    perm = torch.tensor(b.robustness.negative_permutation(n), device=DEV)
    expl = b.robustness.all_distances(q, k[perm])
    for i in (0, 2, 4):
        assert (expl[i] - base[i + 1]).abs().max().item() <= 1e-6 * max(1.0, base[i + 1].abs().max().item()), i
    assert torch.equal(perm[perm], torch.arange(n, device=DEV))                    # involution (eval.py:22-23)
    # identical inputs: zero distance
    ident = b.robustness.all_distances(q, q)
    assert ident[2].abs().max().item() <= 1e-5
    assert ident[0].abs().max().item() <= 1e-6 and ident[4].abs().max().item() <= 1e-5
    # against the reference formulas in fp64 on a sample of rows (rel <= 1e-4, BASELINE.json)
    idx = torch.arange(0, n, 997, device=DEV)
    qs, ks = q[idx].double(), k[idx].double()
    cos = 1 - (qs * ks).sum(1) / ((qs * qs).sum(1) * (ks * ks).sum(1)).sqrt()
    l2 = ((qs - ks) ** 2).mean(1)
    qc, kc = qs - qs.mean(1, keepdim=True), ks - ks.mean(1, keepdim=True)
    pear = 1 - (qc * kc).sum(1) / ((qc * qc).sum(1) * (kc * kc).sum(1)).sqrt()
    for col, exact in ((0, cos), (2, l2), (4, pear)):
        assert ((base[col][idx].double() - exact).abs() <= 1e-4 * exact.abs() + 1e-7).all(), col


def test_cfg2_conv_layer_full_batch_properties():
    """A ResNet-50 3x3 layer and a 1x1 expand layer at the bench's batch of 256: a batch slice inside the full launch
    equals the slice launched alone (same per-output accumulation order whatever the tile walk), zero in ->
    zero out, fused BatchNorm sums == sums of the stored bf16 output."""
    from medsegpretrainimagenet_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(73)
    for (hw, ci, co, k, pad) in ((14, 256, 256, 3, 1), (14, 256, 1024, 1, 0)):
        x = torch.randn((256, hw, hw, ci), device=DEV, generator=g).to(torch.bfloat16)
        w = torch.randn((co, ci, k, k), device=DEV, generator=g) * 0.05
        wf, wd = ops.pack_weights(w)
        ho, wo, pt, pl = ops.conv_out_size(hw, hw, k, k, 1, pad)
        stats = torch.zeros((2, co), device=DEV)
        y = ops.conv_fprop(x, wf, None, co, k, k, 1, pt, pl, ho, wo, stats=stats)
        ys = ops.conv_fprop(x[64:72].contiguous(), wf, None, co, k, k, 1, pt, pl, ho, wo)
        # same per-output accumulation whatever the tile walk: at most a bf16 rounding flip (2^-8 of the value)
        assert _rel(y[64:72].float(), ys.float()) <= 8e-3, (hw, ci, co, k)
        assert (y[64:72] == ys).float().mean().item() >= 0.999
        yf = y.float().reshape(-1, co)
        assert _rel(stats[0], yf.sum(0)) <= 1e-3 and _rel(stats[1], (yf * yf).sum(0)) <= 1e-3
        z = ops.conv_fprop(torch.zeros_like(x), wf, None, co, k, k, 1, pt, pl, ho, wo)
        assert z.abs().max().item() == 0.0
        dy = torch.randn((256, ho, wo, co), device=DEV, generator=g).to(torch.bfloat16)
        dx = ops.conv_dgrad(dy, wd, tuple(x.shape), k, k, 1, pt, pl)
        dxs = ops.conv_dgrad(dy[64:72].contiguous(), wd, (8, hw, hw, ci), k, k, 1, pt, pl)
        assert _rel(dx[64:72].float(), dxs.float()) <= 8e-3, (hw, ci, co, k)
        assert (dx[64:72] == dxs).float().mean().item() >= 0.999
        # wgrad is a sum over the batch: full = sum of two halves (fp32 partials, different split-K: 1e-3 of the range)
        dw = ops.conv_wgrad(x, dy, ci, k, k, 1, pt, pl)
        dwa = ops.conv_wgrad(x[:128].contiguous(), dy[:128].contiguous(), ci, k, k, 1, pt, pl)
        dwb = ops.conv_wgrad(x[128:].contiguous(), dy[128:].contiguous(), ci, k, k, 1, pt, pl)
        assert _rel(dw, dwa + dwb) <= 1e-3


def test_cfg2_resnet50_eval_batch_256_equals_batch_8():
    """Whole network at the bench's size: in eval mode (running statistics) the images are independent, so the first 8
    of 256 must give the logits of those 8 run alone — the run whose parity with the oracle test_hotpath_gpu checks."""
    b = _b200()
    torch.manual_seed(0)
    ref = ref_models.kaiming_init_(ref_models.resnet50_classifier(num_classes=1000))
    gpu = b.convert(copy.deepcopy(ref).to(DEV)).eval()
    g = torch.Generator().manual_seed(74)
    x = torch.randn((256, 3, 224, 224), generator=g).to(DEV)
    with torch.no_grad():
        y_full = gpu(x)
        y_8 = gpu(x[:8].contiguous())
    assert y_full.shape == (256, 1000) and torch.isfinite(y_full).all()
    assert _rel(y_full[:8], y_8) <= 2e-2      # (the kernel variant / tile shape may differ between the two batch sizes)
    # and the bf16 image batch of BASELINE cfg2 is accepted as is
    with torch.no_grad():
        y_bf = gpu(x[:8].to(torch.bfloat16))
    assert _rel(y_bf, gpu(x[:8].to(torch.bfloat16).float())) == 0.0


def test_cfg5_unpooled_levels_streamed_equals_resident():
    """cfg5: unpooled level-5 representations (D = 100 352 = 2048 x 7 x 7) resident, scored (a) in one launch
    and (b) streamed in permutation-closed chunks {i, perm(i)} as the 160 GB levels 1-3 have to be: identical scores.
    Also level 4 (D = 200 704) through the pooled-in-kernel path against pooling first."""
    import medsegpretrainimagenet_b200 as b
    from medsegpretrainimagenet_b200 import robustness as R
    g = torch.Generator(device=DEV).manual_seed(0)
    n, d = 4001, 100352                                      # odd n: the permutation has a fixed point
    q = torch.relu(torch.randn((n, d), device=DEV, generator=g))
    k = torch.relu(q + 0.1 * torch.randn((n, d), device=DEV, generator=g))
    whole = R.robustness_table(q, k)
    streamed = R.robustness_table_streamed(lambda lo, hi: (q[lo:hi], k[lo:hi]), n, rows_per_chunk=514)
    assert torch.equal(whole, streamed)
    chunks = R.symmetric_chunks(n, 514)
    assert len(chunks) >= 8 and all(sum(hi - lo for lo, hi in c) <= 516 for c in chunks)
    del q, k
    # level 4 unpooled maps (1024 x 14 x 14): pool=True inside the kernel == spatial mean first, rel <= 1e-4
    f0 = torch.relu(torch.randn((600, 1024, 14, 14), device=DEV, generator=g))
    f1 = torch.relu(f0 + 0.1 * torch.randn((600, 1024, 14, 14), device=DEV, generator=g))
    fused = R.all_distances(f0, f1, pool=True)
    first = R.all_distances(f0.flatten(2).mean(2), f1.flatten(2).mean(2))
    assert ((fused - first).abs().max() / first.abs().max()).item() <= 1e-4
    # and the unpooled level-4 row length through the plain path: d(q, q) = 0 for all three distances
    same = R.all_distances(f0, f0)
    assert same[0].abs().max().item() <= 1e-6 and same[2].abs().max().item() == 0.0 and same[4].abs().max().item() <= 1e-6


def test_predict_w_model_levels_and_pooling():
    """robustness/eval.py:30-54 (with its missing torch.cat): every level of a ResNet-50 encoder, pooled and unpooled,
    batched; D of the unpooled levels as SURVEY a19 lists them for 224 x 224 inputs."""
    from medsegpretrainimagenet_b200 import models, robustness as R
    torch.manual_seed(0)
    enc = models.kaiming_init_(models.DeepResNet(bias=False)).to(DEV).eval()
    imgs = torch.rand((5, 3, 224, 224))
    want_d = [802816, 802816, 401408, 200704, 100352]
    want_c = [64, 256, 512, 1024, 2048]
    for level in range(5):
        un = R.predict_w_model(enc, imgs, batch_size=2, device=DEV, level=level, pool=False)
        po = R.predict_w_model(enc, imgs, batch_size=2, device=DEV, level=level, pool=True)
        assert un.shape[0] == 5 and un[0].numel() == want_d[level] and tuple(po.shape) == (5, want_c[level])
        assert torch.allclose(po, un.flatten(2).mean(2), rtol=1e-5, atol=1e-6)
    # batching does not change the representation (eval mode: running statistics)
    a = R.predict_w_model(enc, imgs, batch_size=5, device=DEV, level=-1, pool=True)
    b2 = R.predict_w_model(enc, imgs, batch_size=1, device=DEV, level=-1, pool=True)
    assert torch.allclose(a, b2, rtol=2e-2, atol=1e-3)
