"""N-rank == 1-rank parity of the data-parallel step, runnable wherever N GPUs are (SURVEY.md §4 tier 6).

The reference's `nn.DataParallel` (train_model.py:192-197) normalises every replica's shard with its own BatchNorm
statistics; this path synchronises the statistics (SyncBN through the peer-memory kernel or NCCL), reduces batchwise Dice
over GLOBAL sums and averages the gradients, so that N ranks on N shards reproduce the SINGLE-DEVICE step on the
concatenated batch.  Two levels, because a randomly initialised BatchNorm network in bf16 is chaotic (a 1e-7 change of
one statistic — a different fp32 summation order is enough — flips bf16 roundings, and the perturbation grows ~2.7x
per ResNet level: tests/test_hotpath_gpu.py `_check` measures that against the fp32 oracle):

`exchange_parity()` — the three exchanges TEACHER-FORCED (same inputs on both sides, one layer deep, so nothing can
amplify), tight bounds:
    conv -> SyncBN -> ReLU    y, dx rms rel <= 1e-4 (isolated bf16 roundings); running stats rel <= 1e-5;
                              world x averaged (dW, dgamma, dbeta) vs the full-batch gradients rel <= 1e-4
    batchwise Dice            loss rel <= 1e-6, the shard's dL/dp rel <= 1e-5 of the full-batch rows
    GradReducer               the buckets hold mean over ranks of the local gradients (the dW / dgamma line above)

`n_rank_parity()` — one whole step of a ResNet-18 attention U-Net: every rank also runs the full-batch step locally
(same seeds, no process group) and compares
    loss                      relative difference <= 1e-3
    parameter gradients       cosine >= 0.999 (all parameters flattened), norm within 1 % — OR no further from the
                              single-device step than that step is from ITSELF with another summation order
                              (`noise_floor_cosine`: deterministic rows vs atomics, same device, same data; measured
                              in the same call): cosine >= floor - 0.08, norm within 2 %
    BatchNorm running stats   relative difference <= 1e-3 (max-norm over each buffer), or <= 3x the floor's
    confusion counters        all-reduced per-rank counters == counters of the gathered predictions, BIT-EXACT
                              (and within 0.2 % of the pixels of the single-device step's counters)

Used by tests/test_multigpu_gpu.py (under torchrun when >= 2 GPUs are visible) and as the pre-check of
`bench.py --gpus N` (`"parity_n"` in its JSON line): the driver's GPU test box has one GPU."""
from __future__ import annotations

import torch
import torch.distributed as dist

import os

from . import functional as Fn
from . import losses, metrics, models, ops
from .parallel import GradReducer, shard_rows


def _flat(ts):
    return torch.cat([t.reshape(-1).double() for t in ts])


@torch.no_grad()
def _bn_buffers(model):
    return [b for n, b in model.named_buffers() if n.endswith("running_mean") or n.endswith("running_var")]


def n_rank_parity(group=None, device=None, per_rank_batch: int = 4, size: int = 128, seed: int = 0) -> dict:
    """Runs in deterministic mode (torch.use_deterministic_algorithms, as the reference's downstream YAMLs do): the
    reductions then have a fixed order, so a difference between the two steps is a property of the data-parallel
    path (statistics added per rank first, then across ranks), not run-to-run noise — `repeat_grad_cosine` (the
    single-device step executed twice) documents that."""
    was = torch.are_deterministic_algorithms_enabled()
    warn = torch.is_deterministic_algorithms_warn_only_enabled()
    torch.use_deterministic_algorithms(True, warn_only=True)
    try:
        return _n_rank_parity(group, device, per_rank_batch, size, seed)
    finally:
        torch.use_deterministic_algorithms(was, warn_only=warn)


def _n_rank_parity(group, device, per_rank_batch, size, seed) -> dict:
    world = dist.get_world_size(group) if (group is not None and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    g = torch.Generator().manual_seed(seed + 17)
    nb = per_rank_batch * world
    if os.environ.get("MSP_SELFCHECK_DATA", "blobs") == "rand":
        x = torch.rand((nb, 1, size, size), generator=g)
        y = (torch.rand((nb, 1, size, size), generator=g) < 0.3).long()
    else:
        coarse = torch.rand((nb, 1, size // 8, size // 8), generator=g)
        x = torch.nn.functional.interpolate(coarse, size=(size, size), mode="bilinear", align_corners=False)
        x = (x + 0.05 * torch.rand((nb, 1, size, size), generator=g)).clamp_(0, 1)
        y = (x > 0.55).long()
    lo, hi = shard_rows(nb, rank, world)

    def build(grp):
        torch.manual_seed(seed)
        m = models.resnet18_attention_unet(group=grp)
        models.kaiming_init_(m)
        return m.to(device).train()

    # ---- N ranks, one shard each -------------------------------------------------------------------------------
    m_dist = build(group if world > 1 else None)
    params_d = [p for p in m_dist.parameters() if p.requires_grad]
    reducer = GradReducer(params_d, bucket_mb=8.0, group=group if world > 1 else None)
    reducer.zero_grad()
    pred_d = m_dist(x[lo:hi].to(device))
    loss_d = losses.DiceLoss(group=group if world > 1 else None)(pred_d, y[lo:hi].to(device))
    loss_d.backward()
    reducer.finish()
    cnt = metrics.binary_confusion_counts(pred_d.detach(), y[lo:hi].to(device), 0.5)
    counts_d = torch.stack([cnt[k] for k in ("TP", "TN", "FP", "FN")]).clone()
    if world > 1:
        dist.all_reduce(counts_d, group=group)                      # int64 sums are associative: bit-exact
        gathered = [torch.empty_like(pred_d) for _ in range(world)]
        dist.all_gather(gathered, pred_d.detach().contiguous(), group=group)
        pred_all = torch.cat(gathered)
    else:
        pred_all = pred_d.detach()
    cnt_g = metrics.binary_confusion_counts(pred_all, y.to(device), 0.5)
    counts_gathered = torch.stack([cnt_g[k] for k in ("TP", "TN", "FP", "FN")])

    # ---- one device, the concatenated batch ----------------------------------------------------------------------
    m_full = build(None)
    params_f = [p for p in m_full.parameters() if p.requires_grad]
    pred_f = m_full(x.to(device))
    loss_f = losses.DiceLoss()(pred_f, y.to(device))
    loss_f.backward()
    cnt_f = metrics.binary_confusion_counts(pred_f.detach(), y.to(device), 0.5)
    counts_f = torch.stack([cnt_f[k] for k in ("TP", "TN", "FP", "FN")])

    # the single-device step once more: run-to-run reproducibility of the reference point itself
    m_rep = build(None)
    loss_r = losses.DiceLoss()(m_rep(x.to(device)), y.to(device))
    loss_r.backward()
    gr = _flat([p.grad for p in m_rep.parameters() if p.requires_grad])

    gd, gf = _flat([p.grad for p in params_d]), _flat([p.grad for p in params_f])
    rep_cos = float(torch.dot(gr, gf) / (gr.norm() * gf.norm()).clamp_min(1e-30))
    rep_equal = bool(torch.equal(gr, gf)) and float(loss_r.detach()) == float(loss_f.detach())
    cos = float(torch.dot(gd, gf) / (gd.norm() * gf.norm()).clamp_min(1e-30))
    norm_rel = float((gd.norm() - gf.norm()).abs() / gf.norm().clamp_min(1e-30))
    loss_rel = abs(float(loss_d.detach()) - float(loss_f.detach())) / max(abs(float(loss_f.detach())), 1e-30)
    bn_rel = 0.0
    for a, b in zip(_bn_buffers(m_dist), _bn_buffers(m_full)):
        bn_rel = max(bn_rel, float((a - b).abs().max() / b.abs().max().clamp_min(1e-6)))
    exact = bool(torch.equal(counts_d, counts_gathered))
    n_pix = nb * size * size
    cnt_dev = float((counts_d - counts_f).abs().max()) / n_pix
    res = {"world": world, "loss_rel": loss_rel, "grad_cosine": cos, "grad_norm_rel": norm_rel, "bn_running_rel": bn_rel,
           "counters_exchange_bit_exact": exact, "counters_vs_single_device_frac": cnt_dev,
           "loss": float(loss_d.detach()), "loss_single_device": float(loss_f.detach()),
           "repeat_grad_cosine": rep_cos, "repeat_bit_identical": rep_equal}
    if os.environ.get("MSP_SELFCHECK_DETAIL"):                     # which parameters carry the difference
        names = [n for n, p in m_dist.named_parameters() if p.requires_grad]
        err_total = float((gd - gf).norm() ** 2)
        rows = []
        for n, a, b, r in zip(names, params_d, params_f, [p for p in m_rep.parameters() if p.requires_grad]):
            a, b, r = a.grad.double().reshape(-1), b.grad.double().reshape(-1), r.grad.double().reshape(-1)
            rows.append((float(((a - b).norm() ** 2) / max(err_total, 1e-300)), n,
                         float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300)), float(a.norm()), float(b.norm()),
                         float(torch.dot(r, b) / (r.norm() * b.norm()).clamp_min(1e-300))))
        rows.sort(reverse=True)
        res["buffers"] = [(n, float((a - b).abs().max() / b.abs().max().clamp_min(1e-6)))
                          for (n, a), (_, b) in zip(m_dist.named_buffers(), m_full.named_buffers())
                          if n.endswith("running_mean") or n.endswith("running_var")]
        res["worst"] = [{"err_share": round(e, 4), "param": n, "cos": round(c, 5), "norm_n": na, "norm_1": nb,
                         "repeat_cos": round(rc, 5)} for e, n, c, na, nb, rc in rows[:int(os.environ["MSP_SELFCHECK_DETAIL"])]]
    # noise floor: the single-device step against itself with the other summation order (atomics instead of rows)
    torch.use_deterministic_algorithms(False)
    try:
        m_noise = build(None)
        loss_n = losses.DiceLoss()(m_noise(x.to(device)), y.to(device))
        loss_n.backward()
    finally:
        torch.use_deterministic_algorithms(True, warn_only=True)
    gn = _flat([p.grad for p in m_noise.parameters() if p.requires_grad])
    floor_cos = float(torch.dot(gn, gf) / (gn.norm() * gf.norm()).clamp_min(1e-30))
    floor_norm = float((gn.norm() - gf.norm()).abs() / gf.norm().clamp_min(1e-30))
    floor_bn = 0.0
    for a, b in zip(_bn_buffers(m_noise), _bn_buffers(m_full)):
        floor_bn = max(floor_bn, float((a - b).abs().max() / b.abs().max().clamp_min(1e-6)))
    res.update(noise_floor_cosine=floor_cos, noise_floor_norm_rel=floor_norm, noise_floor_bn_rel=floor_bn)
    # (both sides of the comparison are single realisations of the same rounding noise: margins of 0.08 / 3x)
    grads_ok = (cos >= 0.999 and norm_rel <= 1e-2) or (cos >= floor_cos - 0.08 and norm_rel <= 2e-2)
    bn_ok = bn_rel <= max(1e-3, 3.0 * floor_bn)
    res["ok"] = bool(loss_rel <= 1e-3 and grads_ok and bn_ok and exact and cnt_dev <= 2e-3)
    if world > 1:                                                   # the ranks agree on the verdict
        flag = torch.tensor([1 if res["ok"] else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        res["ok"] = bool(flag.item())
    reducer.remove()
    return res


def _rms_rel(a, b) -> float:
    a, b = a.double(), b.double()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-300))


def exchange_parity(group=None, device=None, seed: int = 0) -> dict:
    """The three exchanges one layer deep on identical inputs (see the module docstring)."""
    was = torch.are_deterministic_algorithms_enabled()
    warn = torch.is_deterministic_algorithms_warn_only_enabled()
    torch.use_deterministic_algorithms(True, warn_only=True)
    try:
        return _exchange_parity(group, device, seed)
    finally:
        torch.use_deterministic_algorithms(was, warn_only=warn)


def _exchange_parity(group, device, seed) -> dict:
    world = dist.get_world_size(group) if (group is not None and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    grp = group if world > 1 else None
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    g = torch.Generator().manual_seed(seed + 5)
    nb, c, k, hw = 4 * world, 64, 96, 24
    x = torch.randn((nb, c, hw, hw), generator=g).bfloat16().float()
    dy = torch.randn((nb, k, hw, hw), generator=g).bfloat16().float()
    w0 = (torch.randn((k, c, 3, 3), generator=g) / (9 * c) ** 0.5).bfloat16().float()
    gamma0, beta0 = torch.rand(k, generator=g) + 0.5, torch.randn(k, generator=g) * 0.1
    lo, hi = shard_rows(nb, rank, world)

    def layer(xs, dys, grp_):
        w = torch.nn.Parameter(w0.clone().to(device))
        bn = torch.nn.BatchNorm2d(k).to(device).train()
        with torch.no_grad():
            bn.weight.copy_(gamma0)
            bn.bias.copy_(beta0)
        params = [w, bn.weight, bn.bias]
        red = GradReducer(params, bucket_mb=1.0, group=grp_)
        red.zero_grad()
        xin = xs.to(device).requires_grad_()
        z, stats = Fn.conv2d(Fn.to_nhwc(xin), w, None, 1, 1, want_stats=True)
        out = Fn.to_nchw(Fn.bn_act(z, stats, bn, act=ops.ACT_RELU, group=grp_), k)
        out.backward(dys.to(device))
        red.finish()
        red.remove()
        return out.detach(), xin.grad, [p.grad.detach().clone() for p in params], bn

    y_n, dx_n, g_n, bn_n = layer(x[lo:hi], dy[lo:hi], grp)
    y_1, dx_1, g_1, bn_1 = layer(x, dy, None)
    res = {"world": world,
           "syncbn_y_rms_rel": _rms_rel(y_n, y_1[lo:hi]), "syncbn_dx_rms_rel": _rms_rel(dx_n, dx_1[lo:hi]),
           "syncbn_running_rel": max(_rms_rel(bn_n.running_mean, bn_1.running_mean),
                                     _rms_rel(bn_n.running_var, bn_1.running_var)),
           # the reducer AVERAGES: world x mean over ranks of the local sums = the full-batch gradient
           "reduced_dw_rel": _rms_rel(g_n[0] * world, g_1[0]), "reduced_dgamma_rel": _rms_rel(g_n[1] * world, g_1[1]),
           "reduced_dbeta_rel": _rms_rel(g_n[2] * world, g_1[2])}

    prob = torch.rand((nb, 1, hw * 4, hw * 4), generator=g)
    mask = (torch.rand((nb, 1, hw * 4, hw * 4), generator=g) < 0.3).long()

    def dice(ps, ms, grp_):
        p = ps.to(device).requires_grad_()
        l = losses.DiceLoss(group=grp_)(p, ms.to(device))
        l.backward()
        return float(l.detach()), p.grad

    l_n, dp_n = dice(prob[lo:hi], mask[lo:hi], grp)
    l_1, dp_1 = dice(prob, mask, None)
    res["dice_loss_rel"] = abs(l_n - l_1) / abs(l_1)
    res["dice_grad_rel"] = _rms_rel(dp_n / world, dp_1[lo:hi])     # the local gradient carries the reducer's 1/world
    res["ok"] = bool(res["syncbn_y_rms_rel"] <= 1e-4 and res["syncbn_dx_rms_rel"] <= 1e-4
                     and res["syncbn_running_rel"] <= 1e-5 and res["reduced_dw_rel"] <= 1e-4
                     and res["reduced_dgamma_rel"] <= 1e-4 and res["reduced_dbeta_rel"] <= 1e-4
                     and res["dice_loss_rel"] <= 1e-6 and res["dice_grad_rel"] <= 1e-5)
    if world > 1:
        flag = torch.tensor([1 if res["ok"] else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        res["ok"] = bool(flag.item())
    return res


def main() -> None:
    """`torchrun --nproc-per-node N -m medsegpretrainimagenet_b200.selfcheck` prints one JSON line on rank 0 and exits
    non-zero when the N-rank step differs from the single-device step."""
    import json
    import os
    import sys
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
        if os.environ.get("MSP_SELFCHECK_EXCHANGE", "p2p") == "p2p":
            from . import parallel
            parallel.enable_peer_allreduce(group)
    res = n_rank_parity(group)
    res["exchange"] = exchange_parity(group)
    res["ok"] = bool(res["ok"] and res["exchange"]["ok"])
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
