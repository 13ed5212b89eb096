"""N-rank == 1-rank parity of the data-parallel step, runnable wherever N GPUs are (SURVEY.md §4 tier 6).

The reference's `nn.DataParallel` (train_model.py:192-197) normalises every replica's shard with its own BatchNorm
statistics; this path synchronises the statistics (SyncBN through the peer-memory kernel or NCCL), reduces batchwise Dice
over GLOBAL sums and averages the gradients, so that N ranks on N shards reproduce the SINGLE-DEVICE step on the
concatenated batch.  `n_rank_parity()` checks exactly that on a ResNet-18 attention U-Net: every rank also runs the
full-batch step locally (same seeds, no process group) and compares

    loss                      relative difference <= 1e-3
    parameter gradients       cosine >= 0.999 (all parameters flattened), norm within 1 %
    BatchNorm running stats   relative difference <= 1e-3 (max-norm over each buffer)
    confusion counters        all-reduced per-rank counters == counters of the gathered predictions, BIT-EXACT
                              (and within 0.2 % of the pixels of the single-device step's counters)

Used by tests/test_multigpu_gpu.py (under torchrun when >= 2 GPUs are visible) and as the pre-check of
`bench.py --gpus N` (`"parity_n"` in its JSON line): the driver's GPU test box has one GPU."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import losses, metrics, models
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
    x = torch.rand((nb, 1, size, size), generator=g)
    y = (torch.rand((nb, 1, size, size), generator=g) < 0.3).long()
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
    res["ok"] = bool(loss_rel <= 1e-3 and cos >= 0.999 and norm_rel <= 1e-2 and bn_rel <= 1e-3 and exact
                     and cnt_dev <= 2e-3)
    if world > 1:                                                   # the ranks agree on the verdict
        flag = torch.tensor([1 if res["ok"] else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        res["ok"] = bool(flag.item())
    reducer.remove()
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
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
