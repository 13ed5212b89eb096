"""Host-side logic of the data-parallel path on CPU: world size 2, gloo backend (the GPU path uses the same code with
NCCL).  parallel.GradReducer: bucketed all-reduce driven by post-accumulate-grad hooks, `param.grad` aliased to the
bucket storage, averaged gradients equal the full-batch gradients.  The SyncBN / batchwise-Dice exchanges are sums of
per-rank sufficient statistics: checked here as algebra on the oracle's formulas."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from medsegpretrainimagenet_b200.parallel import GradReducer, shard_rows
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                                    torch.nn.Linear(16, 3))
        frozen = model[2].bias
        frozen.requires_grad_(False)
        params = [p for p in model.parameters() if p.requires_grad]
        reducer = GradReducer(params, bucket_mb=0.0005, group=dist.group.WORLD)   # several small buckets
        assert len(reducer.buckets) >= 3
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn((8, 12), generator=g), torch.randn((8, 3), generator=g)
        lo, hi = shard_rows(8, rank, world)
        for step in range(2):                               # second step: zero_grad keeps the aliasing
            reducer.zero_grad()
            loss = torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi])
            loss.backward()
            reducer.finish()
            for p in params:
                assert p.grad.data_ptr() >= reducer._bucket_of[p].flat.data_ptr()   # still a view of the bucket
        got = [p.grad.clone() for p in params]
        # full-batch reference on every rank
        ref_model = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                                        torch.nn.Linear(16, 3))
        ref_model.load_state_dict(model.state_dict())
        torch.nn.functional.mse_loss(ref_model(x), y).backward()
        ref = [p.grad for n, p in ref_model.named_parameters() if n != "2.bias"]
        ok = all(torch.allclose(a, b, rtol=1e-5, atol=1e-7) for a, b in zip(got, ref))
        # gradient accumulation (train_model.py:53-55, loss/loss.py:83-84: loss / accumulation_scale per fragment):
        # two micro-batches per rank, the first under no_sync(), one exchange after the second == full batch
        reducer.zero_grad()
        mid = (lo + hi) // 2
        with reducer.no_sync():
            (torch.nn.functional.mse_loss(model(x[lo:mid]), y[lo:mid]) / 2).backward()
        (torch.nn.functional.mse_loss(model(x[mid:hi]), y[mid:hi]) / 2).backward()
        reducer.finish()
        ok &= all(torch.allclose(p.grad, b, rtol=1e-5, atol=1e-7) for p, b in zip(params, ref))
        # the two silent failures of the first version now raise: a further backward after the exchange ...
        try:
            torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi]).backward()
            ok = False
        except RuntimeError as e:
            ok &= "no_sync" in str(e)
        # ... and gradients that were detached from the buckets by optimizer.zero_grad(set_to_none=True)
        reducer.zero_grad()
        torch.optim.SGD(params, lr=0.1).zero_grad(set_to_none=True)
        try:
            torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi]).backward()
            ok = False
        except RuntimeError as e:
            ok &= "aliases" in str(e)
        reducer.remove()
        # sufficient statistics: BatchNorm [sum, sum of squares] and Dice [I, Y, S] all-reduced == full batch
        feats = torch.randn((8, 5, 6, 6), generator=g)
        st = torch.stack([feats[lo:hi].sum((0, 2, 3)), (feats[lo:hi] ** 2).sum((0, 2, 3))])
        # the SyncBN choke point (functional._allreduce_sum -> parallel.allreduce_small_sum_): without a peer-memory
        # communicator (CPU / gloo) it is the process group's all-reduce
        from medsegpretrainimagenet_b200.parallel import allreduce_small_sum_, peer_allreduce_for
        assert peer_allreduce_for(dist.group.WORLD) is None
        allreduce_small_sum_(st, dist.group.WORLD)
        ok &= torch.allclose(st[0], feats.sum((0, 2, 3)), rtol=1e-5, atol=1e-5)
        ok &= torch.allclose(st[1], (feats ** 2).sum((0, 2, 3)), rtol=1e-5, atol=1e-5)
        from oracle import ref_losses
        p = torch.rand((8, 3, 6, 6), generator=g)
        m = torch.randint(0, 3, (8, 1, 6, 6), generator=g)
        sums = torch.zeros((3, 3), dtype=torch.float64)
        for c in range(3):
            yy = (m[lo:hi, 0] == c).double()
            pc = p[lo:hi, c].double()
            sums[c] = torch.stack([(yy * pc).sum(), yy.sum(), (pc * pc).sum()])
        dist.all_reduce(sums)
        dice = (2 * sums[:, 0] + 1e-5) / (sums[:, 1] + sums[:, 2] + 1e-5)
        ok &= abs(float(1 - dice.mean()) - float(ref_losses.dice_loss(p, m))) < 1e-6
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_grad_reducer_and_statistics_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_post_accumulate_hook_fires_for_gradients_that_bypass_autograd():
    """parallel.GradReducer counts a bucket's parameters through post-accumulate-grad hooks; convolution weights hand
    autograd None (their gradient is written into `param.grad` by ops._WgradQueue).  The hook must still fire for them."""
    class Bypass(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            ctx.w = w
            return x * w.detach()

        @staticmethod
        def backward(ctx, g):
            ctx.w.grad.add_(1.0)            # "written by the kernel"
            return g, None

    w = torch.nn.Parameter(torch.ones(3))
    w.grad = torch.zeros(3)
    fired = []
    w.register_post_accumulate_grad_hook(lambda p: fired.append(p.grad.clone()))
    Bypass.apply(torch.ones(3, requires_grad=True), w).sum().backward()
    assert len(fired) == 1 and torch.equal(fired[0], torch.ones(3))


def test_shard_rows():
    from medsegpretrainimagenet_b200.parallel import shard_rows
    assert [shard_rows(48, r, 4) for r in range(4)] == [(0, 12), (12, 24), (24, 36), (36, 48)]
