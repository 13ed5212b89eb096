"""Peer-memory all-reduce (csrc/msp_p2p.cu) against NCCL on N GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/test_p2p.py

Checks (1) eager exchanges of several sizes against dist.all_reduce, (2) the same exchanges captured in a CUDA graph and
replayed, (3) latency of one exchange vs one NCCL all-reduce of the same vector (device time, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from medsegpretrainimagenet_b200.parallel import PeerAllReduce

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
par = PeerAllReduce(dist.group.WORLD, max_floats=8192)
g = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
for it in range(60):
    n = (8, 128, 1000, 4096, 8192, 130)[it % 6]
    t = torch.randn(n, device=dev, generator=g)
    ref = t.clone()
    dist.all_reduce(ref)
    par.allreduce_sum_(t)
    worst = max(worst, ((t - ref).abs().max() / ref.abs().max()).item())
assert worst <= 1e-6, worst
# identical on every rank, bit for bit (fixed summation order)
t = torch.randn(4096, device=dev, generator=g)
par.allreduce_sum_(t)
gathered = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(gathered, t)
assert all(torch.equal(gathered[0], x) for x in gathered)
# CUDA graph: 16 exchanges per replay
bufs = [torch.zeros(1024, device=dev) for _ in range(16)]
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for b in bufs:
        par.allreduce_sum_(b)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for b in bufs:
        par.allreduce_sum_(b)
for rep in range(5):
    for i, b in enumerate(bufs):
        b.fill_(float(rank + 1 + i + rep))
    graph.replay()
    torch.cuda.synchronize()
    for i, b in enumerate(bufs):
        exp = sum(float(r + 1 + i + rep) for r in range(world))
        assert torch.all(b == exp), (rep, i, b[:4], exp)
# fused SyncBN statistic exchange (msp_p2p_stats_exchange): rows -> local sums -> global sums -> mean / invstd / running
# stats in ONE launch == reduce_rows -> all-reduce -> bn_finalize, bit for bit (same row order, same rank order)
from medsegpretrainimagenet_b200 import ops
for it, (rows, c) in enumerate([(1, 64), (148, 16), (296, 256), (148, 2048), (37, 520), (1, 4096)]):
    ws = torch.randn((rows, 2, c), device=dev, generator=g).abs_()
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    rm2, rv2 = rm.clone(), rv.clone()
    count = 1000.0 * world
    loc = ops.reduce_rows(ws.clone())
    glob = loc.clone()
    par.allreduce_sum_(glob.view(-1))
    mi_ref = ops.bn_finalize(glob.clone(), count, 1e-5, 0.1, rm, rv)
    mi = torch.empty((2, c), device=dev)
    ws2 = ws.clone()
    par.stats_exchange(ws2, rows, c, reset=True, finalize=(count, 1e-5, 0.1, mi, rm2, rv2))
    assert torch.equal(mi, mi_ref) and torch.equal(rm, rm2) and torch.equal(rv, rv2), (rows, c)
    assert not ws2.any()                                       # reset: the rows are zeroed for the next step
    lo, gl = torch.empty((2, c), device=dev), torch.empty((2, c), device=dev)
    par.stats_exchange(ws.clone(), rows, c, local_out=lo, global_out=gl)
    assert torch.equal(lo, loc) and torch.equal(gl, glob), (rows, c)
# ... captured in a graph and replayed (the exchange number advances inside the kernels: last block to finish)
wss = [torch.zeros((148, 2, 256), device=dev) for _ in range(12)]
outs = [torch.empty((2, 256), device=dev) for _ in range(12)]
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for w_, o_ in zip(wss, outs):
        par.stats_exchange(w_, 148, 256, global_out=o_)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    for w_, o_ in zip(wss, outs):
        par.stats_exchange(w_, 148, 256, global_out=o_)
for rep in range(5):
    for i, w_ in enumerate(wss):
        w_.fill_(float(rank + 1 + i + rep))
    g2.replay()
    torch.cuda.synchronize()
    for i, o_ in enumerate(outs):
        exp = 148.0 * sum(float(r + 1 + i + rep) for r in range(world))
        assert torch.all(o_ == exp), (rep, i, o_[0, :4], exp)
# latency
def timed(fn, iters=200):
    for _ in range(20):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
x = torch.randn(1024, device=dev)
t_p2p = timed(lambda: par.allreduce_sum_(x))
t_nccl = timed(lambda: dist.all_reduce(x))
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(50):
        par.allreduce_sum_(x)
t_p2p_g = timed(lambda: gr.replay(), 20) / 50
gn = torch.cuda.CUDAGraph()
with torch.cuda.graph(gn):
    for _ in range(50):
        dist.all_reduce(x)
t_nccl_g = timed(lambda: gn.replay(), 20) / 50
if rank == 0:
    print(f"p2p all-reduce ok on {world} GPUs: max rel diff vs NCCL {worst:.1e}; 4 KB exchange: peer kernel {t_p2p:.1f} us eager / "
          f"{t_p2p_g:.1f} us in a graph, NCCL {t_nccl:.1f} us eager / {t_nccl_g:.1f} us in a graph", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
