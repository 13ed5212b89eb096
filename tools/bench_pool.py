"""Max-pool forward / backward at the ResNet stem's size (3x3/2 pad 1 on N x 112 x 112 x 64, classification/models.py:49)
and the U-Net encoder's (2x2/2): us per launch (CUDA events, cold L2) against bytes / measured HBM copy bandwidth.
    python tools/bench_pool.py [--batch 256]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medsegpretrainimagenet_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda")
try:
    hbm = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6552.6
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


for (n, h, w, c, k, s, p) in [(args.batch, 112, 112, 64, 3, 2, 1), (max(args.batch // 64, 1), 1024, 1024, 64, 2, 2, 0)]:
    x = torch.randn((n, h, w, c), device=dev).to(torch.bfloat16)
    y, idx = ops.maxpool_fwd(x, k, s, p)
    dy = torch.randn_like(y)
    t_f = timed(lambda: ops.maxpool_fwd(x, k, s, p))
    t_b = timed(lambda: ops.maxpool_bwd(idx, dy, tuple(x.shape), k, s, p))
    b_f = x.numel() * 2 + y.numel() * 2 + idx.numel()
    b_b = x.numel() * 2 + y.numel() * 2 + idx.numel()
    print(f"maxpool {k}x{k}/{s} on {n}x{h}x{w}x{c}: fwd {t_f:7.1f} us ({b_f / t_f / 1e3 / hbm:.2f} of HBM, ideal {b_f / hbm / 1e3:.1f} us)   "
          f"bwd {t_b:7.1f} us ({b_b / t_b / 1e3 / hbm:.2f} of HBM)")
