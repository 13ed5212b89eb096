"""Device timing of the fused BatchNorm kernels on the ResNet-50 (B=256) activation shapes against the HBM roofline
(algorithmic bytes: fwd 4 B/elt (+2 residual), bwd_reduce 6 B/elt, bwd_apply 8 B/elt (+2 shortcut gradient))."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medsegpretrainimagenet_b200 import ops
try:
    BW = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    BW = 6650.0
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shapes = [("stem", 112, 64, 1), ("L0 mid", 56, 64, 6), ("L0 out", 56, 256, 3), ("L1 c1", 56, 128, 1), ("L1 mid", 28, 128, 7),
          ("L1 out", 28, 512, 4), ("L2 c1", 28, 256, 1), ("L2 mid", 14, 256, 11), ("L2 out", 14, 1024, 6),
          ("L3 c1", 14, 512, 1), ("L3 mid", 7, 512, 5), ("L3 out", 7, 2048, 3)]
dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=4):
    ts = []
    for it in range(iters + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it: ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
tot = {"fwd": [0, 0], "red": [0, 0], "app": [0, 0]}
print(f"{'layer':8s} {'x':>3s} {'elts(M)':>8s} |   fwd us   GB/s  frac |   red us   GB/s  frac |   app us   GB/s  frac")
for name, hw, c, cnt in shapes:
    x = torch.randn((B, hw, hw, c), device=dev).to(torch.bfloat16)
    dy = torch.randn_like(x)
    res = torch.randn_like(x) if "out" in name else None
    mi = torch.stack([torch.zeros(c, device=dev), torch.ones(c, device=dev)])
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    y = ops.bn_act_fwd(x, mi, gamma, beta, ops.ACT_RELU, residual=res)
    n = x.numel()
    t_f = timeit(lambda: ops.bn_act_fwd(x, mi, gamma, beta, ops.ACT_RELU, residual=res, out=y))
    sums = ops.bn_act_bwd_reduce(x, y, dy, mi, ops.ACT_RELU)
    t_r = timeit(lambda: ops.bn_act_bwd_reduce(x, y, dy, mi, ops.ACT_RELU))
    dres = torch.empty_like(x) if res is not None else None
    t_a = timeit(lambda: ops.bn_act_bwd_apply(x, y, dy, mi, gamma, ops.ACT_RELU, sums, n // c, dres=dres))
    by = {"fwd": n * (4 + (2 if res is not None else 0)), "red": n * 6, "app": n * (8 + (2 if res is not None else 0))}
    cells = []
    for k, t in (("fwd", t_f), ("red", t_r), ("app", t_a)):
        gbs = by[k] / t / 1e3
        cells.append(f"{t:8.1f} {gbs:6.0f} {gbs / BW:5.2f}")
        tot[k][0] += t * cnt; tot[k][1] += by[k] / BW / 1e3 * cnt
    print(f"{name:8s} {cnt:3d} {n / 1e6:8.1f} | " + " | ".join(cells), flush=True)
for k, v in tot.items():
    print(f"TOTAL {k}: {v[0] / 1e3:.2f} ms measured vs {v[1] / 1e3:.2f} ms roofline -> {v[1] / v[0]:.3f}")
