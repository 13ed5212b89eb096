"""Achieved HBM bandwidth of the loss / metric / robustness kernels at the BASELINE configs' shapes (algorithmic bytes as
defined in DESIGN.md 3.3-3.4, L2 flushed before every timed launch)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medsegpretrainimagenet_b200 as b
from medsegpretrainimagenet_b200 import ops
try:
    BW = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    BW = 6650.0
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
def row(name, t, by):
    print(f"{name:58s} {t:9.1f} us {by / 1e6:9.1f} MB {by / t / 1e3:7.0f} GB/s {by / t / 1e3 / BW:5.2f} of measured HBM", flush=True)
g = torch.Generator(device=dev).manual_seed(0)
# cfg3: 4-class 256x256, batch 24
p4 = torch.softmax(torch.randn((24, 4, 256, 256), device=dev, generator=g), 1)
m4 = torch.randint(0, 4, (24, 1, 256, 256), device=dev, generator=g)
n = 24 * 256 * 256
row("dice sums (cfg3, 4 classes)", timeit(lambda: ops.dice_sums(p4, m4, False, 0, True)), n * (4 * 4 + 8))
sums = ops.dice_sums(p4, m4, False, 0, True); loss, coef = ops.dice_finalize(sums, 0, 1e-5)
row("dice gradient (cfg3)", timeit(lambda: ops.dice_bwd(p4, m4, False, 0, True, coef)), n * (4 * 4 + 8 + 4 * 4))
row("multi-class confusion matrix (cfg3)", timeit(lambda: ops.confusion_multiclass(p4, m4, False)), n * (4 * 4 + 8))
# cfg4: 5-channel multilabel 1024x1024, batch 4
p5 = torch.rand((4, 5, 1024, 1024), device=dev, generator=g)
t5 = (torch.rand((4, 5, 1024, 1024), device=dev, generator=g) < 0.05).float()
n5 = p5.numel()
row("binary confusion counters per channel (cfg4, fp32 target)", timeit(lambda: ops.confusion_binary(p5, t5, 0.5, True)), n5 * 8)
row("BCE loss + gradient (cfg4)", timeit(lambda: ops.bce(p5, t5, True, want_loss=True, want_grad=True)), n5 * 12)
# cfg2: 1000-way softmax CE + top-5, batch 256 (tiny)
lg = torch.randn((256, 1000), device=dev, generator=g); lab = torch.randint(0, 1000, (256, 1), device=dev, generator=g)
row("softmax-CE fwd+bwd (cfg2 head, 256x1000)", timeit(lambda: ops.softmax_ce(lg, lab, 0.1, want_loss=True, want_grad=True)), 256 * 1000 * 8)
# final 1x1 conv + softmax head (cfg3): 16 -> 4 channels
xh = torch.randn((24, 256, 256, 16), device=dev).to(torch.bfloat16); wh = torch.randn((4, 16), device=dev); bh = torch.zeros(4, device=dev)
row("final 1x1 conv + softmax head (cfg3)", timeit(lambda: ops.final_conv_act_fwd(xh, wh, bh, 2)), n * (16 * 2 + 4 * 4))
# cfg5: robustness distances, level 5 pooled (N=50000, D=2048) and unpooled rows (N=2048, D=100352)
for nn_, d in ((50000, 2048), (2048, 100352)):
    q = torch.relu(torch.randn((nn_, d), device=dev, generator=g)); k = torch.relu(q + 0.1 * torch.randn((nn_, d), device=dev, generator=g))
    row(f"robustness: 3 distances x pos/neg pairs (N={nn_}, D={d})", timeit(lambda: ops.rowpair_distances(q, k)), 2 * nn_ * d * 4)
    del q, k
