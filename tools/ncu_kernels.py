"""One launch (after one warm-up launch) of every kernel family of the hot path at a BASELINE-config shape, for
`ncu --set full` (GPU box):

    MSP_NCU_ONCE=1 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
        --clock-control none -k regex:'tapgemm|wgrad_kernel|bn_act|maxpool|dice|confusion|softmax_ce|final_conv|rowpair|bce' \
        -o /tmp/prof_kernels -f python tools/ncu_kernels.py
    python tools/summarize_ncu.py /tmp/prof_kernels.ncu-rep > gpurun_out/ncu_kernels_summary.txt

(`--set full` over all ~30 launches takes 9 GPU-minutes and a 100 MB report; the full-set captures of the dominant
kernel are taken separately, one launch at a time.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medsegpretrainimagenet_b200 import ops

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
B = int(os.environ.get("MSP_NCU_BATCH", "256"))


def twice(fn):
    for _ in range(1 if os.environ.get("MSP_NCU_ONCE") == "1" else 2):   # ncu replays every launch itself
        fn()
    torch.cuda.synchronize()


def conv_layer(hin, ci, co, k, s, pad, kinds=("fprop", "dgrad", "wgrad")):
    ho, wo, pt, pl = ops.conv_out_size(hin, hin, k, k, s, pad)
    x = torch.randn((B, hin, hin, ci), device=dev).to(torch.bfloat16)
    w = torch.randn((co, ci, k, k), device=dev) * 0.05
    wf, wd = ops.pack_weights(w)
    dy = torch.randn((B, ho, wo, co), device=dev).to(torch.bfloat16)
    stats = torch.zeros((2, co), device=dev)
    if "fprop" in kinds:
        twice(lambda: ops.conv_fprop(x, wf, None, co, k, k, s, pt, pl, ho, wo, stats=stats))
    if "dgrad" in kinds:
        twice(lambda: ops.conv_dgrad(dy, wd, tuple(x.shape), k, k, s, pt, pl))
    if "wgrad" in kinds:
        twice(lambda: ops.conv_wgrad(x, dy, ci, k, k, s, pt, pl))


# ---- convolutions (ResNet-50, B=256): tensor-bound 3x3 (tap path), HBM-bound 1x1, halo 3x3, narrow 1x1 ----------------
conv_layer(14, 256, 256, 3, 1, 1)          # L2b1c2: tapgemm<256> fprop+dgrad, wgrad
conv_layer(14, 256, 1024, 1, 1, 0)         # L2b0c3: tapgemm<256> 1x1 expand (epilogue / HBM bound)
conv_layer(28, 128, 128, 3, 1, 1, ("fprop", "dgrad"))   # L1b1c2: tapgemm<128>
conv_layer(56, 64, 64, 3, 1, 1, ("fprop",))             # L0b0c2: tapgemm_halo<64> (resident weights)
conv_layer(56, 256, 64, 1, 1, 0, ("fprop",))            # L0b1c1: tapgemm<64>

# ---- fused BatchNorm kernels on a block output (L1 out: 28x28x512, with residual) -----------------------------------
hw, c = 28, 512
x = torch.randn((B, hw, hw, c), device=dev).to(torch.bfloat16)
dy = torch.randn_like(x)
res = torch.randn_like(x)
n = B * hw * hw
mi = torch.stack([torch.zeros(c, device=dev), torch.ones(c, device=dev)])
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
y = ops.bn_act_fwd(x, mi, gamma, beta, ops.ACT_RELU, residual=res)
twice(lambda: ops.bn_act_fwd(x, mi, gamma, beta, ops.ACT_RELU, residual=res, out=y))
twice(lambda: ops.bn_act_bwd_reduce(x, y, dy, mi, ops.ACT_RELU))
sums = ops.bn_act_bwd_reduce(x, y, dy, mi, ops.ACT_RELU)
dres = torch.empty_like(x)
twice(lambda: ops.bn_act_bwd_apply(x, y, dy, mi, gamma, ops.ACT_RELU, sums, n, dres=dres))

# ---- stem max-pool ---------------------------------------------------------------------------------------------------
xs = torch.randn((B, 112, 112, 64), device=dev).to(torch.bfloat16)
twice(lambda: ops.maxpool_fwd(xs, 3, 2, 1))
yp, idx = ops.maxpool_fwd(xs, 3, 2, 1)
dyp = torch.randn_like(yp)
twice(lambda: ops.maxpool_bwd(idx, dyp, tuple(xs.shape), 3, 2, 1))
del xs, yp, idx, dyp, x, dy, res, y, dres

# ---- losses / metrics (cfg3: 4-class 256x256 batch 24; cfg4: 5-channel 1024x1024 batch 4; cfg2 head) ---------------
p4 = torch.softmax(torch.randn((24, 4, 256, 256), device=dev, generator=g), 1)
m4 = torch.randint(0, 4, (24, 1, 256, 256), device=dev, generator=g)
twice(lambda: ops.dice_sums(p4, m4, False, 0, True))
sums4 = ops.dice_sums(p4, m4, False, 0, True)
loss, coef = ops.dice_finalize(sums4, 0, 1e-5)
twice(lambda: ops.dice_bwd(p4, m4, False, 0, True, coef))
twice(lambda: ops.confusion_multiclass(p4, m4, False))
p5 = torch.rand((4, 5, 1024, 1024), device=dev, generator=g)
t5 = (torch.rand((4, 5, 1024, 1024), device=dev, generator=g) < 0.05).float()
twice(lambda: ops.confusion_binary(p5, t5, 0.5, True))
twice(lambda: ops.bce(p5, t5, True, want_loss=True, want_grad=True))
lg = torch.randn((256, 1000), device=dev, generator=g)
lab = torch.randint(0, 1000, (256, 1), device=dev, generator=g)
twice(lambda: ops.softmax_ce(lg, lab, 0.1, want_loss=True, want_grad=True))
xh = torch.randn((24, 256, 256, 16), device=dev).to(torch.bfloat16)
wh, bh = torch.randn((4, 16), device=dev), torch.zeros(4, device=dev)
twice(lambda: ops.final_conv_act_fwd(xh, wh, bh, 2))

# ---- robustness distances (cfg5 level 5 pooled: N=50000, D=2048) ----------------------------------------------------
q = torch.relu(torch.randn((50000, 2048), device=dev, generator=g))
k = torch.relu(q + 0.1 * torch.randn((50000, 2048), device=dev, generator=g))
twice(lambda: ops.rowpair_distances(q, k))
print("ncu_kernels: done")
