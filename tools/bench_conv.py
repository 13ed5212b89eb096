"""Per-layer device timing of the convolution kernels against the per-layer roofline (GPU box).

    python tools/bench_conv.py [--batch 256] [--model r50|unet50] [--kinds fprop,dgrad,wgrad] [--iters 5]

ideal = max(flops / sustained bf16 peak, algorithmic bytes / measured HBM copy bandwidth); L2 is flushed (a
512 MB memset) before every timed launch, so the numbers are cold-L2 like a real step."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medsegpretrainimagenet_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--model", default="r50")
ap.add_argument("--kinds", default="fprop,dgrad,wgrad")
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--only", default="")
args = ap.parse_args()
try:
    pk = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except Exception:
    pk = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
TC, BW = pk["bf16_tflops_sustained"] * 1e12, pk["hbm_gbs"] * 1e9

def r50_layers(res=224):
    L = [("stem", res, 8, 64, 7, 2, 3)]
    h, cin = res // 4, 64
    for li, (c, n) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3))):
        for b in range(n):
            s = 2 if (b == 0 and li > 0) else 1
            L.append((f"L{li}b{b}c1", h, cin, c, 1, 1, 0))
            L.append((f"L{li}b{b}c2", h, c, c, 3, s, 1))
            h //= s
            L.append((f"L{li}b{b}c3", h, c, 4 * c, 1, 1, 0))
            cin = 4 * c
    return L

def unet50_decoder(res=256):
    # R50 attention U-Net decoder, input res x res; (name, H_in, Cin, Cout, k, stride, pad)
    L = []
    h = res // 32
    chans = [2048, 256, 128, 64, 32, 16]
    skips = [1024, 512, 256, 64, 0]
    x_c = 2048
    for lvl in range(5):
        up_out = x_c // 2
        L.append((f"d{lvl}_up", 2 * h, x_c, up_out, 2, 1, "same"))
        if skips[lvl]:
            sc = skips[lvl]
            L.append((f"d{lvl}_gs", h, x_c, x_c, 1, 1, 0))
            L.append((f"d{lvl}_Ws", 2 * h, sc, x_c, 2, 2, 0))
            L.append((f"d{lvl}_Wg", h, x_c, x_c, 1, 1, 0))
            L.append((f"d{lvl}_psi", h, x_c, sc, 1, 1, 0))
            cin = up_out + sc
        else:
            cin = up_out
        L.append((f"d{lvl}_c0", 2 * h, cin, chans[lvl + 1], 3, 1, 1))
        L.append((f"d{lvl}_c1", 2 * h, chans[lvl + 1], chans[lvl + 1], 3, 1, 1))
        x_c = chans[lvl + 1]
        h *= 2
    return L

layers = r50_layers(224) if args.model == "r50" else r50_layers(256)[:] + unet50_decoder(256)
seen, uniq = {}, []
for l in layers:
    key = l[1:]
    if key in seen:
        seen[key][1] += 1
    else:
        seen[key] = [l[0], 1]
        uniq.append(l)
dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
kinds = args.kinds.split(",")
B = args.batch
tot = {k: [0.0, 0.0] for k in kinds}
print(f"{'layer':10s} {'x':>3s} {'M':>8s} {'N':>5s} {'K':>5s} | " + " | ".join(f"{k:>6s} us  ideal  frac   TF/s" for k in kinds))
for name, hin, ci, co, k, s, pad in uniq:
    if args.only and args.only not in name:
        continue
    cnt = seen[(hin, ci, co, k, s, pad)][1]
    ho, wo, pt, pl = ops.conv_out_size(hin, hin, k, k, s, pad)
    M = B * ho * wo
    rowwin = name == "stem"
    if rowwin:
        ci = 3
        win_px, cpp = ops.rowwin_geometry(ci, k, s)
        wp = max(hin + pl, s * (wo - 1) + win_px); wp += wp & 1
        x = ops.nchw_to_rowwin(torch.randn((B, ci, hin, hin), device=dev), cpp, pl, wp)
        w = torch.randn((co, ci, k, k), device=dev) * 0.05
        wf, wd = ops.pack_weights_rowwin(w, win_px), None
    else:
        x = torch.randn((B, hin, hin, ci), device=dev).to(torch.bfloat16)
        w = torch.randn((co, ci, k, k), device=dev) * 0.05
        wf, wd = ops.pack_weights(w)
    dy = torch.randn((B, ho, wo, co), device=dev).to(torch.bfloat16)
    stats = torch.zeros((2, co), device=dev)
    fl = 2.0 * M * co * ci * k * k
    by = x.numel() * 2 + dy.numel() * 2 + 2 * w.numel()
    ideal = max(fl / TC, by / BW) * 1e6
    cells = []
    for kind in kinds:
        if kind == "dgrad" and name == "stem":
            cells.append(" " * 32)
            continue
        ts = []
        for it in range(args.iters + 1):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if kind == "fprop" and rowwin:
                ops.conv_fprop_rowwin(x, hin, wf, None, co, k, k, s, pt, pl, ho, wo, win_px, stats=stats)
            elif kind == "wgrad" and rowwin:
                ops.conv_wgrad_rowwin(x, hin, dy, ci, k, k, s, pt, pl, win_px)
            elif kind == "fprop":
                ops.conv_fprop(x, wf, None, co, k, k, s, pt, pl, ho, wo, stats=stats)
            elif kind == "dgrad":
                ops.conv_dgrad(dy, wd, tuple(x.shape), k, k, s, pt, pl)
            else:
                ops.conv_wgrad(x, dy, ci, k, k, s, pt, pl)
            e1.record()
            torch.cuda.synchronize()
            if it:
                ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[len(ts) // 2]
        tot[kind][0] += t * cnt
        tot[kind][1] += ideal * cnt
        cells.append(f"{t:9.1f} {ideal:6.1f} {ideal / t:5.2f} {fl / t / 1e6:6.0f}")
    print(f"{name:10s} {cnt:3d} {M:8d} {co:5d} {ci * k * k:5d} | " + " | ".join(cells), flush=True)
    del x, dy, w, wf, wd
for kind in kinds:
    print(f"TOTAL {kind}: {tot[kind][0] / 1e3:.2f} ms measured vs {tot[kind][1] / 1e3:.2f} ms roofline -> {tot[kind][1] / max(tot[kind][0], 1e-9):.3f}")
