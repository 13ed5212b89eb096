"""GPU diagnostic for the folded up-conv dgrad: kernel vs the fold formula evaluated with torch on the same operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from medsegpretrainimagenet_b200 import ops
dev = "cuda"
torch.manual_seed(0)
for (n, cin, cout, h, w) in ((1, 64, 32, 6, 6), (2, 64, 32, 7, 9), (1, 8, 8, 4, 4)):
    wt = (torch.randn(cout, cin, 2, 2) * 0.1).to(torch.bfloat16).float().to(dev)
    dy = torch.randn(n, 2 * h, 2 * w, cout).to(torch.bfloat16).to(dev)
    fold = ops.fold_upconv_weight(wt)                       # (K, C, 1, 9)
    wf9, wd9 = ops.pack_weights(fold, True)
    dx = ops.upconv2x_dgrad(dy, wd9, (n, h, w, cin)).float()
    fw = wd9.float()                                        # [C][9][K]
    gp = F.pad(dy.float(), (0, 0, 2, 2, 2, 2))              # pad W and H by 2
    ref = torch.zeros(n, h, w, cin, device=dev)
    per_tap = []
    t = 0
    for a in (0, 1):
        for b in (0, 1):
            for dr in range(a + 1):
                for dq in range(b + 1):
                    oh, ow = a - 2 * dr, b - 2 * dq
                    sl = gp[:, 2 + oh:2 + oh + 2 * h:2, 2 + ow:2 + ow + 2 * w:2, :]     # (n, h, w, K)
                    ref += torch.einsum("nhwk,ck->nhwc", sl, fw[:, t, :])
                    t += 1
    err = (dx - ref).abs()
    print(f"case {(n, cin, cout, h, w)}: max err {err.max().item():.4f} of {ref.abs().max().item():.3f}")
    print(" per row :", [round(v, 3) for v in err.amax(dim=(0, 2, 3)).tolist()])
    print(" per col :", [round(v, 3) for v in err.amax(dim=(0, 1, 3)).tolist()])
    print(" per img :", [round(v, 3) for v in err.amax(dim=(1, 2, 3)).tolist()])
    # single-tap probes: only folded tap t non-zero
    for t in range(9):
        f1 = torch.zeros_like(fold); f1[:, :, 0, t] = fold[:, :, 0, t]
        _, wd1 = ops.pack_weights(f1, True)
        d1 = ops.upconv2x_dgrad(dy, wd1, (n, h, w, cin)).float()
        offs = [(0, 0), (0, 0), (0, -1), (0, 0), (-1, 0), (1, 1), (1, -1), (-1, 1), (-1, -1)]
        oh, ow = offs[t]
        sl = gp[:, 2 + oh:2 + oh + 2 * h:2, 2 + ow:2 + ow + 2 * w:2, :]
        r1 = torch.einsum("nhwk,ck->nhwc", sl, wd1.float()[:, t, :])
        print(f"   tap {t} off {offs[t]}: err {(d1 - r1).abs().max().item():.4f} of {r1.abs().max().item():.3f}")
