"""Per-kernel device time of ONE eager training step (torch.profiler / CUPTI, warm caches, real overlap-free stream order).
    python tools/profile_step.py [--workload cfg2] [--batch 256]
"""
import argparse, collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medsegpretrainimagenet_b200 as b200
from medsegpretrainimagenet_b200 import models
from medsegpretrainimagenet_b200.parallel import GradReducer

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--batch", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda")
torch.manual_seed(0)
if args.workload == "cfg2":
    model, batch = models.resnet50_classifier(), args.batch or 256
    crit = b200.losses.CrossEntropyLoss(label_smoothing=0.1)
    x = torch.randn((batch, 3, 224, 224), device=dev); y = torch.randint(0, 1000, (batch, 1), device=dev)
    mk = lambda ps: torch.optim.AdamW(ps, lr=0.004, weight_decay=0.05, fused=True)
else:
    model, batch = models.resnet50_attention_unet(out_ch=4, final_activation="softmax"), args.batch or 24
    crit = b200.losses.DiceLoss()
    x = torch.rand((batch, 3, 256, 256), device=dev); y = torch.randint(0, 4, (batch, 1, 256, 256), device=dev)
    mk = lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=1e-4, fused=True)
models.kaiming_init_(model); model.to(dev).train()
params = [p for p in model.parameters() if p.requires_grad]
red = GradReducer(params); opt = mk(params)
def step():
    red.zero_grad(); pred = model(x); loss = crit(pred, y)
    b200.metrics.multiclass_confusion_matrix(pred, y)
    loss.backward(); red.finish(); torch.nn.utils.clip_grad_norm_(params, float("inf"), foreach=True); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type.name != "CUDA": continue
    n = ev.name
    n = n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = re.sub(r"^void\s+", "", n)
    m = re.match(r"([\w:]+)(<[^(]*>)?", n)
    key = (m.group(1).split("::")[-1] + (m.group(2) or ""))[:60] if m else n[:60]
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"{args.workload} batch {batch}: {sum(v[0] for v in agg.values())} kernels, {tot / 1e3:.2f} ms device time in one step")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{k:60s} {c:5d} {us / 1e3:9.3f} ms {us / tot:6.1%} {us / c:9.1f} us")
