"""Diagnostic: loss trajectories of the cfg2 bench step with the weight-pack cache on/off, eager and graphed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medsegpretrainimagenet_b200 as b200
from medsegpretrainimagenet_b200 import models, ops

dev = torch.device("cuda")
def run(graph, steps=12, batch=32):
    torch.manual_seed(0)
    model = models.resnet50_classifier()
    models.kaiming_init_(model)
    model.to(dev).train()
    params = list(model.parameters())
    opt = torch.optim.AdamW(params, lr=0.004, weight_decay=0.05, fused=True, capturable=True)
    crit = b200.losses.CrossEntropyLoss(label_smoothing=0.1)
    g = torch.Generator().manual_seed(1)
    x = torch.randn((batch, 3, 224, 224), generator=g).to(dev)
    y = torch.randint(0, 1000, (batch, 1), generator=g).to(dev)
    def step(x, y):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        return loss.detach()
    out = []
    fn = step
    for i in range(steps):
        if graph and i == 3:
            fn = b200.GraphedStep(step, (x, y), models=[model], warmup=1)
        torch.manual_seed(100 + i)
        out.append(round(float(fn(x, y)), 3))
    return out
print("cache", ops._PACK_ENABLED, "eager", run(False))
print("cache", ops._PACK_ENABLED, "graph", run(True))
