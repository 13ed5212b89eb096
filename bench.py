#!/usr/bin/env python
"""Headline benchmark: train images/sec of the hot path on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--workload cfg2|cfg3|cfg1] [--impl reference]

One "step" = one full training step on one synthetic batch per GPU: forward -> loss -> backward ->
gradient all-reduce (N > 1) -> metrics -> grad-norm -> optimizer, i.e. what train_model.py:51-120 does per
batch.  Default workload = BASELINE.json configs[1]: ResNet-50 ImageNet classification, 3x224x224,
batch 256 per GPU, bf16 activations (weak scaling: per-GPU batch fixed).

Prints ONE JSON line (see the keys at the bottom).  `value` is timed with the batch already resident in HBM;
`e2e` repeats the measurement through the public API with HOST (pinned) batches: H2D copy of the batch and
D2H read of the loss inside the timed region.  `roofline` is for the dominant kernel family (the tcgen05
implicit-GEMM convolutions), from CUDA events around every conv launch of one extra instrumented step.
`--impl reference` times the reference's own algorithm on the host CPU (the oracle port of its PyTorch
modules — the reference itself is Python and does not travel to the GPU box), on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: everything any library prints to fd 1 (NCCL's banner, device printf) is
# diverted to stderr for the whole run, and the result line goes to the saved descriptor.
sys.stdout.flush()
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    os.write(_RESULT_FD, (json.dumps(obj) + "\n").encode())

WORKLOADS = {
    # name: (description, per-GPU batch, input shape, cpu-sample batch)
    "cfg2": ("resnet50_imagenet_cls_3x224x224", 256, (3, 224, 224), 16),
    "cfg3": ("resnet50_attention_unet_acdc_4class_3x256x256", 24, (3, 256, 256), 2),
    "cfg1": ("resnet18_attention_unet_covidqu_binary_1x256x256", 8, (1, 256, 256), 8),
    "cfg4": ("basic_unet_idrid_multilabel5_3x1024x1024", 4, (3, 1024, 1024), 1),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def synthetic_batch(workload, batch, gen, torch):
    _, _, shape, _ = WORKLOADS[workload]
    if workload == "cfg2":
        x = torch.randn((batch, *shape), generator=gen)
        y = torch.randint(0, 1000, (batch, 1), generator=gen)
    elif workload == "cfg3":
        x = torch.rand((batch, *shape), generator=gen)
        y = torch.randint(0, 4, (batch, 1, *shape[1:]), generator=gen)
    elif workload == "cfg4":
        x = torch.rand((batch, *shape), generator=gen)
        y = (torch.rand((batch, 5, *shape[1:]), generator=gen) < 0.05).float()
    else:
        x = torch.rand((batch, *shape), generator=gen)
        y = (torch.rand((batch, 1, *shape[1:]), generator=gen) < 0.3).long()
    return x, y


def reference_arm(args):
    """The reference's algorithm on the host CPU: oracle port of its PyTorch modules, all host threads."""
    import torch
    from oracle import ref_losses, ref_models
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, _, _, cpu_batch = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    if args.workload == "cfg2":
        model = ref_models.resnet50_classifier()
        opt = torch.optim.AdamW(model.parameters(), lr=0.004, betas=(0.9, 0.999), weight_decay=0.05)
        loss_fn = lambda p, y: ref_losses.ce_with_softmax(p, y, 0.1)
    elif args.workload == "cfg3":
        model = ref_models.resnet50_attention_unet(out_ch=4, final_activation="softmax")
        opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
        loss_fn = ref_losses.dice_loss
    elif args.workload == "cfg4":
        model = ref_models.basic_unet(out_ch=5, final_activation="sigmoid")
        opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
        loss_fn = ref_losses.bce_loss_torch
    else:
        model = ref_models.resnet18_attention_unet()
        opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
        loss_fn = ref_losses.dice_loss
    ref_models.kaiming_init_(model).train()
    gen = torch.Generator().manual_seed(1)
    x, y = synthetic_batch(args.workload, cpu_batch, gen, torch)

    def step():
        opt.zero_grad()
        loss = loss_fn(model(x), y)
        v = loss.item()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), float("inf"))
        opt.step()
        return v

    steps, warm = min(args.steps, args.ref_steps), min(args.warmup, 1)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = steps * cpu_batch / dt
    sample = f"{steps} steps x batch {cpu_batch} of the {name} training step (fp32, torch {torch.__version__} CPU)"
    emit({
        "impl": "reference", "metric": "train images/sec", "value": val, "unit": "images/sec",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": name, "per_step_batch": cpu_batch},
        "cpu_baseline": {"value": val, "unit": "images/sec", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def cpu_baseline(workload, budget_s=25.0):
    import torch
    from oracle import ref_losses, ref_models
    name, _, _, cpu_batch = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    if workload == "cfg2":
        model, loss_fn = ref_models.resnet50_classifier(), (lambda p, y: ref_losses.ce_with_softmax(p, y, 0.1))
    elif workload == "cfg3":
        model, loss_fn = ref_models.resnet50_attention_unet(out_ch=4, final_activation="softmax"), ref_losses.dice_loss
    elif workload == "cfg4":
        model, loss_fn = ref_models.basic_unet(out_ch=5, final_activation="sigmoid"), ref_losses.bce_loss_torch
    else:
        model, loss_fn = ref_models.resnet18_attention_unet(), ref_losses.dice_loss
    ref_models.kaiming_init_(model).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01)
    x, y = synthetic_batch(workload, cpu_batch, torch.Generator().manual_seed(1), torch)

    def step():
        opt.zero_grad()
        loss = loss_fn(model(x), y)
        loss.item()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), float("inf"))
        opt.step()

    step()
    n, t0 = 0, time.perf_counter()
    while n < 2 or (time.perf_counter() - t0 < budget_s and n < 8):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n * cpu_batch / dt, "unit": "images/sec", "cores": cores, "kind": "port",
            "sample": f"{n} steps x batch {cpu_batch} of the same training step, oracle port of the reference's "
                      f"PyTorch modules, fp32, {cores} threads"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--ref-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small-allreduce", default="p2p", choices=["p2p", "nccl"],
                    help="SyncBN statistic exchange (N>1): one-kernel all-reduce over NVLink peer memory, or NCCL")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every step from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    import medsegpretrainimagenet_b200 as b200
    from medsegpretrainimagenet_b200 import models, ops
    from medsegpretrainimagenet_b200.parallel import GradReducer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
        if args.small_allreduce == "p2p":
            # SyncBN statistics through the peer-memory kernel (csrc/msp_p2p.cu) instead of ~100 tiny NCCL calls / step
            from medsegpretrainimagenet_b200 import parallel as _par
            if _par.enable_peer_allreduce(group) is None:
                args.small_allreduce = "nccl"        # CUDA IPC unavailable on this box: every rank fell back together
    warm = max(args.warmup, 3)

    name, batch, shape, _ = WORKLOADS[args.workload]
    batch = args.batch or batch
    torch.manual_seed(0)
    if args.workload == "cfg2":
        model = models.resnet50_classifier(group=group)
        crit = b200.losses.CrossEntropyLoss(label_smoothing=0.1)
        make_opt = lambda ps: torch.optim.AdamW(ps, lr=0.004, betas=(0.9, 0.999), weight_decay=0.05, fused=True,
                                                capturable=True)
    elif args.workload == "cfg3":
        model = models.resnet50_attention_unet(out_ch=4, final_activation="softmax", group=group)
        crit = b200.losses.DiceLoss(group=group)
        make_opt = lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=1e-4, fused=True)
    elif args.workload == "cfg4":
        model = models.basic_unet(out_ch=5, final_activation="sigmoid", group=group)
        crit = b200.losses.BCELoss(torch_semantics=True)            # torch.nn.BCELoss
        make_opt = lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=1e-4, fused=True)
    else:
        model = models.resnet18_attention_unet(group=group)
        crit = b200.losses.DiceLoss(group=group)
        make_opt = lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=1e-4, fused=True)
    models.kaiming_init_(model)
    model.to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    reducer = GradReducer(params, bucket_mb=32.0, group=group)
    opt = make_opt(params)
    n_params = sum(p.numel() for p in params)

    gen = torch.Generator().manual_seed(1 + rank)
    x_host, y_host = synthetic_batch(args.workload, batch, gen, torch)
    if args.workload == "cfg2":
        x_host = x_host.to(torch.bfloat16)   # BASELINE configs[1]: "3x224x224 bf16" images (x ~ N(0,1) -> bf16)
    x_host, y_host = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    if args.workload == "cfg2":
        top5 = b200.metrics.Top5Accuracy()
        cm = b200.metrics.MultiClassConfusionMatrix(number_of_classes=1000)
    elif args.workload == "cfg3":
        cm = b200.metrics.MultiClassConfusionMatrix(number_of_classes=4)
        top5 = None
    elif args.workload == "cfg4":
        cm = b200.metrics.ConfusionMatrix(None, threshold=0.5)
        top5 = None
    else:
        cm = b200.metrics.ConfusionMatrix(None, threshold=0.5)
        top5 = None

    def step_core(x, y):
        reducer.zero_grad()
        pred = model(x)
        loss = crit(pred, y)
        if args.workload == "cfg2":
            # device-side counters (the .cpu() of the C x C matrix is deferred to evaluate_batch cadence)
            b200.metrics.multiclass_confusion_matrix(pred, y)
            b200.metrics.topk_correct(pred, y, 5)
        elif args.workload == "cfg3":
            b200.metrics.multiclass_confusion_matrix(pred, y)
        elif args.workload == "cfg4":
            b200.metrics.binary_confusion_counts(pred, y, 0.5, per_channel=True)   # multilabel: 4 x (5,) counters
        else:
            b200.metrics.binary_confusion_counts(pred, y, 0.5)
        loss.backward()                              # loss/loss.py:87
        reducer.finish()
        torch.nn.utils.clip_grad_norm_(params, float("inf"), foreach=True)   # train_model.py:95-98
        opt.step()
        return loss.detach()

    graphed = None

    def step(x, y, read_loss=False):
        out = graphed(x, y) if graphed is not None else step_core(x, y)
        return out.item() if read_loss else None     # loss/loss.py:85 (the one host read of the step)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step(x_dev, y_dev)
    barrier()
    if not args.no_graph:
        # the whole step as ONE CUDA graph (graphs.GraphedStep): same kernels, one launch from the host
        graphed = b200.GraphedStep(step_core, (x_dev, y_dev), models=[model], warmup=1)
        x_dev, y_dev = graphed.static_in          # resident inputs ARE the graph's static buffers (no extra copy)
        for _ in range(2):
            step(x_dev, y_dev)
        barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step(x_dev, y_dev)
    e1.record()
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps   # CPU time to ENQUEUE one step (no sync inside)
    barrier()
    launches = ops.launch_count() - l0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end through the public API with host batches --------------------
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    # public host API: double-buffered prefetch — the pinned host batch i+1 crosses PCIe (inside the timed region)
    # while step i runs; every step still consumes a freshly copied batch and reads its loss back
    feeder = b200.BatchPrefetcher((x_host, y_host), dev)
    reader = b200.ScalarReader(depth=1)             # every step's loss is read back; the host waits one step late
    feeder.put(x_host, y_host)
    for i in range(args.steps):
        xb, yb = feeder.get()
        if i + 1 < args.steps:
            feeder.put(x_host, y_host)
        out = graphed(xb, yb) if graphed is not None else step_core(xb, yb)
        got = reader.push(out)                      # D2H copy of this step's loss (pinned, asynchronous)
        last = got if got is not None else last
    last = (reader.drain() or [last])[-1]
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    # the host -> device copy of one batch by itself (explains e2e when the PCIe link, not the step, is the limit)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xb, yb = feeder.bufs[0]
    h0.record()
    for _ in range(3):
        xb.copy_(x_host, non_blocking=True)
        yb.copy_(y_host, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_ms = h0.elapsed_time(h1) / 3

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- instrumented step: device time of every convolution launch -----------------------------
    timeline = []
    ops.set_conv_timeline(timeline)
    step_core(x_dev, y_dev)
    torch.cuda.synchronize()
    ops.set_conv_timeline(None)
    conv_ms = sum(a.elapsed_time(b) for _, _, a, b in timeline)
    conv_flops = sum(f for _, f, _, _ in timeline)
    by_kind = {}
    for kind, f, a, b in timeline:
        d = by_kind.setdefault(kind, [0.0, 0.0, 0])
        d[0] += f; d[1] += a.elapsed_time(b); d[2] += 1
    pk, pk_src = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    achieved_tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0

    if rank == 0:
        total_images = args.steps * batch * world
        out = {
            "metric": "train images/sec", "value": total_images / (ms * 1e-3), "unit": "images/sec",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": name, "per_gpu_batch": batch, "global_batch": batch * world,
                       "parallelism": f"dp{world}", "params": n_params,
                       "optimizer": type(opt).__name__, "l2": "inputs_larger_than_l2 (batch + activations >> 126 MB)",
                       "step": "fwd+loss+bwd+allreduce+metrics+gradnorm+optimizer",
                       "syncbn_exchange": None if world == 1 else ("peer-memory kernel" if args.small_allreduce == "p2p" else "nccl"),
                       "execution": "eager launches" if graphed is None else "one CUDA graph replay per step"},
            "e2e": {"value": total_images / (ms_e2e * 1e-3), "unit": "images/sec",
                    "h2d_bytes_per_step": (x_host.numel() * x_host.element_size()
                                           + y_host.numel() * y_host.element_size()) * world,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps, "last_loss": last,
                    "input_dtype": str(x_host.dtype).replace("torch.", ""), "h2d_ms_alone": h2d_ms,
                    "h2d_gbs_alone": (x_host.numel() * x_host.element_size()
                                      + y_host.numel() * y_host.element_size()) / (h2d_ms * 1e-3) / 1e9},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "tapgemm_kernel / wgrad_kernel (tcgen05 implicit-GEMM conv)",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None, "peak_source": f"{pk_src} (sustained)",
                         "traffic": None,
                         "conv_ms_per_step": conv_ms, "conv_share_of_step": conv_ms / (ms / args.steps),
                         "algorithmic_gflop_per_step": conv_flops / 1e9,
                         "by_kind": {k: {"gflop": v[0] / 1e9, "ms": v[1], "launches": v[2],
                                         "tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else None}
                                     for k, v in by_kind.items()}},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.workload)
        emit(out)
    if world > 1:
        # Tearing down a NCCL communicator whose collectives were captured in a live CUDA graph can block; the
        # measurement is complete and printed: synchronise, rendezvous once more and leave without the teardown.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
