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
def ncu_traffic(workload_name):
    """Per-launch DRAM bytes of each kernel family from the committed ncu capture of this benchmark's step
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from `ncu --metrics dram__bytes_read.sum,
    dram__bytes_write.sum`): the live run cannot measure DRAM traffic itself."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        return d.get(workload_name, {}), d.get("_source")
    except Exception:
        return {}, None


CONV_KERNELS = ("tapgemm", "wgrad_kernel", "narrow_fprop_kernel")


def conv_times_from_profiler(prof_events, timeline):
    """Device time of every conv call from the profiler's kernel records: the conv kernels (tapgemm* / wgrad_kernel) of
    the instrumented step in stream order, dealt out to the calls in call order by their launch counts."""
    import re
    ks = []
    for ev in prof_events:
        if ev.device_type.name != "CUDA":
            continue
        n = ev.name
        if any(k in n for k in CONV_KERNELS):
            ks.append((ev.time_range.start, ev.device_time / 1e3))      # us -> ms
    ks.sort()
    need = sum(t[6] for t in timeline)
    if len(ks) != need:
        return None
    out, i = [], 0
    for kind, f, a, b, nbytes, name, nl in timeline:
        out.append((kind, f, sum(d for _, d in ks[i:i + nl]), nbytes, name))
        i += nl
    return out


def step_kernel_table(prof_events, top=14):
    import collections
    import re
    agg = collections.OrderedDict()
    lib_full = collections.Counter()
    for ev in prof_events:
        if ev.device_type.name != "CUDA":
            continue
        n = ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        n = re.sub(r"^void\s+", "", n)
        if "at::" in n or "nccl" in n or "Memset" in n or "Memcpy" in n:
            lib_full[re.sub(r"\s+", " ", n)[:140]] += 1
        m = re.match(r"([\w:]+)(<[^(]*>)?", n)
        key = (m.group(1).split("::")[-1] + (m.group(2) or ""))[:48] if m else n[:48]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time / 1e3
    tot = sum(v[1] for v in agg.values()) or 1.0
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    return {"kernels": sum(v[0] for v in agg.values()), "device_ms": tot,
            "top": [{"kernel": k, "launches": c, "ms": round(ms, 4), "share": round(ms / tot, 4)} for k, (c, ms) in rows[:top]],
            "library_kernel_names": [[k, c] for k, c in lib_full.most_common(10)],
            "library_kernels_at_or_torch": sum(c for k, (c, ms) in rows if k.startswith(("at::", "vectorized_", "multi_tensor", "elementwise_kernel", "CatArray", "reduce_kernel", "lpnorm")) or "at::native" in k)}


def graph_timeline(path, replay, barrier):
    """Diagnostic (--graph-timeline PREFIX): CUPTI records of ONE replay of the step's CUDA graph — span, time with at
    least one kernel running, idle gaps and what surrounds the largest ones, per-kernel totals as they run inside the
    graph (peer exchanges include their real waiting time here, unlike the eagerly enqueued profile)."""
    import collections
    import re
    import torch
    from torch.profiler import ProfilerActivity, profile
    barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        replay()
        torch.cuda.synchronize()
    evs = []
    for ev in prof.events():
        if ev.device_type.name != "CUDA":
            continue
        n = re.sub(r"^void\s+", "", ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
        m = re.match(r"([\w:]+)(<[^(]*>)?", n)
        key = (m.group(1).split("::")[-1] + (m.group(2) or ""))[:44] if m else n[:44]
        evs.append((ev.time_range.start, ev.time_range.end, key))
    evs.sort()
    if not evs:
        return
    t0, t1 = evs[0][0], max(e[1] for e in evs)
    busy, cur_end, gaps, last_name = 0.0, evs[0][0], [], "(start)"
    for a, b, k in evs:             # union of the kernels' intervals (streams overlap)
        if a > cur_end:
            gaps.append((a - cur_end, last_name, k))
            cur_end = a
        if b > cur_end:
            busy += b - cur_end
            cur_end = b
            last_name = k
    agg = collections.OrderedDict()
    for a, b, k in evs:
        v = agg.setdefault(k, [0, 0.0])
        v[0] += 1
        v[1] += b - a
    with open(path, "w") as f:
        f.write(f"one graph replay: {len(evs)} kernels, span {(t1 - t0) / 1e3:.3f} ms, some kernel running "
                f"{busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms in {len(gaps)} gaps, sum of kernel durations "
                f"{sum(v[1] for v in agg.values()) / 1e3:.3f} ms\n")
        gsum = collections.Counter()
        for g, a, b in gaps:
            gsum[(a, b)] += g
        f.write("idle time by (kernel before -> kernel after):\n")
        for (a, b), g in gsum.most_common(25):
            f.write(f"  {g:9.1f} us  {a} -> {b}\n")
        f.write("kernels:\n")
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"  {k:46s} {c:5d} {us / 1e3:9.3f} ms {us / c:9.1f} us\n")


def conv_roofline(timeline, ms_per_step, workload_name):
    """`roofline` object: the dominant kernel (largest share of device time among the conv kernel variants, which are
    60 % of the step) with its algorithmic FLOPs or bytes over its CUDA-event time, plus the per-kernel list."""
    pk, pk_src = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    hbm = pk.get("hbm_gbs")
    traffic, traffic_src = ncu_traffic(workload_name)
    fam = {}
    by_kind = {}
    for kind, f, t, nbytes, name in timeline:
        d = fam.setdefault(name or "conv", {"launches": 0, "ms": 0.0, "gflop": 0.0, "gbyte": 0.0, "ideal_ms": 0.0})
        d["launches"] += 1; d["ms"] += t; d["gflop"] += f / 1e9; d["gbyte"] += nbytes / 1e9
        d["ideal_ms"] += max(f / (peak_tf * 1e12), nbytes / (hbm * 1e9)) * 1e3
        k = by_kind.setdefault(kind, [0.0, 0.0, 0])
        k[0] += f; k[1] += t; k[2] += 1
    kernels = []
    for name, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        tensor_ms, hbm_ms = d["gflop"] / peak_tf, d["gbyte"] / hbm * 1e3      # GFLOP / (TFLOP/s) = ms; GB / (GB/s) = s
        bound = "tensor" if tensor_ms >= hbm_ms else "hbm"
        e = {"kernel": name, "launches": d["launches"], "ms": d["ms"], "share_of_step": d["ms"] / ms_per_step,
             "algorithmic_gflop": d["gflop"], "algorithmic_gbyte": d["gbyte"],
             "tflops": d["gflop"] / d["ms"] if d["ms"] > 0 else None,
             "gbs": d["gbyte"] / d["ms"] * 1e3 if d["ms"] > 0 else None,
             "bound": bound, "frac_of_tensor_peak": (d["gflop"] / d["ms"]) / peak_tf if d["ms"] > 0 else None,
             "frac_of_hbm_peak": (d["gbyte"] / d["ms"] * 1e3) / hbm if d["ms"] > 0 else None,
             "frac_of_per_launch_roofline": d["ideal_ms"] / d["ms"] if d["ms"] > 0 else None}
        tr = traffic.get(name)
        if tr:
            e["dram_bytes_per_launch_ncu"] = tr.get("dram_bytes_per_launch")
            e["algorithmic_bytes_per_launch"] = d["gbyte"] * 1e9 / d["launches"]
        kernels.append(e)
    conv_ms = sum(d["ms"] for d in fam.values())
    conv_gflop = sum(d["gflop"] for d in fam.values())
    dom = kernels[0] if kernels else None
    out = {"bound": None, "kernel": None, "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None}
    if dom:
        if dom["bound"] == "tensor":
            out.update(bound="tensor", achieved=dom["tflops"], peak=peak_tf, unit="TFLOP/s", frac=dom["frac_of_tensor_peak"])
        else:
            out.update(bound="hbm", achieved=dom["gbs"], peak=hbm, unit="GB/s", frac=dom["frac_of_hbm_peak"])
        out["kernel"] = f"{dom['kernel']} (tcgen05 implicit-GEMM conv; dominant kernel: {dom['share_of_step']:.0%} of the step)"
        out["traffic"] = dom.get("dram_bytes_per_launch_ncu")
        out["traffic_source"] = traffic_src if out["traffic"] is not None else None
        out["algorithmic_per_launch"] = {"gflop": dom["algorithmic_gflop"] / dom["launches"],
                                         "bytes": dom["algorithmic_gbyte"] * 1e9 / dom["launches"]}
        out["avg_launch_us"] = dom["ms"] * 1e3 / dom["launches"]
    out["peak_source"] = f"{pk_src} (bf16 sustained {peak_tf} TFLOP/s, HBM copy {hbm} GB/s)"
    all_tf = conv_gflop / conv_ms if conv_ms > 0 else 0.0
    out["all_conv_kernels"] = {"achieved_tflops": all_tf, "frac_of_tensor_peak": all_tf / peak_tf if peak_tf else None,
                               "frac_of_per_launch_roofline": (sum(d["ideal_ms"] for d in fam.values()) / conv_ms) if conv_ms else None,
                               "conv_ms_per_step": conv_ms, "conv_share_of_step": conv_ms / ms_per_step,
                               "algorithmic_gflop_per_step": conv_gflop}
    out["by_kind"] = {k: {"gflop": v[0] / 1e9, "ms": v[1], "launches": v[2],
                          "tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else None} for k, v in by_kind.items()}
    out["kernels"] = kernels
    return out


def run_workload(args, workload, steps, warm, world, rank, local, dev, group, with_cpu_baseline, sample_clocks=True):
    """One workload on this rank -> the JSON fields (rank 0) or None."""
    import torch
    import torch.distributed as dist
    import medsegpretrainimagenet_b200 as b200
    from medsegpretrainimagenet_b200 import models, ops, optim as mopt
    from medsegpretrainimagenet_b200.parallel import GradReducer

    name, batch, shape, _ = WORKLOADS[workload]
    batch = (args.batch or batch) if workload == args.workload else batch
    torch.manual_seed(0)
    torch_opt = args.optimizer == "torch"
    # the reference's YAMLs: use_deterministic_algorithms true for every downstream U-Net config
    # (config/downstream/acdc/resnet50_attention_unet.yaml:233), false for pretraining (pretraining/resnet50/simple.yaml:99)
    det = (workload != "cfg2") if args.deterministic == "yaml" else (args.deterministic == "on")
    torch.use_deterministic_algorithms(det, warn_only=True)
    torch.utils.deterministic.fill_uninitialized_memory = False    # torch.empty stays uninitialised (no fill kernels)
    if workload == "cfg2":
        # the pretraining YAML's own model: the sequential [DeepResNet, AdaptiveAvgPool2d, Flatten, Linear]
        # (config/pretraining/resnet50/simple.yaml:23-33), executed as one pass of the B200 interpreter
        model = models.resnet50_pretraining_model(group=group)
        crit = b200.losses.CrossEntropyLoss(label_smoothing=0.1)
        if torch_opt:
            make_opt = lambda ps: torch.optim.AdamW(ps, lr=0.004, betas=(0.9, 0.999), weight_decay=0.05, fused=True, capturable=True)
        else:
            make_opt = lambda ps: mopt.AdamW(ps, lr=0.004, betas=(0.9, 0.999), weight_decay=0.05)
    else:
        if workload == "cfg3":
            model = models.resnet50_attention_unet(out_ch=4, final_activation="softmax", group=group)
            crit = b200.losses.DiceLoss(group=group)
        elif workload == "cfg4":
            model = models.basic_unet(out_ch=5, final_activation="sigmoid", group=group)
            crit = b200.losses.BCELoss(torch_semantics=True)            # torch.nn.BCELoss
        else:
            model = models.resnet18_attention_unet(group=group)
            crit = b200.losses.DiceLoss(group=group)
        if torch_opt:
            make_opt = lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=1e-4, fused=True)
        else:
            make_opt = lambda ps: mopt.SGD(ps, lr=0.05, momentum=0.9, weight_decay=1e-4)
    models.kaiming_init_(model)
    model.to(dev).train()
    ops.set_wgrad_stream(torch.cuda.Stream(device=dev) if args.wgrad_stream else None)
    params = [p for p in model.parameters() if p.requires_grad]
    reducer = GradReducer(params, bucket_mb=32.0, group=group)
    opt = make_opt(params)
    n_params = sum(p.numel() for p in params)
    clip = (lambda: torch.nn.utils.clip_grad_norm_(params, float("inf"), foreach=True)) if torch_opt \
        else (lambda: mopt.clip_grad_norm_(params, float("inf")))

    gen = torch.Generator().manual_seed(1 + rank)
    x_host, y_host = synthetic_batch(workload, batch, gen, torch)
    if workload == "cfg2":
        x_host = x_host.to(torch.bfloat16)   # BASELINE configs[1]: "3x224x224 bf16" images (x ~ N(0,1) -> bf16)
    x_host, y_host = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step_core(x, y):
        reducer.zero_grad()
        pred = model(x)
        loss = crit(pred, y)
        if workload == "cfg2":
            # device-side counters (the .cpu() of the C x C matrix is deferred to evaluate_batch cadence)
            b200.metrics.multiclass_confusion_matrix(pred, y)
            b200.metrics.topk_correct(pred, y, 5)
        elif workload == "cfg3":
            b200.metrics.multiclass_confusion_matrix(pred, y)
        elif workload == "cfg4":
            b200.metrics.binary_confusion_counts(pred, y, 0.5, per_channel=True)   # multilabel: 4 x (5,) counters
        else:
            b200.metrics.binary_confusion_counts(pred, y, 0.5)
        loss.backward()                              # loss/loss.py:87
        reducer.finish()
        clip()                                       # train_model.py:95-98 (max_norm = inf: measured and logged)
        opt.step()                                   # train_model.py:107
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
    if rank == 0 and sample_clocks:
        sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step(x_dev, y_dev)
    e1.record()
    host_ms = (time.perf_counter() - t_host0) * 1e3 / steps   # CPU time to ENQUEUE one step (no sync inside)
    barrier()
    launches = ops.launch_count() - l0
    if graphed is not None:
        launches = graphed.launches_per_replay * steps
    ms = e0.elapsed_time(e1)
    clocks = None      # the sampler keeps running through the end-to-end timed region below (both are under load)

    # ---- timed region 2: end to end through the public API with host batches --------------------
    # public host API: double-buffered prefetch — the pinned host batch i+1 crosses PCIe (inside the timed region)
    # while step i runs; every step still consumes a freshly copied batch and reads its loss back.  The feeder's two
    # device buffers and the reader's pinned scalar are allocated BEFORE the timed region (one-off set-up: a cudaMalloc
    # of 2 x 77 MB inside it cost one run 6 ms per step over 20 steps)
    feeder = b200.BatchPrefetcher((x_host, y_host), dev)
    reader = b200.ScalarReader(depth=1)             # every step's loss is read back; the host waits one step late
    if graphed is not None:
        graphed.after_copy = feeder.release         # the feeder's buffer is free once it sits in the graph's static inputs
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    feeder.put(x_host, y_host)
    for i in range(steps):
        xb, yb = feeder.get()
        if i + 1 < steps:
            feeder.put(x_host, y_host)
        out = graphed(xb, yb) if graphed is not None else step_core(xb, yb)
        got = reader.push(out)                      # D2H copy of this step's loss (pinned, asynchronous)
        last = got if got is not None else last
    last = (reader.drain() or [last])[-1]
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    if sample_clocks:
        # nvidia-smi needs ~0.1-0.3 s to deliver its first sample: if the two timed regions were shorter than that,
        # keep the same step running (untimed, every rank the same count) until samples under load exist
        extra = 0
        if rank == 0 and len(sampler.lines) < 2 and sampler.proc is not None:
            extra = int(min(200, max(1, 700.0 / max(ms / steps, 0.05))))
        if world > 1:
            te = torch.tensor([extra], dtype=torch.int64, device=dev)
            dist.broadcast(te, src=0)
            extra = int(te.item())
        for _ in range(extra):
            step(x_dev, y_dev)
        barrier()
    if rank == 0 and sample_clocks:
        clocks = sampler.stop()
        clocks["sampled_over"] = "both timed regions" + (f" + {extra} extra untimed steps" if extra else "")

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

    if graphed is not None:
        graphed.after_copy = None
    if args.graph_timeline and graphed is not None:
        graph_timeline(args.graph_timeline + (f".{workload}.rank{rank}.txt"), lambda: step(x_dev, y_dev), barrier)
        graphed.after_copy = feeder.release

        def e2e_iters():        # six iterations of the end-to-end loop (steady state: the host runs ahead of the device)
            feeder.put(x_host, y_host)
            for i in range(6):
                xb_, yb_ = feeder.get()
                if i + 1 < 6:
                    feeder.put(x_host, y_host)
                reader.push(graphed(xb_, yb_))
            reader.drain()
        graph_timeline(args.graph_timeline + (f".{workload}.rank{rank}.e2e.txt"), e2e_iters, barrier)

    # ---- instrumented step: device time of every convolution launch -----------------------------
    # per-kernel device times from the profiler (CUPTI activity records) of one eagerly enqueued step; if the profiler is
    # unavailable, CUDA-event pairs around every conv call (which include the host's launch gaps at small batch)
    timeline, prof_events, step_kernels, timing_src = [], None, None, "cuda events around every conv call"
    ops.set_conv_timeline(timeline)
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step_core(x_dev, y_dev)
            torch.cuda.synchronize()
        prof_events = list(prof.events())
    except Exception as e:   # noqa: BLE001
        print(f"[bench] profiler unavailable ({e}); falling back to event pairs", file=sys.stderr)
        timeline.clear()
        step_core(x_dev, y_dev)
        torch.cuda.synchronize()
    ops.set_conv_timeline(None)
    timed = conv_times_from_profiler(prof_events, timeline) if prof_events else None
    if timed is not None:
        timing_src = "profiler (CUPTI) per-kernel device time of one eagerly enqueued step"
        step_kernels = step_kernel_table(prof_events)
    else:
        timed = [(kind, f, a.elapsed_time(b), nbytes, name) for kind, f, a, b, nbytes, name, _ in timeline]
    timeline = timed

    out = None
    if rank == 0:
        total_images = steps * batch * world
        h2d_bytes = x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size()
        out = {
            "metric": "train images/sec", "value": total_images / (ms * 1e-3), "unit": "images/sec",
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": name, "per_gpu_batch": batch, "global_batch": batch * world,
                       "parallelism": f"dp{world}", "params": n_params,
                       "optimizer": f"{type(opt).__module__}.{type(opt).__name__}",
                       "l2": "inputs_larger_than_l2 (batch + activations >> 126 MB)",
                       "step": "fwd+loss+bwd+allreduce+metrics+gradnorm+optimizer",
                       "syncbn_exchange": None if world == 1 else ("peer-memory kernel" if args.small_allreduce == "p2p" else "nccl"),
                       "execution": "eager launches" if graphed is None else "one CUDA graph replay per step",
                       "wgrad_stream": bool(args.wgrad_stream), "deterministic_reductions": det},
            "e2e": {"value": total_images / (ms_e2e * 1e-3), "unit": "images/sec",
                    "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / steps, "last_loss": last,
                    "input_dtype": str(x_host.dtype).replace("torch.", ""), "h2d_ms_alone": h2d_ms,
                    "h2d_gbs_alone": h2d_bytes / (h2d_ms * 1e-3) / 1e9},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms,
            "clocks": clocks,
            "roofline": conv_roofline(timeline, ms / steps, name),
        }
        out["roofline"]["kernel_time_source"] = timing_src
        if step_kernels is not None:
            out["step_kernels"] = step_kernels
        if with_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(workload)
    # release this workload's graph, pools and model before the next one
    ops.set_wgrad_stream(None)
    del graphed, feeder, reader, timeline, opt, reducer, model, params
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--secondary", default="auto", choices=["auto", "none", *WORKLOADS],
                    help="second workload measured in the same run and reported under `secondary` (auto: cfg3, the R50 "
                         "attention U-Net the north-star targets are stated on, when the primary is cfg2)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--ref-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small-allreduce", default="p2p", choices=["p2p", "nccl"],
                    help="SyncBN statistic exchange (N>1): one-kernel all-reduce over NVLink peer memory, or NCCL")
    ap.add_argument("--optimizer", default="b200", choices=["b200", "torch"],
                    help="gradient norm + optimizer step: this repo's multi-tensor kernels (default) or torch.optim")
    ap.add_argument("--wgrad-stream", type=int, default=int(os.environ.get("MSP_WGRAD_STREAM", "1")),
                    help="1: weight-gradient kernels on a side stream (graph branch) next to the dgrad / BN-backward chain")
    ap.add_argument("--deterministic", default="yaml", choices=["yaml", "on", "off"],
                    help="fixed-order reductions (torch.use_deterministic_algorithms): as the workload's reference YAML says, or forced")
    ap.add_argument("--graph-timeline", default="", help="diagnostic: write PREFIX.<workload>.rank<r>.txt with the CUPTI timeline "
                                                          "summary of one graph replay")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every step from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the N-rank == 1-rank pre-check (N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    parity = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
        if args.small_allreduce == "p2p":
            # SyncBN statistics through the peer-memory kernel (csrc/msp_p2p.cu) instead of ~100 tiny NCCL calls / step
            from medsegpretrainimagenet_b200 import parallel as _par
            if _par.enable_peer_allreduce(group) is None:
                args.small_allreduce = "nccl"        # CUDA IPC unavailable on this box: every rank fell back together
        if not args.no_parity_check:
            # SURVEY.md §4 tier 6: the N-rank step (SyncBN, global Dice, averaged gradients) == the single-device
            # step on the concatenated batch, checked on these very GPUs before anything is timed
            from medsegpretrainimagenet_b200.selfcheck import exchange_parity, n_rank_parity
            parity = n_rank_parity(group, dev)
            parity["exchange"] = exchange_parity(group, dev)
            parity["ok"] = bool(parity["ok"] and parity["exchange"]["ok"])
    warm = max(args.warmup, 3)

    out = run_workload(args, args.workload, args.steps, warm, world, rank, local, dev, group,
                       with_cpu_baseline=(world == 1 and not args.no_cpu_baseline))
    sec = args.secondary
    if sec == "auto":
        sec = "cfg3" if args.workload == "cfg2" else "none"
    if sec != "none" and sec != args.workload:
        s = run_workload(args, sec, args.steps, warm, world, rank, local, dev, group, with_cpu_baseline=False,
                         sample_clocks=False)
        if rank == 0:
            out["secondary"] = {k: s[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                                                  "config", "e2e", "gpu_launches", "host_enqueue_ms_per_step", "roofline",
                                                  "step_kernels") if k in s}
    if rank == 0:
        if parity is not None:
            out["parity_n"] = "ok" if parity["ok"] else "FAILED"
            out["parity_n_detail"] = parity
        emit(out)
    if world > 1:
        # every CUDA graph that captured NCCL work was released inside run_workload; tear the communicator down
        # properly.  A watchdog ends the process if the teardown blocks (the measurement is complete and printed).
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        dog = threading.Timer(60.0, lambda: os._exit(0))
        dog.daemon = True
        dog.start()
        dist.destroy_process_group()
        dog.cancel()


if __name__ == "__main__":
    main()
