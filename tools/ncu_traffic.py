"""Per-kernel DRAM traffic of one benchmark step from an ncu launch list -> profiles/ncu_traffic.json (read by bench.py
for `roofline.traffic` and `roofline.kernels[].dram_bytes_per_launch_ncu`).

GPU box (plain run first, then the same command under ncu; one workload per capture):
    python bench.py --workload cfg2 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --secondary none > gpurun_out/plain.log 2>&1 &&
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/ncu_traffic_cfg2.csv python bench.py --workload cfg2 --steps 1 --warmup 3 --no-graph \
        --no-cpu-baseline --secondary none
Here:
    python tools/ncu_traffic.py cfg2=gpurun_out/ncu_traffic_cfg2.csv cfg3=gpurun_out/ncu_traffic_cfg3.csv

Launches are grouped by kernel name incl. the leading template argument (`tapgemm_kernel<256>`), over ALL launches of
the capture (warm-up, timed and instrumented steps alike: the per-launch mean is what is stored)."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NAMES = {"cfg2": "resnet50_imagenet_cls_3x224x224", "cfg3": "resnet50_attention_unet_acdc_4class_3x256x256",
         "cfg1": "resnet18_attention_unet_covidqu_binary_1x256x256", "cfg4": "basic_unet_idrid_multilabel5_3x1024x1024"}


def family(n):
    n = n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = re.sub(r"^void\s+", "", n)
    m = re.match(r"([\w:]+)(<\(?int\)?(\d+)[^>]*>|<(\d+)[^>]*>)?", n)
    if not m:
        return n[:40]
    base = m.group(1).split("::")[-1]
    num = m.group(3) or m.group(4)
    return f"{base}<{num}>" if num else base


def parse(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    cur = {}
    for r in csv.DictReader(lines):
        key = (r["ID"], r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = (r.get("Metric Unit") or "").lower()
        name = r["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        elif name.startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(unit, 1e-3)
        cur.setdefault(key, {})[name] = v
    for (_, kname), m in cur.items():
        d = per.setdefault(family(kname), {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
        d["launches"] += 1
        d["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        d["us"] += m.get("gpu__time_duration.sum", 0.0)
    return {k: {"launches": d["launches"], "dram_bytes_per_launch": d["dram_bytes"] / d["launches"],
                "us_per_launch_cold": d["us"] / d["launches"]} for k, d in per.items()}


def main():
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        out = json.load(open(out_path))
    except Exception:
        out = {}
    srcs = []
    for arg in sys.argv[1:]:
        wl, path = arg.split("=", 1)
        out[NAMES.get(wl, wl)] = parse(path)
        srcs.append(os.path.basename(path))
    out["_source"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                      "over `bench.py --steps 1 --warmup 3 --no-graph` (" + ", ".join(srcs) + "); per-launch means")
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
    for wl, d in out.items():
        if wl.startswith("_"):
            continue
        print(wl)
        for k, v in sorted(d.items(), key=lambda kv: -kv[1]["dram_bytes_per_launch"] * kv[1]["launches"])[:12]:
            print(f"  {k:40s} {v['launches']:5d} launches  {v['dram_bytes_per_launch'] / 1e6:10.2f} MB/launch  {v['us_per_launch_cold']:8.1f} us")


if __name__ == "__main__":
    main()
