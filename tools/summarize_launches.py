"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family."""
import csv, re, sys, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], v * scale))
def family(n):
    n = n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = re.sub(r"^void\s+", "", n)
    m = re.match(r"([\w:]+)(<\d+>)?", n)
    return (m.group(1).split("::")[-1] + (m.group(2) or "")) if m else n[:40]
agg = collections.OrderedDict()
for n, us in rows:
    k = family(n)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot/1e3:.2f} ms total device time (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':58s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} {c:8d} {us/1e3:10.3f} {us/tot:7.1%} {us/c:9.1f}")
