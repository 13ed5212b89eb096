"""Per-layer roofline of the ResNet-50 (cfg2) convolutions: ideal time = max(flops / TC peak, bytes / HBM peak).
Bytes = each operand read once + output written once (bf16 activations, bf16 packed weights, fp32 dW)."""
import json, sys
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pk = json.load(open("MEASURED_PEAKS.json"))
TC, BW = pk["bf16_tflops_sustained"] * 1e12, pk["hbm_gbs"] * 1e9
layers = []  # (name, H_in, Cin, Cout, k, stride)
layers.append(("stem", 224, 8, 64, 7, 2))
h, cin = 56, 64
for li, (c, n) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3))):
    for b in range(n):
        s = 2 if (b == 0 and li > 0) else 1
        layers.append((f"L{li}b{b}c1", h, cin, c, 1, 1))
        layers.append((f"L{li}b{b}c2", h, c, c, 3, s))
        h = h // s
        layers.append((f"L{li}b{b}c3", h, c, 4 * c, 1, 1))
        cin = 4 * c
tot = {"fprop": [0, 0, 0], "dgrad": [0, 0, 0], "wgrad": [0, 0, 0]}
rows = []
for name, hin, ci, co, k, s in layers:
    ho = hin // s
    M = B * ho * ho
    fl = 2.0 * M * co * ci * k * k
    xin = B * hin * hin * ci * 2
    yout = M * co * 2
    w = co * ci * k * k
    for kind, by in (("fprop", xin + yout + 2 * w), ("dgrad", xin + yout + 2 * w), ("wgrad", xin + yout + 4 * w)):
        if kind == "dgrad" and name == "stem":
            continue
        t = max(fl / TC, by / BW)
        tot[kind][0] += fl; tot[kind][1] += by; tot[kind][2] += t
        rows.append((name, kind, M, co, ci * k * k, fl / 1e9, by / 1e6, t * 1e6, "TC" if fl / TC > by / BW else "HBM"))
if "-v" in sys.argv:
    for r in rows:
        if r[1] == "fprop":
            print("%-10s %-6s M=%-8d N=%-5d K=%-5d %8.1f GF %8.1f MB ideal %7.1f us %s" % r)
for k, v in tot.items():
    print(f"{k}: {v[0]/1e12:.2f} TF, {v[1]/1e9:.2f} GB, ideal {v[2]*1e3:.2f} ms  (pure TC {v[0]/TC*1e3:.2f} ms, pure HBM {v[1]/BW*1e3:.2f} ms)")
print("sum ideal ms:", sum(v[2] for v in tot.values()) * 1e3)
