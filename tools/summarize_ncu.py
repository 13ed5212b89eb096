"""Summarise an `ncu --set full` report (tools/ncu_kernels.py): one line per profiled launch with the figures the
roofline claims rest on — duration, DRAM bytes read / written, achieved DRAM GB/s against the measured HBM copy peak,
tensor-pipe activity, registers, shared memory.  Runs where `ncu` is installed (build container or GPU box):

    python tools/summarize_ncu.py gpurun_out/prof_kernels.ncu-rep [--second-only]"""
import csv, io, json, os, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
    "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
]


def main():
    rep = sys.argv[1]
    second_only = "--second-only" in sys.argv
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {c: i for i, c in enumerate(hdr)}
    try:
        hbm = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0

    def val(r, name, scale_to=None):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return float("nan")
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        if scale_to == "bytes":
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u.split("/")[0], 1)
        if scale_to == "us":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(u, 1)
        return v

    print(f"# {os.path.basename(rep)}: ncu --set full --clock-control none (cold caches, serialised launches; shares and "
          f"byte counts are what to read, absolute times are slower than in a step)\n# HBM peak used: {hbm:.0f} GB/s (MEASURED_PEAKS.json copy bandwidth)")
    print(f"{'kernel':44s} {'grid':>6s} {'regs':>4s} {'smem KB':>7s} | {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>6s} {'of HBM':>6s} "
          f"{'dram%':>6s} | {'tensor%':>7s} {'tmem-op%':>8s} {'warps%':>6s}")
    seen = {}
    for r in data:
        name = r[col["Kernel Name"]]
        short = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        us = val(r, "gpu__time_duration.sum", "us")
        rd, wr = val(r, "dram__bytes_read.sum", "bytes"), val(r, "dram__bytes_write.sum", "bytes")
        if rd != rd:   # sections without the read / write split: total from the DRAM byte rate
            i = col.get("dram__bytes.sum.per_second")
            rate = float(r[i].replace(",", "")) * {"byte/s": 1, "kbyte/s": 1e3, "mbyte/s": 1e6, "gbyte/s": 1e9, "tbyte/s": 1e12}.get(units[i].lower().replace("second", "s"), 1)
            rd, wr = rate * us * 1e-6, 0.0
        key = (short, r[col["launch__grid_size"]], r[col["launch__shared_mem_per_block_dynamic"]], round(rd / 4e6))
        seen[key] = seen.get(key, 0) + 1
        if second_only and seen[key] != 2:
            continue
        gbs = (rd + wr) / us / 1e3
        smem = val(r, "launch__shared_mem_per_block_dynamic", "bytes") / 1e3
        print(f"{short[:44]:44s} {int(val(r, 'launch__grid_size')):6d} {int(val(r, 'launch__registers_per_thread')):4d} {smem:7.1f} | "
              f"{us:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {gbs:6.0f} {gbs / hbm:6.2f} "
              f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} | "
              f"{max(val(r, 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'), val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')):7.1f} "
              f"{val(r, 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):8.1f} "
              f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f}")


if __name__ == "__main__":
    main()
