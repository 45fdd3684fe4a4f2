#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i`) into the numbers the roofline needs."""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
peaks = json.load(open("MEASURED_PEAKS.json")) if len(sys.argv) < 3 else json.load(open(sys.argv[2]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
def f(k):
    try: return float(m[k].replace(",", ""))
    except Exception: return float("nan")
def scaled(k, to):
    v, unit = f(k), u.get(k, "")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1,
            "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1}.get(unit, 1)
    return v * mult
t = scaled("gpu__time_duration.sum", "s")
rd, wr = scaled("dram__bytes_read.sum", "B"), scaled("dram__bytes_write.sum", "B")
out = {
    "kernel": m.get("Kernel Name", "")[:60], "duration_ms": t * 1e3,
    "dram_read_GB": rd / 1e9, "dram_write_GB": wr / 1e9, "dram_TBps": (rd + wr) / t / 1e12,
    "frac_of_measured_copy_bw": (rd + wr) / t / 1e9 / peaks["hbm_gbs"],
    "dram_throughput_pct_of_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": f("launch__registers_per_thread"), "grid": f("launch__grid_size"),
    "block": f("launch__block_size"),
}
print(json.dumps(out, indent=1))
