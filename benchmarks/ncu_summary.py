#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_*` launch list."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault(r[idi], {"k": r[ki][:46]})[r[mi]] = float(r[vi].replace(",", ""))
print(f"{'id':>4} {'kernel':46} {'ms':>8} {'rd MB':>9} {'wr MB':>9} {'TB/s':>6}")
for k, v in d.items():
    t = v.get("gpu__time_duration.sum", 0) / 1e6
    rd, wr = v.get("dram__bytes_read.sum", 0) / 1e6, v.get("dram__bytes_write.sum", 0) / 1e6
    print(f"{k:>4} {v['k']:46} {t:8.3f} {rd:9.1f} {wr:9.1f} {(rd + wr) / 1e3 / max(t, 1e-9):6.2f}")
