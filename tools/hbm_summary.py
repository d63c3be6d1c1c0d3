"""Summarises the ncu CSV of tools/hbm_prof.py: per kernel (last launch) duration, DRAM bytes, achieved GB/s and % of the measured copy peak."""
import collections
import csv
import json
import os
import re
import sys

peak = 6539.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rows = collections.OrderedDict()
with open(sys.argv[1], newline="") as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
for r in csv.DictReader(lines):
    k = (r["ID"], re.sub(r"\(.*$", "", re.sub(r"^void\s+", "", r["Kernel Name"])).replace("vb::", ""))
    rows.setdefault(k, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
last = collections.OrderedDict()
for (i, name), m in rows.items():
    if "FillFunctor" in name or "step_inc" in name:
        continue
    last[name] = m
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
print(f"{'kernel':44s} {'us':>8s} {'read MB':>9s} {'write MB':>9s} {'GB/s':>8s} {'% of measured copy peak':>24s} {'ncu dram % of peak':>19s}")
for name, m in last.items():
    t = m["gpu__time_duration.sum"][0] * scale[m["gpu__time_duration.sum"][1]]
    rd = m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
    wr = m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
    gbs = (rd + wr) / t / 1e3
    pct = m.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", (float("nan"), ""))[0]
    print(f"{name[:44]:44s} {t:8.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {gbs:8.0f} {100 * gbs / peak:23.1f}% {pct:18.1f}%")
