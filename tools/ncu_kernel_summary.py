"""Per-launch summary of an `ncu -i X.ncu-rep --page raw --csv` dump: duration, tensor-pipe activity, issue activity, MUFU (XU) pipe,
DRAM bytes, registers.  Usage: python tools/ncu_kernel_summary.py raw.csv [label ...] (labels name the launches in order)."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
labels = sys.argv[2:]
hdr = next(r for r in rows if "Kernel Name" in r)
hi = rows.index(hdr)
units = rows[hi + 1]
idx = {h: i for i, h in enumerate(hdr)}


def col(*cands):
    for c in cands:
        for h in hdr:
            if h == c or h.startswith(c):
                return idx[h]
    return None


def num(r, i):
    try:
        return float(r[i].replace(",", ""))
    except Exception:
        return float("nan")


c_dur = col("gpu__time_duration.sum")
c_tensor = col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_tensor")
c_issue = col("smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active")
c_xu = col("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")
c_rd, c_wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
c_regs = col("launch__registers_per_thread")
c_dramp = col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3}
print(f"{'launch':52s} {'us':>8s} {'tensor%':>8s} {'issue%':>7s} {'xu%':>6s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'dram%':>6s} {'regs':>5s}")
n = 0
for r in rows[hi + 2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*$", "", re.sub(r"^void\s+", "", r[idx["Kernel Name"]])).replace("vb::", "")
    lab = labels[n] if n < len(labels) else name[:52]
    n += 1
    dur = num(r, c_dur) * scale.get(units[c_dur], 1.0)
    rd = num(r, c_rd) * scale.get(units[c_rd], 1.0) / 1e6
    wr = num(r, c_wr) * scale.get(units[c_wr], 1.0) / 1e6
    f = lambda c: f"{num(r, c):.1f}" if c is not None else "-"
    print(f"{lab[:52]:52s} {dur:8.1f} {f(c_tensor):>8s} {f(c_issue):>7s} {f(c_xu):>6s} {rd:11.1f} {wr:11.1f} {f(c_dramp):>6s} {f(c_regs):>5s}")
