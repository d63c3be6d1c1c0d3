"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares of ONE training step.

The step is cut out of the list as the launches between two consecutive `adam_kernel` launches (the last kernel of a step).
Usage: python tools/launch_summary.py gpurun_out/launches_r1d.csv > profiles/r1d_launches_summary.txt"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)          # drop the argument list
    return name.replace("vb::", "")[:66]


def main(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            rows.append((r["Kernel Name"], v / 1e3 if r["Metric Unit"] == "ns" else v))
    ends = [i for i, (n, _) in enumerate(rows) if "adam_kernel" in n]
    if len(ends) >= 2:
        step = rows[ends[-2] + 1:ends[-1] + 1]
    else:
        step = rows
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for n, v in step:
        k = short(n)
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + v)
    print(f"launches in the step: {len(step)}; sum of kernel durations {tot / 1e3:.2f} ms")
    print(f"{'kernel':66s} {'n':>4s} {'us':>10s} {'share':>7s}")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:66s} {c:4d} {t:10.1f} {100 * t / tot:6.1f}%")
    g = sum(t for k, (c, t) in agg.items() if k.startswith("gemm_kernel"))
    print(f"tcgen05 GEMM kernel share of the step (sum over instantiations): {100 * g / tot:.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
