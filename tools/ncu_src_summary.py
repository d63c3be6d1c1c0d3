"""Summarises an `ncu -i X.ncu-rep --page source --csv` dump: stall-reason totals and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def num(x):
    try: return int(float(x))
    except Exception: return 0
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[idx["Address"]].startswith("0x")]
samples = sum(num(r[idx["# Samples"]]) for r in data)
tot = {h: sum(num(r[idx[h]]) for r in data) for h in stalls}
print("kernel rows", len(data), "samples", samples)
for h, v in sorted(tot.items(), key=lambda x: -x[1])[:12]:
    print(f"  {h:28s} {v:8d} {100 * v / max(samples, 1):5.1f}%")
for r in sorted(data, key=lambda r: -num(r[idx["# Samples"]]))[:topn]:
    s = num(r[idx["# Samples"]])
    why = {h[6:]: num(r[idx[h]]) for h in stalls if num(r[idx[h]]) > 0.2 * max(s, 1)}
    print(f"{s:6d} {r[idx['Source']].strip()[:64]:64s} {why}")
