"""cuobjdump opcode counts of libvitb200.so per kernel family: the SASS evidence that the hot kernels are Blackwell-native
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add) and where the
remaining mma.sync (HMMA) instructions live.  Usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "vitb200", "libvitb200.so")], capture_output=True, text=True).stdout
ops = ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "MUFU", "FHFMA", "FHADD", "FFMA2", "FMNMX3", "LDG.E.ENL2.256")
fam = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        d = re.sub(r"\(.*$", "", re.sub(r"^void ", "", d)).replace("vb::", "")
        d = re.sub(r"<.*$", "<...>", d) if d.count("<") else d
        cur = fam.setdefault(d, {"n": 0, **{o: 0 for o in ops}})
        cur["n"] += 1
        continue
    if cur is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for o in ops:
            if o == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    cur[o] += 1
            elif op == o or op.startswith(o + "."):
                cur[o] += 1
print("SASS opcode counts per kernel family of vitb200/libvitb200.so (cuobjdump -sass; sm_100a only).  `inst` = instantiations.")
print(f"{'kernel':44s} {'inst':>4s} " + " ".join(f"{o[:12]:>12s}" for o in ops))
tot = {o: 0 for o in ops}
for k, v in fam.items():
    print(f"{k[:44]:44s} {v['n']:4d} " + " ".join(f"{v[o]:12d}" for o in ops))
    for o in ops:
        tot[o] += v[o]
print(f"{'TOTAL':44s} {sum(v['n'] for v in fam.values()):4d} " + " ".join(f"{tot[o]:12d}" for o in ops))
