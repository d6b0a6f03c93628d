#!/usr/bin/env python
"""Per-instruction hot spots from an .ncu-rep source page.  usage: ncu_hot.py rep kernel_index [topN] [stallcol]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; ki = int(sys.argv[2]); topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kernels = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kernels.append(cur); continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None and len(r) > 5: cur["rows"].append(r)
K = kernels[ki]; h = K["hdr"]
iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
sc = {c: i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c}
tot = sum(int(r[iSm]) for r in K["rows"])
print(K["name"][:100], "instructions", len(K["rows"]), "samples", tot)
order = sorted(range(len(K["rows"])), key=lambda i: -int(K["rows"][i][iSm]))[:topn]
for i in sorted(order):
    r = K["rows"][i]
    st = sorted(((int(r[j]), c[6:]) for c, j in sc.items() if int(r[j]) > 0), reverse=True)[:3]
    print(f"{i:5d} smp {int(r[iSm]):5d} ({100*int(r[iSm])/tot:4.1f}%) exe {int(r[iE]):8d}  {r[iS].strip()[:70]:70s} {st}")
