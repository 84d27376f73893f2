#!/usr/bin/env python
"""Per-role stall breakdown of the warp-specialised logmel kernel from an .ncu-rep (source page).
Usage: ncu_roles.py rep ntiles.  Roles are delimited by the USETMAXREG instructions in SASS order."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; nt = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); h = rows[1]; data = rows[2:]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
cut = [i for i, r in enumerate(data) if "USETMAXREG" in r[iS]]
names = ["setup", "fft", "mel", "L", "K+tail"]
bounds = [0] + cut[:4] + [len(data)] if len(cut) >= 4 else [0, len(data)]
tot = sum(float(r[iN] or 0) for r in data)
print(f"role boundaries at SASS rows {cut}; total samples {tot:.0f}")
for k in range(len(bounds) - 1):
    seg = data[bounds[k]:bounds[k + 1]]
    s = sum(float(r[iN] or 0) for r in seg); e = sum(float(r[iE] or 0) for r in seg) / nt
    agg = collections.Counter()
    for r in seg:
        for c in stall:
            agg[c] += float(r[h.index(c)] or 0)
    top = ", ".join(f"{c[6:]} {100 * v / max(s, 1):.0f}%" for c, v in agg.most_common(6))
    print(f"{names[k] if k < len(names) else k:8s} samples {100 * s / tot:5.1f}%  instr/tile {e:8.1f}   {top}")
print("-- top sampled instructions")
L = sorted(((float(r[iN] or 0), i, r[iS].strip()[:64], float(r[iE] or 0) / nt) for i, r in enumerate(data)), reverse=True)[:16]
for s, i, t, e in L:
    print(f"{100 * s / tot:5.1f}%  row {i:5d}  exec/tile {e:7.1f}  {t}")
