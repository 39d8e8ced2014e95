#!/usr/bin/env python
"""Print selected metrics of an .ncu-rep (read here, no GPU): tools/ncu_metrics.py rep.ncu-rep kernel-substr pat1 pat2 ..."""
import csv, subprocess, sys
rep, ksub, pats = sys.argv[1], sys.argv[2], sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if ksub not in d["Kernel Name"]:
        continue
    print("==", d["Kernel Name"][:70])
    for k in hdr:
        if any(p in k for p in pats) and d[k] not in ("", "n/a"):
            print(f"   {k} = {d[k]} {units[hdr.index(k)]}")
