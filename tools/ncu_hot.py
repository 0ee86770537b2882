#!/usr/bin/env python
"""Top stall-sample SASS lines of one launch in an ncu report (source page)."""
import csv, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
print(rows[0][:2])
hdr = rows[1]
si, ss, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > ss]
tot = sum(int(r[ss] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][ss] or 0))[:top]:
    stalls = {h: int(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h and v.isdigit() and int(v) > 0}
    best = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    print(f"{idx:5d} {int(r[ss]):7d} {100*int(r[ss])/max(tot,1):5.1f}%  exec {r[ie]:>8s}  {r[si].strip()[:70]:70s} {best}")
