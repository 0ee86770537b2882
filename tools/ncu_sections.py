#!/usr/bin/env python
"""Stall samples of one launch grouped by code section: SASS rows between landmark instructions
(barriers, mbarrier waits, MMA issue).  usage: ncu_sections.py report.ncu-rep launch_skip"""
import csv, subprocess, sys, re
rep, skip = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
si, ss, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > ss]
tot = sum(int(r[ss] or 0) for r in body)
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
sec_start, acc, inst, stalls, marks = 0, 0, 0, {}, []
def flush(end, label):
    global acc, inst, stalls, marks, sec_start
    if acc:
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        print(f"[{sec_start:5d},{end:5d}) {acc:7d} {100*acc/tot:5.1f}%  warp-inst {inst:>10d}  {' '.join(marks)[:60]:60s} {top}")
    acc, inst, stalls, marks, sec_start = 0, 0, {}, [], end
for idx, r in enumerate(body):
    src = r[si]
    acc += int(r[ss] or 0)
    inst += int(r[ie] or 0)
    for i, h in stall_cols:
        if r[i].isdigit() and int(r[i]):
            stalls[h] = stalls.get(h, 0) + int(r[i])
    m = re.search(r"(UTCHMMA|UTCBAR|LDTM|BAR\.SYNC|MUFU\.\w+|UBLKCP|LDGSTS|SYNCS\.PHASECHK|SHFL|STG|EXIT)", src)
    if m and (not marks or marks[-1] != m.group(1)):
        marks.append(m.group(1))
    if "BAR.SYNC" in src or ("BRA" in src and idx > 0 and "SYNCS.PHASECHK" in body[idx - 1][si]):
        flush(idx + 1, src)
flush(len(body), "end")
print("total", tot)
