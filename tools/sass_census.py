#!/usr/bin/env python
"""SASS census of the shipped library: per-kernel counts of the instruction mnemonics that prove the tcgen05 / TMEM /
TMA claims (cuobjdump -sass of vit-cnn_b200/csrc/libvitcnn.so).  usage: python tools/sass_census.py > profiles/r02_sass_census.txt"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vit-cnn_b200", "csrc", "libvitcnn.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cols = ["UTCHMMA.2CTA", "UTCHMMA", "UTCHMMA(A=tmem)", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UBLKCP", "HMMA", "MUFU.EX2", "MUFU.TANH"]
counts, order, cur = {}, [], None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter(); order.append(cur)
        continue
    if cur is None: continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*)", line)
    if not m: continue
    op, rest = m.group(1), m.group(2)
    c = counts[cur]
    if op.startswith("UTCHMMA"):
        c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        if rest.lstrip().startswith("tmem["): c["UTCHMMA(A=tmem)"] += 1
    for k in ("LDTM", "STTM", "UTCBAR", "UTMALDG", "UBLKCP", "HMMA"):
        if op.startswith(k): c[k] += 1
    if op.startswith("MUFU.EX2"): c["MUFU.EX2"] += 1
    if op.startswith("MUFU.TANH"): c["MUFU.TANH"] += 1
print("SASS census of vit-cnn_b200/csrc/libvitcnn.so (cuobjdump -sass, sm_100a; static instruction counts per kernel; tools/sass_census.py).")
print("UTCHMMA = tcgen05.mma (UTCHMMA.2CTA = cta_group::2; A=tmem: the A operand is read from tensor memory), LDTM / STTM = tcgen05.ld / tcgen05.st")
print("(TMEM <-> registers), UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (tensor-map TMA load), UBLKCP = cp.async.bulk (1-D bulk copy,")
print("either direction), HMMA = mma.sync (legacy tensor path).\n")
print(f"{'kernel':58s}" + "".join(f"{c:>16s}" for c in cols))
for k in order:
    c = counts[k]
    if not any(c[x] for x in cols): continue
    print(f"{k[:58]:58s}" + "".join(f"{c[x]:>16d}" for x in cols))
print("\nkernels without any of these instructions: " + ", ".join(k for k in order if not any(counts[k][x] for x in cols)))
