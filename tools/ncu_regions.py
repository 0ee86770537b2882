#!/usr/bin/env python
"""Stall samples and executed instructions of one profiled launch of tokens_tm_kernel, grouped by code region.
Regions are cut at landmark instructions found in the SASS (the first LDTM after each hand-off).
usage: ncu_regions.py report.ncu-rep [patches]"""
import csv, re, subprocess, sys, collections
rep = sys.argv[1]
npatch = float(sys.argv[2]) if len(sys.argv) > 2 else 131072.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 10]
ia, isrc, isamp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp] or 0) for r in data)
inst_tot = sum(int(r[ia] or 0) for r in data)
print(f"total samples {tot}, warp instructions {inst_tot} = {inst_tot / npatch:.0f} per patch")
# landmarks in order of appearance: the issuer warp's code starts at the first UTCHMMA, the row threads' at the first LDGSTS
# (stem-input fetch) behind it; code the compiler moved out of line (polling slow paths, a second copy of the issuer's MMAs) sits
# behind the last MUFU.TANH + cls code and is reported as "tail of the listing"
idx = lambda pat, lo=0: [i for i, r in enumerate(data) if re.search(pat, r[isrc]) and i >= lo]
first_mma = idx(r'UTCHMMA')[0]
ldgsts = idx(r'LDGSTS', first_mma)[0]
ex2 = idx(r'MUFU\.EX2', ldgsts)
tanh = idx(r'MUFU\.TANH', ldgsts)
late_mma = [i for i in idx(r'UTCHMMA') if i > tanh[-1]]
end_rows = late_mma[0] - 40 if late_mma else len(data)
regions = [("setup", 0, first_mma - 40), ("issuer", first_mma - 40, ldgsts - 60), ("fetch+fusion+LN1+qkv", ldgsts - 60, ex2[0] - 40),
           ("attention", ex2[0] - 40, tanh[0] - 400), ("proj+LN2", tanh[0] - 400, tanh[0] - 20), ("fc1 GELU", tanh[0] - 20, tanh[-1] + 40),
           ("fc2+LN3+cls", tanh[-1] + 40, end_rows), ("tail of the listing", end_rows, len(data))]
for name, a, b in regions:
    s = collections.Counter(); n = 0; inst = 0
    for r in data[a:b]:
        n += int(r[isamp] or 0); inst += int(r[ia] or 0)
        for i, h in stall:
            if r[i].isdigit(): s[h] += int(r[i])
    top = ', '.join(f'{k} {100 * v / max(n, 1):.0f}%' for k, v in s.most_common(7))
    print(f'{name:22s} [{a:5d},{b:5d}) samples {n:7d} ({100 * n / tot:4.1f}%) inst/patch {inst / npatch:7.0f}  {top}')
