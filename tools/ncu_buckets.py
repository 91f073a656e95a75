#!/usr/bin/env python
"""Instruction / stall-sample share per source line of a profiled kernel (ncu --set full --import-source on; needs -lineinfo).
Usage: python tools/ncu_buckets.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"] if False else
                     ["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
tot_ex = sum(int(r[iEx]) for r in data); tot_s = sum(int(r[iS]) for r in data)
print('instructions', tot_ex, 'samples', tot_s)
# contiguous regions with similar execution counts = loops
regions = []
cur = None
for k, r in enumerate(data):
    ex = int(r[iEx])
    if cur and ex > 0 and 0.8 < ex / max(cur['ex0'], 1) < 1.25:
        cur['hi'] = k; cur['ex'] += ex; cur['s'] += int(r[iS])
    else:
        if cur: regions.append(cur)
        cur = dict(lo=k, hi=k, ex0=ex, ex=ex, s=int(r[iS]))
regions.append(cur)
regions.sort(key=lambda x: -x['ex'])
for g in regions[:N]:
    print('%5d-%5d  n=%3d  per-instr %9d  instr %5.1f%%  samples %5.1f%%   %s' % (
        g['lo'], g['hi'], g['hi'] - g['lo'] + 1, g['ex0'], 100.0 * g['ex'] / tot_ex, 100.0 * g['s'] / tot_s,
        data[g['lo']][iSrc].strip()[:50]))
