#!/usr/bin/env python
"""Top stalled SASS instructions of a profiled kernel (ncu --set full --import-source on), read here without a GPU.
Usage: python tools/ncu_hot.py gpurun_out/prof.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
tot = sum(int(r[iS]) for r in data)
print(rows[0][1][:100], 'total samples', tot, 'instructions executed', sum(int(r[iEx]) for r in data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
top = sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:N]
for k in sorted(top):
    r = data[k]
    st = sorted(((hdr[i], int(r[i])) for i in stall_cols if int(r[i]) > 0), key=lambda x: -x[1])[:2]
    print(k, r[iSrc].strip()[:72], r[iS], r[iEx], st)
