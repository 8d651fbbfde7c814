#!/usr/bin/env python3
"""Per-kernel times of the last full frame in an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
L = [(r[ix['Kernel Name']].split('(')[0].replace('void ', '').replace('vpt::', ''), float(r[ix['Metric Value']]) / 1e3) for r in rows[hi + 1:] if len(r) == len(hdr)]
gi = [i for i, (n, t) in enumerate(L) if n.startswith('genKernel')]
s, e = gi[-2], gi[-1]
tot = 0
for n, t in L[s:e]:
    print("%-34s %8.1f us" % (n[:34], t)); tot += t
print("frame total %.1f us" % tot)
