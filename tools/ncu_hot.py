#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: hottest SASS lines, stall mix, and instruction share per source line.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv ; python tools/ncu_hot.py src.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] not in ('Address', 'Kernel Name')]
f = lambda r, k: float(r[idx[k]] or 0)
ti = sum(f(r, 'Instructions Executed') for r in data); ts = sum(f(r, '# Samples') for r in data)
tt = sum(f(r, 'Thread Instructions Executed') for r in data)
print("sass rows %d, warp-inst %.3g, thread-inst %.3g (avg %.1f thr/inst), samples %d" % (len(data), ti, tt, tt / max(ti, 1), ts))
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:N]:
    print("%s %5.2f%%smp %5.2f%%inst thr%5s | %s" % (r[idx['Address']][-5:], 100 * f(r, '# Samples') / ts, 100 * f(r, 'Instructions Executed') / ti,
                                                   r[idx['Avg. Threads Executed']], r[idx['Source']][:100]))
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = {h: sum(f(r, h) for r in data) for h in st}; s = sum(tot.values())
print("stalls:", {k: round(100 * v / s, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
