#!/usr/bin/env python3
"""Attribute an ncu SASS profile to CUDA source lines of the kernel body.
usage: tools/ncu_lines.py <report.ncu-rep> <object.o|.so> <kernel regex> [launch index] [top N]
Joins `ncu --page source --print-source sass --csv` (samples / instructions per SASS row) with `nvdisasm -gi` (inline call
chains, needs -lineinfo) of the same binary by instruction index, and sums per OUTERMOST line (the line in the kernel that
the inlined helpers were called from) and per innermost function line."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, obj, kre = sys.argv[1], sys.argv[2], sys.argv[3]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = starts[min(launch, len(starts) - 1)]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == 'Address':
        break
    data.append(r)
# ncu prints every row twice in some versions: dedupe by address
seen, d2 = set(), []
for r in data:
    if r[0] in seen: continue
    seen.add(r[0]); d2.append(r)
data = d2
name = None
for r in rows[:hi][::-1]:
    if r and r[0] == 'Kernel Name': name = r[1]; break
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
lines = []
for cub in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    chain, fn, fresh = [], None, True
    for l in txt.splitlines():
        m = re.match(r'^(_Z\w+):$', l)
        if m: fn = m.group(1); chain = []; fresh = True; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh: chain = []; fresh = False          # first annotation after an instruction starts a new chain (innermost first)
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m and fn:
            lines.append((fn, int(m.group(1), 16), tuple(chain), m.group(2)))
            fresh = True                                  # instructions without an annotation inherit the previous chain
# pick the function whose demangled name matches the kernel regex and whose length matches
byfn = collections.defaultdict(list)
for fn, off, ch, txt in lines: byfn[fn].append((off, ch, txt))
cands = [fn for fn in byfn if re.search(kre, fn) and len(byfn[fn]) == len(data)]
if not cands:
    cands = [fn for fn in byfn if re.search(kre, fn)]
    print("warning: no function with %d instructions; candidates %s" % (len(data), [(c, len(byfn[c])) for c in cands]))
fn = cands[min(launch, len(cands) - 1)] if len(cands) > 1 and len(set(len(byfn[c]) for c in cands)) == 1 else cands[0]
ins = byfn[fn]
f = lambda r, k: float(r[idx[k]] or 0)
ts = sum(f(r, '# Samples') for r in data); ti = sum(f(r, 'Instructions Executed') for r in data)
outer, inner = collections.Counter(), collections.Counter()
outer_i, inner_i = collections.Counter(), collections.Counter()
for (off, ch, txt), r in zip(ins, data):
    o = ch[-1] if ch else ("?", 0); i = ch[0] if ch else ("?", 0)
    outer[o] += f(r, '# Samples'); inner[i] += f(r, '# Samples')
    outer_i[o] += f(r, 'Instructions Executed'); inner_i[i] += f(r, 'Instructions Executed')
print("%s: %d SASS rows, %.3g warp-instructions, %d samples" % (fn, len(data), ti, ts))
src = {}
def text(file, line):
    for d in ("real-time-path-tracing-voxel-blocks_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, file)
        if os.path.exists(p):
            if p not in src: src[p] = open(p).read().splitlines()
            return src[p][line - 1].strip()[:90] if 0 < line <= len(src[p]) else ""
    return ""
print("-- by kernel-body line (outermost call site): %samples %instructions")
for k, v in outer.most_common(top):
    print("%5.1f%% %5.1f%%  %s:%d  %s" % (100 * v / ts, 100 * outer_i[k] / ti, k[0], k[1], text(*k)))
print("-- by innermost line")
for k, v in inner.most_common(top):
    print("%5.1f%% %5.1f%%  %s:%d  %s" % (100 * v / ts, 100 * inner_i[k] / ti, k[0], k[1], text(*k)))
