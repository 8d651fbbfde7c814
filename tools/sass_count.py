#!/usr/bin/env python3
"""Static SASS instruction counts per kernel of an object / shared library (cuobjdump -sass): total and the top opcodes.
usage: python tools/sass_count.py file.sass [name-filter]"""
import re, sys
txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if flt not in name:
        continue
    ins = [l for l in f.split('\n') if re.search(r'/\*[0-9a-f]{4}\*/', l)]
    ops = {}
    for l in ins:
        m = re.search(r'\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m:
            o = m.group(1).split('.')[0]
            ops[o] = ops.get(o, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:12]
    print(name[:64], len(ins), top)
