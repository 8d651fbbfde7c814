#!/usr/bin/env python3
"""Export the blue-noise sampler tables (data, not code) used by the reference RNG.

Source of the numbers: /root/reference/renderer/util/RandGenData.h:15-39 (the
OPTIMIZED_BLUE_NOISE_SPP == 4 block: Owen-scrambled Sobol 256spp x 256d, the
128x128x8 scrambling tile and the 128x128x8 ranking tile; Heitz et al. 2019).
Output layout (little endian, raw bytes):
    sobol      65536 B
    scrambling 131072 B
    ranking    131072 B
Run once in the build container (the reference tree does not exist on the GPU box).
"""
import re, sys, pathlib
import numpy as np

src = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/renderer/util/RandGenData.h")
dst = pathlib.Path(sys.argv[2] if len(sys.argv) > 2 else
                   "real-time-path-tracing-voxel-blocks_b200/data/bluenoise_tables.bin")
text = src.read_text()
beg = text.index("#if OPTIMIZED_BLUE_NOISE_SPP == 4")
end = text.index("#endif", beg)
block = text[beg:end]
arrays = {}
for m in re.finditer(r"unsigned char (h_\w+)\[([^\]]+)\]\s*=\s*\{([^}]*)\}", block):
    name, body = m.group(1), m.group(3)
    vals = np.array([int(v) for v in body.replace("\n", " ").split(",") if v.strip()], dtype=np.uint8)
    arrays[name] = vals
sob, scr, rnk = arrays["h_sobol_256spp_256d"], arrays["h_scramblingTile"], arrays["h_rankingTile"]
assert sob.size == 65536 and scr.size == 131072 and rnk.size == 131072, (sob.size, scr.size, rnk.size)
dst.parent.mkdir(parents=True, exist_ok=True)
dst.write_bytes(sob.tobytes() + scr.tobytes() + rnk.tobytes())
print("wrote", dst, dst.stat().st_size, "bytes")
