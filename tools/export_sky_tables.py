#!/usr/bin/env python3
"""Export the Hosek-Wilkie sky-model coefficient tables (data, not code) the reference's sky generator reads.

Source of the numbers: /root/reference/renderer/sky/SkyData.h (skyDataSets :3, skyDataSetsRad :556, hSolarDatasets :629,
hLimbDarkeningDatasets :2442 — the published spectral datasets of Hosek & Wilkie 2012/2013, 10 channels).
Output layout (little-endian float32): skyDataSets[540] skyDataSetsRad[60] hSolarDatasets[1800] hLimbDarkeningDatasets[60].
Run once in the build container (the reference tree does not exist on the GPU box).
"""
import re, sys, pathlib
import numpy as np

src = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/renderer/sky/SkyData.h").read_text()
dst = pathlib.Path(sys.argv[2] if len(sys.argv) > 2 else "real-time-path-tracing-voxel-blocks_b200/data/sky_tables.bin")


def arr(name, n):
    m = re.search(r"static const float %s\[\]\s*=\s*\{(.*?)\};" % name, src, re.S)
    body = re.sub(r"//.*", "", m.group(1))
    vals = np.array([float(t.rstrip("fF")) for t in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?[fF]?", body)], np.float32)
    assert vals.size == n, (name, vals.size)
    return vals


out = np.concatenate([arr("skyDataSets", 540), arr("skyDataSetsRad", 60), arr("hSolarDatasets", 1800), arr("hLimbDarkeningDatasets", 60)])
dst.write_bytes(out.astype("<f4").tobytes())
print("wrote", dst, dst.stat().st_size, "bytes")
