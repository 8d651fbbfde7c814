#!/usr/bin/env python3
"""Debug aid: per-frame, per-plane error statistics of the CUDA denoiser chain vs the oracle (same scenario as
tests/test_gpu_parity.py::test_denoiser_chain_matches_oracle), no assertions."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tests", os.path.join("real-time-path-tracing-voxel-blocks_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import common, oracle as O, vpt, vpt_scenes as S
O.build()
W, H = 256, 160
for firefly in (1, 0):
    inp = common.scene_inputs((2, 1, 2))
    g = common.setup(vpt.Vpt(W, H), inp, spp=1, total=3, diffuse=1)
    o = common.setup(O.Oracle(W, H), inp, spp=1, total=3, diffuse=1)
    p = S.default_denoising_params()
    p["enableFireflyFilter"] = firefly
    cam = common.scene_camera(W, H)
    prev = cam
    for f in range(5):
        g.render(cam, prev, f); o.render(cam, prev, f)
        g.write("Illumination", o.read("Illumination"))
        g.write_reservoirs(f & 1, o.read_reservoirs(f & 1))
        g.denoise(p, cam, prev, f, f + 1); o.denoise(p, cam, prev, f, f + 1)
        line = []
        for name in ("Illumination", "HistoryLength", "PrevIllumination", "PrevFastIllumination", "IlluminationPing", "IlluminationOutput"):
            a, b = g.read(name), o.read(name)
            m, outl, dmax = common.rel_err_stats(a, b)
            line.append("%s %.1e/%.3f/%.2g" % (name[:12], m, outl, dmax))
        print("firefly", firefly, "frame", f, " | ".join(line), flush=True)
        prev = cam
        if f >= 1:
            cam = vpt.camera_set_yaw_pitch(cam, cam[15] + np.float32(0.5 * np.pi / 180.0), cam[16])
