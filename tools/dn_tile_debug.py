#!/usr/bin/env python3
"""Bring-up check of the shared-memory tile kernels, one kernel family per process (a CUDA fault is sticky): runs a few frames
of render + denoise with VPT_DN_TILE_MASK = 4 / 1 / 2 / 7 and prints the sync status, the TMA timeout counter and the distance to
the gather-kernel result. usage: python tools/dn_tile_debug.py [mask]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    mask = sys.argv[1]
    os.environ["VPT_DN_TILE_MASK"] = mask
    for p in ("tests", os.path.join("real-time-path-tracing-voxel-blocks_b200", "python")):
        sys.path.insert(0, os.path.join(ROOT, p))
    import numpy as np
    import common, vpt, vpt_scenes as S
    W, H = 256, 160
    inp = common.scene_inputs((2, 1, 2))
    t = common.setup(vpt.Vpt(W, H), inp, spp=1, total=3, diffuse=1)
    os.environ["VPT_DN_GATHER"] = "1"
    g = common.setup(vpt.Vpt(W, H), inp, spp=1, total=3, diffuse=1)
    p = S.default_denoising_params()
    cam = common.scene_camera(W, H)
    try:
        for f in range(4):
            for c in (t, g):
                c.render(cam, cam, f); c.denoise(p, cam, cam, f, f + 1); c.sync()
            for name in ("IlluminationOutput", "PrevIllumination"):
                a, b = t.read(name), g.read(name)
                print("mask", mask, "frame", f, name, "equal" if np.array_equal(a, b) else "max abs diff %g, mre %g" % (np.abs(a - b).max(), common.rel_err_stats(a, b)[0]))
        print("mask", mask, "OK, tma timeouts:", vpt.lib().vpt_debug_tma_timeouts())
    except Exception as ex:
        print("mask", mask, "FAILED:", ex)
        try:
            print("tma timeouts:", vpt.lib().vpt_debug_tma_timeouts())
        except Exception as ex2:
            print("counter unreadable:", ex2)
else:
    for m in ("4", "1", "2", "7"):
        subprocess.run([sys.executable, os.path.abspath(__file__), m])
