#!/usr/bin/env python3
"""Quick A/B of the large-world (cfg5-shaped) trace: 1024x256x1024 voxels (masks walked through L1/L2), 1920x1080, 8 spp, bounce
limits 8/2, cfg5's camera; prints ms per frame, Grays/s and DDA steps per ray. usage: [VPT_LIB=...] tools/cfg5_quick.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-path-tracing-voxel-blocks_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common, vpt
W, H, SPP = 1920, 1080, 8
inp = common.scene_inputs((32, 8, 32))
g = common.setup(vpt.Vpt(W, H), inp, spp=SPP, total=8, diffuse=2)
cam = vpt.camera_from_scene(W, H, [512.0, 200.0, 512.0], [-0.321564, -0.35, -0.946799], 90.0)
g.set_profiling(True)
g.render(cam, cam, 0)
rays, steps = g.counters()
g.set_profiling(False)
for f in range(1, 3):
    g.render(cam, cam, f)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stream = torch.cuda.ExternalStream(g.stream())
n = 5
e0.record(stream)
for f in range(3, 3 + n):
    g.render(cam, cam, f)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"lib": os.environ.get("VPT_LIB", "libvpt.so"), "ms_per_frame": round(ms, 3), "rays_per_frame": rays, "steps_per_ray": round(steps / rays, 1),
                  "grays_per_s": round(rays / (ms * 1e-3) / 1e9, 3)}))
