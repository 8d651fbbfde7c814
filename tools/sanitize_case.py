#!/usr/bin/env python3
"""Small end-to-end case for compute-sanitizer memcheck: every kernel of the product path at a tiny size (trace with ReSTIR,
continuation depth rounds, multi-wave, edits, generated sky, denoiser chain incl. firefly / history-fix lists, tone map, textured materials, HitDistReconstruction / PrePass, the block picker)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", os.path.join("real-time-path-tracing-voxel-blocks_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import common, vpt, vpt_scenes as S
W, H = 101, 67   # not multiples of the 8x4 tiles / 32x8 blocks
inp = common.scene_inputs((2, 1, 2))
mats = inp["materials"].copy(); mats[1]["roughness"] = 0.0; mats[1]["metallic"] = 1
g = common.setup(vpt.Vpt(W, H), dict(inp, materials=mats), spp=3, total=4, diffuse=2)
g.set_wave_budget(2 * 13 * 17 * 32)
g.generate_sky(S.DEFAULT_SKY_PARAMS, S.load_sky_tables())
p = S.default_denoising_params(yaml_overrides=False)   # 5 a-trous iterations: 12 spatial passes, steps up to 2048
cam = common.scene_camera(W, H)
prev = cam
for f in range(6):
    if f == 3: g.set_voxel(30, 30, 38, 3)
    g.render(cam, prev, f)
    g.denoise(p, cam, prev, f, f + 1)
    prev = cam
    cam = vpt.camera_set_yaw_pitch(cam, cam[15] + np.float32(0.01), cam[16])
# textured materials + the default-off denoiser passes + the picker
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_textures as T
g.set_textures(*T.textured_scene(vpt, len(mats), size=32))
p2 = p.copy(); p2["enableHitDistanceReconstruction"] = 1; p2["enablePrePass"] = 1; p2["atrousIterationNum"] = 1
for f in range(6, 9):
    pk = g.pick_voxel(cam[6:9], cam[9:12])
    if pk["hitSurface"]: g.set_voxel(*pk["deletePos"], 0)
    g.render(cam, prev, f)
    g.denoise(p2, cam, prev, f, f + 1)
    prev = cam
rgb8, ldr = g.tonemap(vpt.default_tonemapping_params())
out = g.read("IlluminationOutput")
assert np.isfinite(out).all() and rgb8.shape == (H, W, 3)
print("sanitize case ok", float(out.mean()), g.counters())
