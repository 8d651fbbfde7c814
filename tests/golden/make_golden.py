#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/. Run in the BUILD container, where /root/reference exists:

    python tests/golden/make_golden.py

  perlin_ref.npz      outputs of the REFERENCE's own PerlinNoiseGenerator (voxelengine/Noise.cpp + ext/PerlinNoise.hpp
                      compiled as they lie into oracle/_ref/libref_noise.so): 4096 random points + the 2x1x2 chunk maps.
  imagediff_ref.json  outputs of the REFERENCE's own ImageDiff.cpp (oracle/_ref/libref_imagediff.so) on seeded image pairs.
  camera_kat.json     the known answers of the reference's camera unit test (renderer/test/camera/test.cpp:137-259).
  oracle_cfg1.json    digest of the oracle's config-1 render (256x256, 1 spp, 1 bounce): drift detector for the oracle itself.
The fixtures travel to the GPU box; /root/reference does not."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "real-time-path-tracing-voxel-blocks_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402

O.build()
ref_noise = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_noise.so"))
ref_noise.ref_perlin_noise_map.argtypes = [C.c_int, C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
rng = np.random.default_rng(20261018)
xs = rng.uniform(-4.0, 6.0, 4096).astype(np.float32)
ys = rng.uniform(-4.0, 6.0, 4096).astype(np.float32)
out = np.zeros(4096, np.float32)
ref_noise.ref_perlin_noise_map(4, 124, 4096, xs.ctypes.data, ys.ctypes.data, out.ctypes.data)
# chunk maps exactly as initVoxelsMultiChunk samples them (VoxelSceneGen.cu:361-376), 2x1x2 chunks
maps = np.zeros((4, 32, 32), np.float32)
for c in range(4):
    cx, cz = c % 2, c // 2
    gx = (cx * 32 + np.arange(32)).astype(np.float32)
    gz = (cz * 32 + np.arange(32)).astype(np.float32)
    freq = np.float32(1.0) / np.float32(64)
    X, Z = np.meshgrid(gx * freq, gz * freq)  # [z][x]
    fx, fz = np.ascontiguousarray(X.ravel(), np.float32), np.ascontiguousarray(Z.ravel(), np.float32)
    o = np.zeros(1024, np.float32)
    ref_noise.ref_perlin_noise_map(4, 124, 1024, fx.ctypes.data, fz.ctypes.data, o.ctypes.data)
    maps[c] = o.reshape(32, 32)
np.savez_compressed(os.path.join(HERE, "perlin_ref.npz"), xs=xs, ys=ys, values=out, chunk_maps_2x1x2=maps)

ref_id = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_imagediff.so"))
ref_id.ref_imagediff.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
cases = []
for i, (w, h, noise) in enumerate([(64, 48, 0), (64, 48, 2), (96, 64, 9), (80, 80, 40), (33, 17, 120)]):
    r = np.random.default_rng(100 + i)
    a = r.integers(0, 256, (h, w, 3), dtype=np.uint8)
    a = (a.astype(np.float32) * 0.25 + np.linspace(0, 180, w, dtype=np.float32)[None, :, None]).astype(np.uint8)
    b = np.clip(a.astype(np.int32) + r.integers(-noise, noise + 1, a.shape), 0, 255).astype(np.uint8)
    res, flags = np.zeros(3, np.float32), np.zeros(4, np.int32)
    ref_id.ref_imagediff(a.ctypes.data, b.ctypes.data, w, h, 3, res.ctypes.data, flags.ctypes.data)
    cases.append(dict(seed=100 + i, w=w, h=h, noise=noise, rmse=float(res[0]), ssim=float(res[1]), ratio=float(res[2]),
                      differentPixels=int(flags[0]), isIdentical=bool(flags[1]), isVeryClose=bool(flags[2]), isClose=bool(flags[3])))
json.dump(cases, open(os.path.join(HERE, "imagediff_ref.json"), "w"), indent=1)

kat = dict(source="renderer/test/camera/test.cpp:137-259", width=800, height=600, yaw=0.0, pitch=0.0, tol=1e-3,
           tanHalfFovX=1.0, tanHalfFovY=float(np.tan(np.float32(90.0 * (600.0 / 800.0) * np.pi / 180.0) * 0.5)),
           uv_to_dir=[dict(uv=[0.5, 0.5], dir=[0.0, 0.0, 1.0]),
                      dict(uv=[0.0, 0.0], dir=(np.array([1, -0.66818, 1]) / np.linalg.norm([1, -0.66818, 1])).tolist()),
                      dict(uv=[1.0, 1.0], dir=(np.array([-1, 0.66818, 1]) / np.linalg.norm([1, 0.66818, 1])).tolist()),
                      dict(uv=[0.0, 1.0], dir=(np.array([1, 0.66818, 1]) / np.linalg.norm([1, 0.66818, 1])).tolist()),
                      dict(uv=[1.0, 0.0], dir=(np.array([-1, -0.66818, 1]) / np.linalg.norm([1, 0.66818, 1])).tolist())],
           round_trip_uvs=[[0.0, 0.0], [0.5, 0.5], [1.0, 1.0], [0.0, 1.0], [1.0, 0.0], [0.25, 0.75], [0.3, 0.7]])
json.dump(kat, open(os.path.join(HERE, "camera_kat.json"), "w"), indent=1)

import common  # noqa: E402
import vpt_scenes as S  # noqa: E402
inp = common.scene_inputs((2, 1, 2), noise_fn=O.perlin_noise_chunks, alias_fn=O.build_alias_table)
o = common.setup(O.Oracle(256, 256), inp, spp=1, total=1, diffuse=1)
cam = O.camera_from_scene(256, 256, S.SCENE_CAMERA["position"], S.SCENE_CAMERA["direction"], S.SCENE_CAMERA["fov"])
o.render(cam, cam, 0)
hits = o.read("PrimaryHits")
ill = o.read("Illumination")
grid = o.get_grid()
dig = dict(config="cfg1: 2x1x2 chunks, 256x256, 1 spp, 1 bounce, iterationIndex 0, scene_export.yaml camera",
           grid_sha256=hashlib.sha256(grid.tobytes()).hexdigest(), solid_voxels=int((grid != 0).sum()),
           primary_hits_sha256=hashlib.sha256(hits.tobytes()).hexdigest(), hit_fraction=float((hits[..., 3] >= 0).mean()),
           face_histogram=np.bincount(hits[..., 3][hits[..., 3] >= 0], minlength=7).tolist(),
           mean_radiance=[float(v) for v in ill[..., :3].reshape(-1, 3).mean(0)], rays=int(o.counters()[0]))
json.dump(dig, open(os.path.join(HERE, "oracle_cfg1.json"), "w"), indent=1)
print("golden fixtures written to", HERE)
