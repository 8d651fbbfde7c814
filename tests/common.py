"""Shared scene setup for the parity tests: the same bytes go to the CUDA path (vpt.Vpt) and to the oracle."""
import numpy as np

import vpt_scenes as S


def scene_inputs(chunks=(2, 1, 2), constant_sky=None, noise_fn=None, alias_fn=None):
    """Host-side inputs. noise_fn/alias_fn default to the product's host helpers (libvpt host code, no GPU)."""
    import vpt
    noise_fn = noise_fn or vpt.perlin_noise_chunks
    alias_fn = alias_fn or vpt.build_alias_table
    cx, cy, cz = chunks
    sky, sun, sun_dir = S.synthetic_sky(constant=constant_sky)
    if constant_sky is not None:
        sun_w = np.ones(sun.shape[0] * sun.shape[1], np.float32)  # black sun: uniform alias table, zero radiance
    else:
        sun_w = S.sky_pdf_weights(sun)
    mats, b2m = S.default_materials()
    return dict(chunks=chunks, tables=S.load_tables(), noise=noise_fn(cx, cy, cz), materials=mats, b2m=b2m,
                sky=sky, sun=sun, sun_dir=sun_dir, sky_alias=alias_fn(S.sky_pdf_weights(sky)), sun_alias=alias_fn(sun_w))


def setup(ctx, inp, spp=1, total=3, diffuse=1, restir=1):
    cx, cy, cz = inp["chunks"]
    ctx.set_tables(inp["tables"])
    ctx.generate_terrain(cx, cy, cz, inp["noise"])
    ctx.set_materials(inp["materials"], inp["b2m"])
    ctx.set_sky(inp["sky"], inp["sun"], inp["sky_alias"], inp["sun_alias"], inp["sun_dir"])
    ctx.set_trace_params(spp, total, diffuse, restir)
    return ctx


def scene_camera(width, height, chunks=(2, 1, 2)):
    """data/scene/scene_export.yaml camera. The reference scene is 2x1x2 chunks; the 16-chunk (4x1x4) terrain is
    sampled at half the noise frequency and is 8 voxels higher under the camera, so the eye is lifted by 8."""
    import vpt
    pos = list(S.SCENE_CAMERA["position"])
    if chunks[0] >= 4:
        pos[1] += 8.0
    return vpt.camera_from_scene(width, height, pos, S.SCENE_CAMERA["direction"], S.SCENE_CAMERA["fov"])


def rel_err_stats(a, b, floor=1e-3):
    """mean relative error and outlier fraction of two float images."""
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    d = np.abs(a - b)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    rel = d / denom
    return float(rel.mean()), float((rel > 1e-3).mean()), float(d.max())
