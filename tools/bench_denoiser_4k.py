#!/usr/bin/env python3
"""Config 3 on one GPU: the denoiser chain alone at 3840x2160 on the synthetic G-buffer, per-pass CUDA-event times
and achieved algorithmic GB/s against the measured HBM peak. usage: python tools/bench_denoiser_4k.py [frames] > out.json"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", os.path.join("real-time-path-tracing-voxel-blocks_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
sys.path.insert(0, ROOT)
import vpt, vpt_scenes as S
import bench
W, H = 3840, 2160
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
g = vpt.Vpt(W, H)
p = S.default_denoising_params()
cam = vpt.camera_init(W, H); cam[6:9] = (0.0, 6.0, 0.0); cam = vpt.camera_set_yaw_pitch(cam, 0.0, 0.0)
g.set_profiling(True)
base = S.synthetic_gbuffer(W, H, 0)
best = None
for f in range(frames):
    # the G-buffer is static; only the noisy radiance changes per frame (hash of the frame index)
    noisy = S.synthetic_gbuffer(W, H, f)["Illumination"] if f < 3 else base["Illumination"] * np.float32(1.0 + 0.01 * (f % 5))
    g.begin_external_frame()
    for name in ("Depth", "NormalRoughness", "Material", "Albedo"):
        g.write(name, base[name])
    g.write("Illumination", noisy)
    g.denoise(p, cam, cam, f, f + 1)
    g.sync()
    t = g.timings()
    if f >= 3 and (best is None or t["denoise_total_ms"] < best["denoise_total_ms"]):
        best = t
peak, src = bench.load_peaks()
npix = W * H
rows = []
for name, key, mult in (("prep_firefly_sky", "firefly", 1), ("temporal", "temporal", 1), ("history_clamp", "history_clamp", 1),
                        ("atrous_smem", "atrous_smem", 1), ("atrous", "atrous", int(best["atrous_passes"]))):
    ms = float(best[{"firefly": "firefly_ms", "temporal": "temporal_ms", "history_clamp": "history_clamp_ms", "atrous_smem": "atrous_smem_ms",
                     "atrous": "atrous_ms"}[key]])
    b = bench.PASS_BYTES[key][0] * mult * npix + (16 * npix if key == "atrous" else 0)
    rows.append({"name": name, "ms": round(ms, 4), "algorithmic_bytes": b, "gbs": round(b / (ms * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(b / (ms * 1e-3) / 1e9 / peak, 3)})
chain_b = sum(r["algorithmic_bytes"] for r in rows)
print(json.dumps({"workload": "cfg3: denoiser chain alone, 3840x2160 synthetic G-buffer, shipped settings (4 spatial passes), 1 GPU",
                  "denoise_total_ms": round(float(best["denoise_total_ms"]), 4), "chain_algorithmic_bytes": chain_b,
                  "chain_gbs": round(chain_b / (float(best["denoise_total_ms"]) * 1e-3) / 1e9, 1), "chain_frac_of_hbm_peak": round(chain_b / (float(best["denoise_total_ms"]) * 1e-3) / 1e9 / peak, 3),
                  "hbm_peak_gbs": peak, "peak_source": src, "passes": rows}))
