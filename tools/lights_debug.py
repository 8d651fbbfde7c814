import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", os.path.join("real-time-path-tracing-voxel-blocks_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, common, vpt, oracle as O, vpt_scenes as S
from test_lights import lantern_inputs, place_lanterns, LANTERN
O.build()
W, H = 320, 192
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1
restir = int(sys.argv[2]) if len(sys.argv) > 2 else 1
inp = lantern_inputs()
g = common.setup(vpt.Vpt(W, H), inp, spp=spp, total=3, diffuse=1, restir=restir)
o = common.setup(O.Oracle(W, H), inp, spp=spp, total=3, diffuse=1, restir=restir)
cam = common.scene_camera(W, H)
for f in range(3):
    g.render(cam, cam, f); o.render(cam, cam, f)
    a, b = g.read("Illumination")[..., :3], o.read("Illumination")[..., :3]
    mre, tail, dmax = common.rel_err_stats(a, b)
    rg, ro = g.read_reservoirs(f & 1), o.read_reservoirs(f & 1)
    nl = len(o.lights()[0])
    local = (ro["lightData"] != 0) & ((ro["lightData"] & 0x7FFFFFFF) < nl)
    same = rg["lightData"] == ro["lightData"]
    wrel = np.abs(rg["weightSum"] - ro["weightSum"]) / np.maximum(np.abs(ro["weightSum"]), 1e-6)
    print("frame", f, "spp", spp, "restir", restir, "mre %.2e tail %.2e dmax %.3g" % (mre, tail, dmax), "| reservoirs same %.4f" % same.mean(), "local", int(local.sum()),
          "same among local %.4f" % (same[local].mean() if local.any() else 1), "weight rel>1e-3 among same: %.4f" % (wrel[same] > 1e-3).mean(),
          "rays", g.counters()[0], o.counters()[0])
    rel = np.abs(a - b).max(-1) / np.maximum(np.maximum(np.abs(a), np.abs(b)).max(-1), 1e-3)
    bad = rel > 1e-2
    print("   pixels off >1e-2:", int(bad.sum()), "of which reservoir differs:", int((bad & ~same).sum()), "local-light pixels:", int((bad & local).sum()))
    if f == 0:
        place_lanterns([g, o], o.read("PrimaryHits"), W, H, [(0.5, 0.55), (0.3, 0.7), (0.75, 0.6)])
    if f >= 1 and len(sys.argv) > 3:
        tilesX = (W + 7) // 8
        nslots = tilesX * ((H + 3) // 4) * 32
        planes = []
        for which in range(6):
            buf = np.zeros((nslots, 4), np.uint32)
            assert vpt.lib().vpt_debug_read_wave(g.ctx, which, buf.ctypes.data_as(__import__("ctypes").c_void_p), __import__("ctypes").c_size_t(nslots)) == 0
            planes.append(buf)
        ys, xs = np.nonzero(local & ~same)
        for k in range(min(6, len(ys))):
            y, x = ys[k], xs[k]
            slot = ((y // 4) * tilesX + x // 8) * 32 + (y % 4) * 8 + (x % 8)
            print("   slot", slot, "candC.x", planes[0][slot, 0], "ris.x", hex(planes[1][slot, 0]), "rstA.x", hex(planes[2][slot, 0]), "rstB", [hex(v) for v in planes[3][slot, :2]],
                  "lightA", planes[4][slot].view(np.float32), "light2A", planes[5][slot].view(np.float32))
        for k in range(min(8, len(ys))):
            y, x = ys[k], xs[k]
            print("   px", x, y, "gpu", [hex(int(rg["lightData"][y, x])), hex(int(rg["uvData"][y, x])), float(rg["weightSum"][y, x]), float(rg["targetPdf"][y, x]), float(rg["M"][y, x])],
                  "orc", [hex(int(ro["lightData"][y, x])), hex(int(ro["uvData"][y, x])), float(ro["weightSum"][y, x]), float(ro["targetPdf"][y, x]), float(ro["M"][y, x])], "rad", a[y, x], b[y, x])
