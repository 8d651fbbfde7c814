"""Local emissive lights: the exposed faces of emissive voxels as triangle lights in the reference's own
LightInfo packing (renderer/shaders/Light.h), 8 alias-sampled RIS candidates (closesthit.cu:330-375), BSDF-ray hits on emissive
faces (:854-901), finite-tmax visibility rays (:616-626, 736-755, 801-820) and the light-id remap after an edit (Restir.h:48-79).
CPU part: the oracle's packing helpers against numpy / closed forms. GPU part: CUDA path vs oracle."""
import numpy as np
import pytest

import common
import vpt_scenes as S

LANTERN = 13  # block id mapped to the emissive material below


def lantern_inputs(chunks=(2, 1, 2)):
    inp = common.scene_inputs(chunks)
    mats, b2m = S.default_materials()
    mats = np.concatenate([mats, np.zeros(1, S.MATERIAL_DTYPE)])
    mats[12]["albedo"] = (12.0, 9.0, 5.0)
    mats[12]["isEmissive"] = 1
    mats[12]["materialId"] = 12
    mats[12]["roughness"] = 1.0
    mats[12]["uvScale"] = 1.0
    b2m = b2m.copy()
    b2m[LANTERN] = 12
    inp["materials"], inp["b2m"] = mats, b2m
    return inp


def place_lanterns(ctxs, hits, w, h, where):
    """One lantern two voxels above the surface seen at each (fx, fy) fraction of the image."""
    out = []
    for fx, fy in where:
        hx, hy, hz, f = hits[int(fy * h), int(fx * w)]
        if f < 0:
            continue
        for c in ctxs:
            c.set_voxel(int(hx), int(hy) + 2, int(hz), LANTERN)
        out.append((int(hx), int(hy) + 2, int(hz)))
    return out


# ------------------------------------------------------------------------------------------------ CPU: oracle helpers
def test_fp16_packing_matches_numpy(oracle_lib):
    O = oracle_lib
    rng = np.random.default_rng(7)
    vals = np.concatenate([rng.standard_normal(4000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 4000).astype(np.float32),
                           np.array([0, -0.0, 1, 65504, 65519.99, 65520, 1e9, 2 ** -24, 2 ** -25, 2 ** -25 * 1.0001, 6.1e-5, 6.0975e-5, np.inf], np.float32)])
    with np.errstate(over="ignore"):
        ref = vals.astype(np.float16)
    for v, r in zip(vals, ref):
        assert O.f32_to_f16_bits(v) == int(r.view(np.uint16)), v
        assert O.f16_bits_to_f32(int(r.view(np.uint16))) == float(r.astype(np.float32)), r


def test_oracle_light_list_of_a_single_lantern(oracle_lib):
    """A lone emissive voxel in the air: 6 exposed faces -> 12 triangles of area 1/2, outward normals, alias pmf 1/12; resting on
    the ground it loses the covered bottom face."""
    O = oracle_lib
    inp = lantern_inputs()
    o = common.setup(O.Oracle(64, 48), inp)
    assert len(o.lights()[0]) == 0
    o.set_voxel(20, 30, 21, LANTERN)
    li, al, keys = o.lights()
    assert len(li) == 12 and len(keys) == 6
    lin = 20 + 64 * (21 + 64 * 30)
    assert list(keys) == [(lin << 3) | f for f in range(6)]
    assert np.allclose(al["p"], 1.0 / 12.0) and np.all(al["q"] == 1.0)
    # fp16 edge lengths of exactly 1, fp16 radiance
    assert np.all(li["scalars"] == (0x3C00 | (0x3C00 << 16)))
    assert np.all(li["radiance"][:, 0] == (int(np.float16(12.0).view(np.uint16)) | (int(np.float16(9.0).view(np.uint16)) << 16)))
    # centroids of the two triangles of the +y face (face 0): A = (20,31,21), u = +z, v = +x
    assert np.allclose(li["center"][0], (20 + 1 / 3, 31.0, 21 + 1 / 3), atol=1e-6)
    assert np.allclose(li["center"][1], (21 - 1 / 3, 31.0, 22 - 1 / 3), atol=1e-6)
    # a lantern standing on a solid voxel has 5 exposed faces
    o2 = common.setup(O.Oracle(64, 48), inp)
    g = o2.get_grid()
    ys = [y for y in range(32) if o2_voxel(g, 20, y, 21) != 0]
    top = max(ys)
    o2.set_voxel(20, top + 1, 21, LANTERN)
    g = o2.get_grid()
    exposed = sum(1 for dx, dy, dz in ((0, 1, 0), (0, -1, 0), (-1, 0, 0), (1, 0, 0), (0, 0, 1), (0, 0, -1)) if o2_voxel(g, 20 + dx, top + 1 + dy, 21 + dz) == 0)
    assert 1 <= exposed <= 5
    assert len(o2.lights()[0]) == 2 * exposed


def o2_voxel(grid_bytes, x, y, z, cx=2, cz=2):
    chunk = (x >> 5) + cx * ((z >> 5) + cz * (y >> 5))
    return grid_bytes.reshape(-1)[chunk * 32768 + (x & 31) + 32 * ((z & 31) + 32 * (y & 31))]


def test_oracle_lanterns_light_the_scene(oracle_lib):
    O = oracle_lib
    W, H = 128, 80
    inp = lantern_inputs()
    o = common.setup(O.Oracle(W, H), inp, spp=1, total=3, diffuse=1)
    cam = common.scene_camera(W, H)
    o.render(cam, cam, 0)
    base = o.read("Illumination")[..., :3].mean()
    assert place_lanterns([o], o.read("PrimaryHits"), W, H, [(0.5, 0.5)])
    o.render(cam, cam, 1)
    li = o.lights()[0]
    r = o.read_reservoirs(1)
    local = (r["lightData"] != 0) & ((r["lightData"] & 0x7FFFFFFF) < len(li))
    assert local.sum() > 50                       # pixels whose reservoir holds a local light
    assert o.read("Illumination")[..., :3].mean() > base * 1.01


# ------------------------------------------------------------------------------------------------ GPU: parity
@pytest.mark.gpu
def test_light_list_is_bit_identical_to_the_oracle(oracle_lib):
    import vpt
    O = oracle_lib
    inp = lantern_inputs((4, 1, 4))
    g = common.setup(vpt.Vpt(64, 48), inp)
    o = common.setup(O.Oracle(64, 48), inp)
    rng = np.random.default_rng(3)
    for _ in range(40):   # lanterns in the air, on the ground, buried, touching each other, on the grid boundary
        x, y, z = int(rng.integers(0, 128)), int(rng.integers(0, 32)), int(rng.integers(0, 128))
        for c in (g, o):
            c.set_voxel(x, y, z, LANTERN)
            c.set_voxel(min(x + 1, 127), y, z, LANTERN)
    for c in (g, o):
        c.set_voxel(0, 31, 0, LANTERN); c.set_voxel(127, 0, 127, LANTERN)
    lg, ag, kg = g.lights()
    lo, ao, ko = o.lights()
    assert len(lg) == len(lo) > 100
    assert np.array_equal(kg, ko)
    assert lg.tobytes() == lo.tobytes()
    assert ag.tobytes() == ao.tobytes()


@pytest.mark.gpu
def test_lantern_scene_matches_oracle_with_edits(oracle_lib):
    """4 spp, limits 3/1, ReSTIR on, moving camera; lanterns appear after frame 0, one is removed and one added after frame 2 (the
    light list changes between frames: the previous reservoirs' light ids go through the remap)."""
    import vpt
    O = oracle_lib
    W, H = 320, 192
    inp = lantern_inputs()
    g = common.setup(vpt.Vpt(W, H), inp, spp=4, total=3, diffuse=1)
    o = common.setup(O.Oracle(W, H), inp, spp=4, total=3, diffuse=1)
    p = S.default_denoising_params()
    cam = common.scene_camera(W, H)
    prev = cam
    lanterns = []
    for f in range(5):
        g.render(cam, prev, f)
        o.render(cam, prev, f)
        hg, ho = g.read("PrimaryHits"), o.read("PrimaryHits")
        assert np.array_equal(hg, ho), (f, int((hg != ho).any(-1).sum()))
        for name in ("Depth", "Material", "NormalRoughness", "Albedo"):
            assert np.array_equal(g.read(name), o.read(name)), (f, name)
        mre, tail, _ = common.rel_err_stats(g.read("Illumination")[..., :3], o.read("Illumination")[..., :3])
        assert mre <= 1e-3 and tail <= 2e-2, (f, mre, tail)
        rg, ro = g.read_reservoirs(f & 1), o.read_reservoirs(f & 1)
        same = (rg["lightData"] == ro["lightData"]).mean()
        assert same > 0.99, (f, same)
        if f >= 1:
            nl = len(o.lights()[0])
            local = (ro["lightData"] != 0) & ((ro["lightData"] & 0x7FFFFFFF) < nl)
            assert local.sum() > 200, (f, int(local.sum()))      # the lanterns are actually sampled
            assert ((rg["lightData"] == ro["lightData"]) & local).sum() > 0.97 * local.sum()
        assert g.counters()[0] <= o.counters()[0] + 50
        g.denoise(p, cam, prev, f, f + 1)
        o.denoise(p, cam, prev, f, f + 1)
        assert np.array_equal(g.read("HistoryLength"), o.read("HistoryLength")), f
        if f == 0:
            lanterns = place_lanterns([g, o], ho, W, H, [(0.5, 0.55), (0.3, 0.7), (0.75, 0.6)])
            assert len(lanterns) >= 2
            assert g.lights()[0].tobytes() == o.lights()[0].tobytes()
        if f == 2:
            x, y, z = lanterns[0]
            for c in (g, o):
                c.set_voxel(x, y, z, 0)
                c.set_voxel(x + 1, y + 1, z, LANTERN)
        prev = cam
        cam = vpt.camera_set_yaw_pitch(cam, cam[15] + np.float32(0.4 * np.pi / 180.0), cam[16])


@pytest.mark.gpu
@pytest.mark.parametrize("lanterns", [False, True])
def test_throughput_path_is_bit_identical_to_the_instrumented_path(lanterns):
    """vpt_set_profiling(0) — what bench.py times — selects the kernel instances without the step counter and the per-launch
    events. Same engine, same results: every plane, the reservoirs and the ray count are bit-identical to the instrumented
    (default) path the other parity tests run, with and without local lights (finite-tmax visibility rays, closest-hit BSDF rays)."""
    import vpt
    W, H = 320, 192
    inp = lantern_inputs() if lanterns else common.scene_inputs((2, 1, 2))
    a = common.setup(vpt.Vpt(W, H), inp, spp=3, total=3, diffuse=1)
    b = common.setup(vpt.Vpt(W, H), inp, spp=3, total=3, diffuse=1)
    b.set_profiling(False)
    p = S.default_denoising_params()
    cam = common.scene_camera(W, H)
    prev = cam
    for f in range(3):
        for c in (a, b):
            c.render(cam, prev, f)
        for name in ("PrimaryHits", "Depth", "Material", "NormalRoughness", "GeoNormalThinfilm", "MaterialParameter", "Albedo", "Illumination"):
            assert a.read(name).tobytes() == b.read(name).tobytes(), (f, name)
        assert a.read_reservoirs(f & 1).tobytes() == b.read_reservoirs(f & 1).tobytes(), f
        assert a.counters()[0] == b.counters()[0], f
        assert a.counters()[1] > 10 * a.counters()[0] and b.counters()[1] == 0   # steps are only counted on the instrumented path
        for c in (a, b):
            c.denoise(p, cam, prev, f, f + 1)
        for name in ("HistoryLength", "IlluminationOutput"):
            assert a.read(name).tobytes() == b.read(name).tobytes(), (f, name)
        if f == 0 and lanterns:
            assert len(place_lanterns([a, b], a.read("PrimaryHits"), W, H, [(0.5, 0.55), (0.3, 0.7), (0.75, 0.6)])) >= 2
        prev = cam
        cam = vpt.camera_set_yaw_pitch(cam, cam[15] + np.float32(0.4 * np.pi / 180.0), cam[16])
