"""SURVEY 8a rows D2 (HitDistReconstruction, HitDistReconstruction.h:50-161) and D3 (PrePass, PrePass.h:6-149): both are off in
the shipped global_settings.yaml and on the path when enabled. CPU tests: identities of the oracle restatement; gpu tests:
the CUDA passes against the oracle, alone and inside the full chain."""
import numpy as np
import pytest

import common
import vpt_scenes as S


def _params(**kw):
    p = S.default_denoising_params()
    for k, v in kw.items():
        p[k] = v
    return p


def _only(**kw):
    base = dict(enableHitDistanceReconstruction=0, enablePrePass=0, enableTemporalAccumulation=0, enableHistoryFix=0,
                enableHistoryClamping=0, enableSpatialFiltering=0, enableFireflyFilter=0)
    base.update(kw)
    return _params(**base)


def _flat_scene(ctx, W, H, illum, depth=None, frame=0):
    """A fronto-parallel plane: constant normal, smoothly varying depth, one material."""
    gb = S.synthetic_gbuffer(W, H, frame)
    gb["NormalRoughness"][...] = (0.0, 0.0, -1.0, 1.0)
    gb["Depth"][...] = 10.0 if depth is None else depth
    gb["Material"][...] = 3.0
    gb["Albedo"][...] = (1.0, 1.0, 1.0, 1.0)
    gb["Illumination"] = illum.astype(np.float32)
    ctx.begin_external_frame()
    for name in ("Illumination", "Depth", "NormalRoughness", "Material", "Albedo"):
        ctx.write(name, gb[name])
    return gb


def test_oracle_hitdist_reconstruction_identities(oracle_lib):
    import vpt
    O = oracle_lib
    W, H = 48, 32
    o = O.Oracle(W, H)
    cam = vpt.camera_set_yaw_pitch(vpt.camera_init(W, H), 0.0, 0.0)
    rng = np.random.default_rng(1)
    illum = np.zeros((H, W, 4), np.float32)
    illum[..., :3] = rng.random((H, W, 3))
    illum[..., 3] = 7.5
    holes = rng.random((H, W)) < 0.3
    illum[holes, 3] = 0.0                           # missing hit distances
    _flat_scene(o, W, H, illum)
    o.denoise(_only(enableHitDistanceReconstruction=1), cam, cam, 0, 1)
    ping = o.read("IlluminationPing")
    assert np.array_equal(ping[..., :3], illum[..., :3])                     # radiance passes through
    assert np.allclose(ping[..., 3], 7.5, rtol=1e-5)                         # every hole is filled from its neighbours
    # a pixel whose whole 5x5 neighbourhood has no hit distance stays 0
    illum2 = illum.copy()
    illum2[..., 3] = 0.0
    illum2[10, 10, 3] = 4.0
    _flat_scene(o, W, H, illum2)
    o.denoise(_only(enableHitDistanceReconstruction=1), cam, cam, 0, 1)
    ping = o.read("IlluminationPing")[..., 3]
    assert ping[10, 10] == 4.0 and ping[20, 30] == 0.0
    assert np.all(ping[8:13, 8:13] == 4.0) and ping[10, 13] == 0.0           # 5x5 footprint, nothing beyond
    # the composite of this configuration reads Ping (finalResultBuffer = 1, Denoiser.cu:96)
    assert np.allclose(o.read("IlluminationOutput")[..., :3], illum2[..., :3])


def test_oracle_prepass_identities(oracle_lib):
    import vpt
    O = oracle_lib
    W, H = 64, 48
    o = O.Oracle(W, H)
    cam = vpt.camera_set_yaw_pitch(vpt.camera_init(W, H), 0.0, 0.0)
    const = np.zeros((H, W, 4), np.float32)
    const[...] = (0.25, 0.5, 0.75, 3.0)
    # the reconstruction fills Ping with the (constant) input; the pre-blur of a constant is the constant
    _flat_scene(o, W, H, const)
    o.denoise(_only(enableHitDistanceReconstruction=1, enablePrePass=1), cam, cam, 0, 5)
    out = o.read("Illumination")
    assert np.allclose(out, const, rtol=2e-6)
    # noise is smoothed (variance drops), mean is preserved on a flat receiver
    rng = np.random.default_rng(2)
    noisy = const.copy()
    noisy[..., :3] += (rng.random((H, W, 3)).astype(np.float32) - 0.5) * 0.2
    noisy[..., 3] = 40.0                            # long hit distance -> full 30-pixel radius scaled by hit distance factor
    # a true plane 10 units in front of the camera, facing it: depth = 10 / cos(angle to the view axis)
    dirs = np.array([[O.uv_to_world_direction(cam, (x + 0.5) / W, (y + 0.5) / H) for x in range(W)] for y in range(H)], np.float32)
    axis = O.uv_to_world_direction(cam, 0.5, 0.5)
    plane_depth = (10.0 / (dirs @ axis)).astype(np.float32)
    gb = _flat_scene(o, W, H, noisy, depth=plane_depth)
    o.write("NormalRoughness", np.concatenate([np.broadcast_to(-axis, (H, W, 3)), np.ones((H, W, 1), np.float32)], -1).astype(np.float32))
    o.denoise(_only(enableHitDistanceReconstruction=1, enablePrePass=1), cam, cam, 0, 5)
    out = o.read("Illumination")
    inner = (slice(8, H - 8), slice(8, W - 8))
    v_out, v_in = out[inner][..., 0].var(), noisy[inner][..., 0].var()
    assert v_out < 0.5 * v_in, (v_out, v_in)
    assert abs(out[inner][..., 0].mean() - noisy[inner][..., 0].mean()) < 0.01
    # frame index rotates the Poisson disc: two frame indices give different results
    _flat_scene(o, W, H, noisy, depth=plane_depth)
    o.denoise(_only(enableHitDistanceReconstruction=1, enablePrePass=1), cam, cam, 0, 6)
    assert not np.array_equal(out, o.read("Illumination"))


def test_weyl_rotator_matches_closed_form(oracle_lib):
    """Weyl1D(0.5, n) = fract(0.5 + n*10368889 / 2^24) with the int multiply wrapping (DenoiserCommon.h:336-339)."""
    O = oracle_lib
    for n in (0, 1, 7, 207, 208, 5000):
        m = np.array([n], np.uint32) * np.uint32(10368889)
        m = m.view(np.int32)[0]
        x = np.float32(0.5) + np.float32(m) / np.float32(16777216.0)
        frac = np.float32(x - np.trunc(x))
        angle = np.float32(frac * np.float32(90.0 * np.pi / 180.0))
        rot = O.prepass_rotator(n)
        assert np.allclose(rot, [np.cos(angle), np.sin(angle), -np.sin(angle), np.cos(angle)], atol=2e-6), n


@pytest.mark.gpu
def test_hitdist_and_prepass_alone_match_oracle(oracle_lib):
    import vpt
    O = oracle_lib
    W, H = 320, 200
    g, o = vpt.Vpt(W, H), O.Oracle(W, H)
    cam = vpt.camera_init(W, H)
    cam[6:9] = (0.0, 6.0, 0.0)
    cam = vpt.camera_set_yaw_pitch(cam, 0.0, 0.0)
    rng = np.random.default_rng(5)
    for f, flags in enumerate((dict(enableHitDistanceReconstruction=1), dict(enableHitDistanceReconstruction=1, enablePrePass=1))):
        gb = S.synthetic_gbuffer(W, H, f)
        hd = (rng.random((H, W)).astype(np.float32) * 20.0 + 0.5)
        hd[rng.random((H, W)) < 0.25] = 0.0
        gb["Illumination"][..., 3] = hd
        for ctx in (g, o):
            ctx.begin_external_frame()
            for name in ("Illumination", "Depth", "NormalRoughness", "Material", "Albedo"):
                ctx.write(name, gb[name])
            ctx.denoise(_only(**flags), cam, cam, 0, 3 + f)
        a, b = g.read("IlluminationPing"), o.read("IlluminationPing")
        m, outl, dmax = common.rel_err_stats(a, b)
        assert m <= 1e-5 and outl <= 1e-3, ("ping", flags, m, outl, dmax)
        if "enablePrePass" in flags:
            a, b = g.read("Illumination"), o.read("Illumination")
            m, outl, dmax = common.rel_err_stats(a, b)
            assert m <= 2e-5 and outl <= 2e-3, ("prepass", m, outl, dmax)
            assert not np.array_equal(b, gb["Illumination"])
        m, outl, dmax = common.rel_err_stats(g.read("IlluminationOutput"), o.read("IlluminationOutput"))
        assert m <= 2e-5 and outl <= 2e-3, ("output", flags, m, outl, dmax)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [dict(enableHitDistanceReconstruction=1, enablePrePass=1), dict(enablePrePass=1)],
                         ids=["hitdist+prepass", "prepass-on-stale-ping"])
def test_full_chain_with_prepasses_matches_oracle(oracle_lib, flags):
    """Rendered frames + the whole chain with the optional passes on. With the reconstruction off the pre-pass blurs whatever
    the previous frame left in IlluminationPing (the reference does exactly that): Ping must therefore evolve identically."""
    import vpt
    O = oracle_lib
    W, H = 256, 160
    inp = common.scene_inputs((2, 1, 2))
    g = common.setup(vpt.Vpt(W, H), inp, spp=1, total=3, diffuse=1)
    o = common.setup(O.Oracle(W, H), inp, spp=1, total=3, diffuse=1)
    p = _params(**flags)
    cam = common.scene_camera(W, H)
    prev = cam
    for f in range(4):
        g.render(cam, prev, f)
        o.render(cam, prev, f)
        g.write("Illumination", o.read("Illumination"))
        g.write_reservoirs(f & 1, o.read_reservoirs(f & 1))
        g.denoise(p, cam, prev, f, f + 1)
        o.denoise(p, cam, prev, f, f + 1)
        assert np.array_equal(g.read("HistoryLength"), o.read("HistoryLength")), f
        for name, tol, tail in (("Illumination", 5e-5, 5e-3), ("IlluminationOutput", 1e-4, 1e-2), ("IlluminationPing", 2e-4, 1e-2)):
            m, outl, dmax = common.rel_err_stats(g.read(name), o.read(name))
            assert m <= tol and outl <= tail, (f, name, m, outl, dmax)
        prev = cam
        if f >= 1:
            cam = vpt.camera_set_yaw_pitch(cam, cam[15] + np.float32(0.5 * np.pi / 180.0), cam[16])
