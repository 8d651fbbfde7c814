"""Sky / sun generator (SURVEY 8f "next" row 1): the oracle's restatement of SkyModel::update (renderer/sky/Sky.cu) and the
CUDA generator behind vpt_generate_sky."""
import numpy as np
import pytest

import common
import vpt_scenes as S


@pytest.fixture(scope="module")
def sky(oracle_lib):
    tb = S.load_sky_tables()
    return oracle_lib.generate_sky(S.DEFAULT_SKY_PARAMS, tb), tb


def test_sun_direction_and_host_state_match_oracle(oracle_lib, sky):
    import vpt
    (_, _, _, _, sd), tb = sky
    # default SkyParams (timeOfDay 0.25, axis 45 deg): the sun the reference computes (Sky.cu:362-367)
    assert np.allclose(sd, S.reference_sun_dir(), atol=2e-7)
    for params in (S.DEFAULT_SKY_PARAMS, (0.4, 30.0, 90.0, 0.5), (0.12, 70.0, 200.0, 1.0)):
        c1, r1, s1 = oracle_lib.sky_state(params, tb)
        c2, r2, s2 = vpt.sky_state(params, tb)     # libvpt host code (no GPU)
        assert np.array_equal(c1, c2) and np.array_equal(r1, r2) and np.array_equal(s1, s2), params
        assert abs(float(np.linalg.norm(s1)) - 1.0) < 1e-6


def test_sky_map_properties(oracle_lib, sky):
    (skym, sun, sky_pdf, sun_pdf, sd), tb = sky
    H, W = skym.shape[:2]
    assert np.isfinite(skym).all() and (skym >= 0).all() and np.isfinite(sun).all() and (sun >= 0).all()
    lum = S.luminance(skym[..., :3])
    assert np.allclose(sky_pdf.reshape(H, W), lum, rtol=1e-6, atol=1e-9)
    assert np.allclose(sun_pdf.reshape(sun.shape[:2]), S.luminance(sun[..., :3]), rtol=1e-6)
    # the brightest sky texel of the upper hemisphere looks towards the sun (equal-area sphere map, LinearMath.h:1857-1863)
    y, x = np.unravel_index(np.argmax(lum[H // 2:]), (H // 2, W))
    v = (y + H // 2 + 0.5) / H; u = (x + 0.5) / W
    yy = 2 * v - 1; r = np.sqrt(1 - yy * yy)
    d = np.array([r * np.cos(2 * np.pi * u), yy, r * np.sin(2 * np.pi * u)])
    assert float(d @ sd) > 0.98   # within ~10 degrees (the circumsolar peak is pulled towards the horizon)
    # lower hemisphere (Sky.cu:285-307): row 0 is the mist colour = mean upper-hemisphere luminance / 2 (sum over half the map / all texels)
    mist = lum[H // 2:].astype(np.float64).sum() / (W * H)
    assert np.allclose(skym[0, :, :3], mist, rtol=1e-5)
    # the row just below the horizon: smoothstep blend of mist and the horizon row with w = clamp((v + 0.4) * 2), v = -0.5/256
    w = np.clip((-0.5 / (H // 2) + 0.4) * 2.0, 0, 1); sw = w * w * (3 - 2 * w)
    assert np.allclose(skym[H // 2 - 1, :, :3], mist + sw * (skym[H // 2, :, :3] - mist), rtol=1e-4, atol=1e-6)
    # solar disc: limb darkening, u = 0 is the disc centre (LinearMath.h:1871-1885)
    sl = S.luminance(sun[..., :3])
    assert sl[:, 0].mean() > sl[:, -1].mean() > 0
    # brightness is a pure scale
    sky2, sun2, _, _, _ = oracle_lib.generate_sky((0.25, 45.0, 0.0, 0.5), tb)
    assert np.allclose(sky2[H // 2:], 0.5 * skym[H // 2:], rtol=1e-6, atol=1e-9) and np.allclose(sun2, 0.5 * sun, rtol=1e-6)


@pytest.mark.gpu
def test_cuda_sky_matches_oracle_and_renders(oracle_lib, sky):
    """vpt_generate_sky (device Hosek-Wilkie evaluation + host alias tables) vs the oracle: maps to libm precision, then a
    render lit by the generated sky on both sides (primary hits exact, radiance mean relative error <= 1e-3)."""
    import vpt
    (skym, sun, sky_pdf, sun_pdf, sd), tb = sky
    W, H = 192, 128
    inp = common.scene_inputs((2, 1, 2))
    g = common.setup(vpt.Vpt(W, H), inp, spp=2, total=3, diffuse=1)
    g.generate_sky(S.DEFAULT_SKY_PARAMS, tb)
    gs, gu, gd = g.read_sky()
    assert np.array_equal(gd, sd)
    for a, b, name in ((gs, skym, "sky"), (gu, sun, "sun")):
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-6)
        assert rel[..., :3].max() < 2e-4 and rel[..., :3].mean() < 2e-6, (name, float(rel.max()), float(rel.mean()))
    o = common.setup(oracle_lib.Oracle(W, H), inp, spp=2, total=3, diffuse=1)
    o.set_sky(skym, sun, oracle_lib.build_alias_table(sky_pdf), oracle_lib.build_alias_table(sun_pdf), sd)
    cam = common.scene_camera(W, H)
    for f in range(2):
        g.render(cam, cam, f)
        o.render(cam, cam, f)
        assert np.array_equal(g.read("PrimaryHits"), o.read("PrimaryHits"))
        mean_rel, outliers, _ = common.rel_err_stats(g.read("Illumination")[..., :3], o.read("Illumination")[..., :3])
        # the alias tables are built from maps that differ in the last bits, so a handful of sky/sun samples land on a
        # neighbouring texel: the > 1e-3 tail is allowed 2 %
        assert mean_rel <= 1e-3 and outliers <= 2e-2, (f, mean_rel, outliers)


def _hdr_test_image(w=96, h=64):
    rng = np.random.default_rng(3)
    img = np.zeros((h, w, 4), np.float32)
    img[..., :3] = np.exp(rng.uniform(-9.0, 3.0, (h, w, 3))).astype(np.float32)   # 1e-4 .. 20: spans the toe, the shoulder, the clamp
    img[0, :8, :3] = 0.0
    img[1, :8, :3] = np.float32(0.0031308 / 0.8) * np.linspace(0.5, 1.5, 8, dtype=np.float32)[:, None]  # around the sRGB knee
    return img


def test_output_stage_oracle_properties(oracle_lib):
    """FilmicToneMapping (manual exposure) + PNG conversion restatement: identities of the curve and of the conversion."""
    import vpt
    img = _hdr_test_image()
    p = vpt.default_tonemapping_params()
    assert float(p["manualExposure"][0]) == 10.0 and int(p["curve"][0]) == 0           # C++ defaults (GlobalSettings.h:148-167)
    p["manualExposure"] = 0.8
    rgb8, ldr = oracle_lib.tonemap(img, p)
    assert rgb8.shape == (64, 96, 3) and ldr.shape == (64, 96, 4)
    assert (ldr[..., :3] >= 0).all() and (ldr[..., :3] <= 1).all() and (ldr[..., 3] == 1).all()
    assert (rgb8[-1, :8] == 0).all()                                                    # black stays black; y is flipped (row 0 -> last row)
    assert np.array_equal(rgb8, (np.clip(ldr[::-1, :, :3], 0, 1) * np.float32(255.0)).astype(np.uint8))
    # grey ramp: monotone in the input for every curve
    ramp = np.zeros((1, 256, 4), np.float32); ramp[0, :, :3] = np.linspace(0, 8, 256, dtype=np.float32)[:, None]
    for curve in (0, 1, 2):
        p["curve"] = curve
        r8, rl = oracle_lib.tonemap(ramp, p)
        assert (np.diff(rl[0, :, 0]) >= -1e-6).all(), curve
    # saturation 0 -> grey; gain/lift act before the sRGB encode
    p["curve"] = 0; p["saturation"] = 0.0
    _, g = oracle_lib.tonemap(img, p)
    assert np.allclose(g[..., 0], g[..., 1], atol=1e-6) and np.allclose(g[..., 1], g[..., 2], atol=1e-6)


def test_output_stage_settings_loaders(tmp_path):
    import vpt
    y = tmp_path / "s.yaml"
    y.write_text("denoising:\n  phiLuminance: 3\npostprocess:\n  manualExposure: 0.8\n  toneMappingCurve: 2\n  whitePoint: 4\n  contrast: 1.1\n"
                 "  saturation: 0.9\n  gain: 1.05\n  lift: 0.01\n  enableBloom: true\nsky:\n  timeOfDay: 0.3\n  sunAxisAngle: 50\n  sunAxisRotate: 10\n  skyBrightness: 0.5\n")
    p, rc = vpt.load_tonemapping_settings(str(y))
    assert rc == 0
    assert (float(p["manualExposure"][0]), int(p["curve"][0]), float(p["whitePoint"][0])) == (np.float32(0.8), 2, 4.0)
    assert np.allclose([p["contrast"][0], p["saturation"][0], p["gain"][0], p["lift"][0]], [1.1, 0.9, 1.05, 0.01])
    s, rc = vpt.load_sky_settings(str(y))
    assert rc == 0 and np.allclose([s["timeOfDay"][0], s["sunAxisAngle"][0], s["sunAxisRotate"][0], s["skyBrightness"][0]], [0.3, 50, 10, 0.5])
    _, rc = vpt.load_tonemapping_settings(str(tmp_path / "missing.yaml"))
    assert rc != 0                                                                      # like LoadFromYAML returning false


@pytest.mark.gpu
def test_cuda_output_stage_matches_oracle(oracle_lib):
    """vpt_tonemap vs the oracle on a synthetic HDR frame, all three curves: 8-bit output identical up to +-1 LSB on a few
    samples (device powf vs glibc powf at a truncation boundary), IDENTICAL by the reference's own classification."""
    import imagediff
    import vpt
    img = _hdr_test_image(192, 128)
    g = vpt.Vpt(192, 128)
    g.write("IlluminationOutput", img)
    for curve, contrast in ((0, 1.0), (1, 1.0), (2, 1.2)):
        p = vpt.default_tonemapping_params()
        p["manualExposure"] = 0.8; p["curve"] = curve; p["contrast"] = contrast; p["whitePoint"] = 4.0
        g8, gl = g.tonemap(p)
        o8, ol = oracle_lib.tonemap(img, p)
        assert np.abs(gl - ol).max() < 2e-6, curve
        d = np.abs(g8.astype(np.int32) - o8.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 2e-3, (curve, int(d.max()), float((d > 0).mean()))
        assert imagediff.compare(g8, o8)["isIdentical"]
