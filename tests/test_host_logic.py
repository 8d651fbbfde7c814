"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/vpt.h declares (no compute calls
without a GPU), the loaders mirror the reference parsers, host helpers agree with the oracle, failure is loud without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import vpt_scenes as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SETTINGS_YAML = """# Global Settings Configuration File
# Generated automatically

denoising:
  enableHitDistanceReconstruction: false
  enablePrePass: false
  enableTemporalAccumulation: true
  enableHistoryFix: yes
  enableHistoryClamping: on
  enableSpatialFiltering: 1
  enableFireflyFilter: TRUE
  maxAccumulatedFrameNum: 30
  maxFastAccumulatedFrameNum: 6
  phiLuminance: 2
  lobeAngleFraction: 0.5
  roughnessFraction: 0.15
  depthThreshold: 0.003
  atrousIterationNum: 1
  disocclusionThreshold: 0.01
  disocclusionThresholdAlternate: 0.05
  denoisingRange: 500000

postprocess:
  manualExposure: 0.8
  enableBloom: true
sky:
  timeOfDay: 0.25
"""

SCENE_YAML = """# Scene Configuration File
camera:
  position: [35.6184, 11.8733, 42.0387]
  direction: [-0.321564, -0.0129988, -0.946799]
  up: [0, 1, 0]
  fov: 90

character:
  position: [32.0, 10.0, 38.0]

chunk_config:
  chunksX: 4
  chunksY: 1
  chunksZ: 4
"""


def test_library_exports_every_declared_symbol():
    import vpt
    L = vpt.lib()
    header = open(os.path.join(ROOT, "include", "vpt.h")).read()
    declared = set(re.findall(r"\b(vpt_[a-z0-9_]+)\s*\(", header))
    declared.discard("vpt_ctx")
    assert declared == set(vpt.EXPORTS), declared ^ set(vpt.EXPORTS)
    for name in sorted(declared):
        assert hasattr(L, name), name


def test_pod_layouts_match_the_header():
    import oracle
    header = open(os.path.join(ROOT, "include", "vpt.h")).read()
    assert oracle.CAMERA_FLOATS * 4 == 212                       # Camera.h POD
    assert oracle.RESERVOIR_DTYPE.itemsize == 20 and oracle.ALIAS_DTYPE.itemsize == 12
    assert S.MATERIAL_DTYPE.itemsize == 48 and S.DENOISE_DTYPE.itemsize == 68
    fields = re.findall(r"^\s+(?:int32_t|float) (\w+);", header[header.index("typedef struct VptDenoisingParams"):header.index("} VptDenoisingParams")], re.M)
    assert tuple(fields) == S.DENOISE_DTYPE.names                 # same order as GlobalSettings.h:124-140 grouping in the C ABI
    for name, val in re.findall(r"VPT_BUF_(\w+) = (\d+)", header):
        assert oracle.BUF[name] == int(val)


def test_no_gpu_fails_loudly():
    """No CPU fallback: without a usable sm_100 device vpt_create returns VPT_ERR_CUDA and says why."""
    import torch
    import vpt
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vpt.VptError) as e:
        vpt.Vpt(64, 64)
    assert "vpt_create" in str(e.value) and "(2)" in str(e.value)


def test_settings_loader(tmp_path):
    import vpt
    f = tmp_path / "global_settings.yaml"
    f.write_text(SETTINGS_YAML)
    p, ok = vpt.load_denoising_settings(str(f))
    assert ok
    d = vpt.default_denoising_params()
    assert d["atrousIterationNum"][0] == 5 and p["atrousIterationNum"][0] == 1      # C++ default vs shipped YAML
    for k in ("enableTemporalAccumulation", "enableHistoryFix", "enableHistoryClamping", "enableSpatialFiltering", "enableFireflyFilter"):
        assert p[k][0] == 1                                                           # parseBool: true/1/yes/on, any case
    assert p["enableHitDistanceReconstruction"][0] == 0 and p["enablePrePass"][0] == 0
    assert np.isclose(p["depthThreshold"][0], 0.003) and p["denoisingRange"][0] == 500000.0 and p["maxAccumulatedFrameNum"][0] == 30.0
    assert p.tobytes() == S.default_denoising_params().tobytes()                      # == the params the benchmarks use
    q, ok2 = vpt.load_denoising_settings(str(tmp_path / "missing.yaml"))
    assert not ok2 and q.tobytes() == d.tobytes()                                     # returns false, values stay default
    ref_yaml = "/root/reference/data/settings/global_settings.yaml"
    if os.path.exists(ref_yaml):
        r, ok3 = vpt.load_denoising_settings(ref_yaml)
        assert ok3 and r.tobytes() == p.tobytes()


def test_scene_loader(tmp_path):
    import vpt
    f = tmp_path / "scene.yaml"
    f.write_text(SCENE_YAML)
    sc = vpt.load_scene_config(str(f))
    assert sc["loaded"] and np.allclose(sc["position"], [35.6184, 11.8733, 42.0387]) and sc["fov"] == 90.0
    assert abs(np.linalg.norm(sc["direction"]) - 1.0) < 1e-6 and sc["chunks"] == (4, 1, 4)       # direction normalised (SceneConfig.cpp:108)
    miss = vpt.load_scene_config(str(tmp_path / "nope.yaml"))
    assert not miss["loaded"] and np.allclose(miss["position"], [20, 15, 20])                    # CameraConfig defaults (SceneConfig.h:10-16)
    assert np.allclose(miss["direction"], np.array([-1, -0.3, -1]) / np.linalg.norm([-1, -0.3, -1]), atol=1e-6)
    cam = vpt.camera_from_scene(256, 256, sc["position"], sc["direction"], sc["fov"])
    assert np.allclose(cam[6:9], sc["position"]) and abs(cam[4] - 1.0) < 1e-6                    # tan(45 deg)


def test_alias_table_matches_oracle_and_is_a_distribution(oracle_lib):
    import vpt
    rng = np.random.default_rng(4)
    w = rng.random(4097).astype(np.float32) ** 4
    a, b = vpt.build_alias_table(w), oracle_lib.build_alias_table(w)
    assert a.tobytes() == b.tobytes()
    assert abs(a["p"].sum() - 1.0) < 1e-4 and (a["q"] <= 1.0 + 1e-6).all() and (a["q"] >= 0).all()
    # reconstruct the distribution from (q, alias): P(i) = (q_i + sum_{j: alias_j = i} (1 - q_j)) / n
    n = w.size
    rec = a["q"].astype(np.float64).copy()
    has = a["alias"] >= 0
    np.add.at(rec, a["alias"][has], 1.0 - a["q"][has].astype(np.float64))
    assert np.allclose(rec / n, w / w.sum(), atol=2e-6)


def test_sharding_helpers():
    import vpt_shard
    for n in (1, 2, 4, 8):
        got = sorted(k for r in range(n) for k in vpt_shard.samples_of(r, n, 4 * n))
        assert got == list(range(4 * n))
        for h in (1080, 2160, 544, 36):
            if h < 4 * n:
                continue
            b = vpt_shard.row_bands(h, n)
            assert b[0] == 0 and b[-1] == h and all(x % 4 == 0 for x in b[:-1]) and all(b[i] < b[i + 1] for i in range(n))
    with pytest.raises(ValueError):
        vpt_shard.row_bands(8, 4)


def test_vpt_offline_cli_help():
    import subprocess
    exe = os.path.join(ROOT, "real-time-path-tracing-voxel-blocks_b200", "vpt_offline")
    if not os.path.exists(exe):
        pytest.skip("vpt_offline not built")
    out = subprocess.run([exe, "--help"], capture_output=True, text=True, timeout=30)
    assert out.returncode == 0
    for flag in ("--width", "--height", "--output", "--scene", "--frames", "--test-canonical", "--update-canonical", "--canonical-image",
                 "--comment", "--test-sequence", "--test-remove20", "--test-remove-circle"):   # mainOffline.cpp:57-133
        assert flag in out.stdout, flag


def test_fastdiv_matches_integer_division():
    """csrc/vpt_fastdiv.h (index decoding in the kernels: path -> slot/sample, slot -> tile, voxel index -> x,y,z, light
    texel index -> x,y): the magic-number quotient/remainder equals // and % for every divisor the kernels can meet."""
    import ctypes as C
    import vpt
    L = vpt.lib()
    L.vpt_debug_fastdiv.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.vpt_debug_fastdiv.restype = None
    q, r = C.c_uint32(), C.c_uint32()
    rng = np.random.default_rng(7)
    divisors = [1, 2, 3, 5, 7, 30, 32, 34, 66, 128, 130, 160, 240, 258, 480, 960, 1024, 1026, 1056, 2073600, 8294400, 33177600, 2**31 - 1]
    for d in divisors:
        ns = np.concatenate([np.arange(0, 300), np.array([d - 1, d, d + 1, 2 * d - 1, 2 * d, 2**31 - 1, 2**32 - 1], np.uint64) % 2**32,
                             rng.integers(0, 2**32, 300)]).astype(np.uint64)
        for n in ns:
            L.vpt_debug_fastdiv(int(n), d, C.byref(q), C.byref(r))
            assert (q.value, r.value) == (int(n) // d, int(n) % d), (int(n), d, q.value, r.value)


def test_world_chunk_files_roundtrip(tmp_path):
    """WorldSceneManager chunk storage (WorldSceneManager.cpp:240-308, 310-458): FNV-1a-64 file names, raw 32768-byte chunks,
    the scene file's chunk_config / chunks sections; bad records are skipped like the reference does."""
    import vpt
    rng = np.random.default_rng(5)
    chunks = (2, 1, 2)
    ids = (rng.integers(0, 13, 4 * 32768) * (rng.random(4 * 32768) < 0.3)).astype(np.uint8)
    # known answer: FNV-1a 64 with the REFERENCE's offset basis 1469598103934665603 (WorldSceneManager.cpp:242 — the standard
    # basis 14695981039346656037 with its last digit missing; file names must match the reference's, so the quirk is kept)
    def fnv(b):
        h = 1469598103934665603
        for v in b.tobytes():
            h = ((h ^ v) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
        return "%016x" % h
    hashes = [fnv(ids[i * 32768:(i + 1) * 32768]) for i in range(4)]
    assert [vpt.chunk_hash(ids[i * 32768:(i + 1) * 32768]) for i in range(4)] == hashes
    scene = str(tmp_path / "scene.yaml")
    cam9 = np.array([35.6184, 11.8733, 42.0387, -0.321564, -0.0129988, -0.946799, 0, 1, 0], np.float32)
    assert vpt.save_world(scene, str(tmp_path), chunks, ids, cam9, 90.0) == 0
    for h in hashes:
        assert (tmp_path / (h + ".bin")).stat().st_size == 32768
    text = open(scene).read()
    assert "chunk_config:\n  chunksX: 2\n  chunksY: 1\n  chunksZ: 2\n" in text and ("  3: " + hashes[3]) in text
    # the scene loader reads the camera and chunk_config back
    sc = vpt.load_scene_config(scene)
    assert sc["loaded"] and np.allclose(sc["position"], cam9[:3], rtol=1e-5) and sc["fov"] == 90.0 and sc["chunks"] == chunks
    got = np.zeros_like(ids)
    rc, ok, bad = vpt.load_world(scene, str(tmp_path), chunks, got)
    assert (rc, ok, bad) == (0, 4, 0) and np.array_equal(got, ids)
    # a truncated file, a missing file and an out-of-range index are skipped; the other chunks still load
    (tmp_path / (hashes[1] + ".bin")).write_bytes(b"\0" * 100)
    (tmp_path / (hashes[2] + ".bin")).unlink()
    with open(scene, "a") as f:
        f.write("  9: %s\n" % hashes[0])
    got = np.full_like(ids, 255)
    rc, ok, bad = vpt.load_world(scene, str(tmp_path), chunks, got)
    assert rc == 0 and ok == 2 and bad == 3
    assert np.array_equal(got[:32768], ids[:32768]) and (got[32768:3 * 32768] == 255).all() and np.array_equal(got[3 * 32768:], ids[3 * 32768:])
    assert vpt.load_world(str(tmp_path / "nope.yaml"), str(tmp_path), chunks, got)[0] != 0


def test_image_diff_matches_reference_golden(tmp_path):
    """vpt_image_diff (the --test-canonical comparison of the offline entry) against the fixture generated by the reference's
    own ImageDiff.cpp (tests/golden/imagediff_ref.json, made by tests/golden/make_golden.py), in memory and through PNG files."""
    import json
    import vpt
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "imagediff_ref.json")))

    def words(img):
        u = img.astype(np.uint32)
        return u[..., 0] | (u[..., 1] << 8) | (u[..., 2] << 16) | (np.uint32(255) << 24)
    for i, c in enumerate(cases):
        r = np.random.default_rng(c["seed"])
        a = r.integers(0, 256, (c["h"], c["w"], 3), dtype=np.uint8)
        a = (a.astype(np.float32) * 0.25 + np.linspace(0, 180, c["w"], dtype=np.float32)[None, :, None]).astype(np.uint8)
        b = np.clip(a.astype(np.int32) + r.integers(-c["noise"], c["noise"] + 1, a.shape), 0, 255).astype(np.uint8)
        got = vpt.image_diff(words(a), words(b), 3)
        assert got["differentPixels"] == c["differentPixels"] and got["totalPixels"] == c["w"] * c["h"]
        assert abs(got["rmse"] - c["rmse"]) <= 1e-3 * max(1.0, c["rmse"])
        assert abs(got["ssim"] - c["ssim"]) <= 1e-4
        assert (bool(got["isIdentical"]), bool(got["isVeryClose"]), bool(got["isClose"])) == (c["isIdentical"], c["isVeryClose"], c["isClose"])
        if i < 3:
            PIL = pytest.importorskip("PIL.Image")
            pa, pb, pd = str(tmp_path / "a.png"), str(tmp_path / "b.png"), str(tmp_path / "d.png")
            PIL.fromarray(a, "RGB").save(pa); PIL.fromarray(b, "RGB").save(pb)
            got2 = vpt.image_diff_files(pa, pb, pd)
            assert got2 == got
            d = np.asarray(PIL.open(pd).convert("RGB")).astype(np.int32)
            want = np.minimum(np.abs(a.astype(np.int32) - b.astype(np.int32)) * 3, 255)   # ImageDiff.cpp:161-180
            assert np.array_equal(d, want)
    with pytest.raises(vpt.VptError):
        vpt.image_diff_files(str(tmp_path / "a.png"), str(tmp_path / "missing.png"))
    PIL = pytest.importorskip("PIL.Image")
    PIL.fromarray(np.zeros((5, 7, 3), np.uint8), "RGB").save(str(tmp_path / "s.png"))
    with pytest.raises(vpt.VptError):
        vpt.image_diff_files(str(tmp_path / "a.png"), str(tmp_path / "s.png"))      # different sizes


def test_dda_axis_choice_forms_agree():
    """The DDA engines pick the axis to step as ONE 3-input minimum plus equality predicates (z wins ties, then y, then x; csrc/vpt_dda.cu)
    where the oracle / VoxelEngine::performRayTraversal (voxelengine/VoxelEngine.cu:1133-1162) nest two '<' tests. Same axis and the
    same tCur for every combination of ties, infinities (axis-parallel rays: tMax = tDelta = FLT_MAX) and signed zeros (an origin
    on a voxel boundary with a negative step gives tMax = -0; +0 never occurs as an axis' tMax)."""
    rng = np.random.default_rng(5)
    n = 400_000
    special = np.array([-0.0, 1.0, 1.5, 2.0, 3.25, 1e-30, np.finfo(np.float32).max], np.float32)
    t = np.where(rng.random((n, 3)) < 0.5, rng.choice(special, size=(n, 3)), (rng.random((n, 3)) * 4).astype(np.float32)).astype(np.float32)
    tx, ty, tz = t[:, 0], t[:, 1], t[:, 2]
    xy = tx < ty
    ta = np.where(xy, tx, ty)
    az = ta < tz
    axis_ref = np.where(az, np.where(xy, 0, 1), 2)
    tcur_ref = np.where(az, ta, tz)
    m = np.minimum(np.minimum(tx, ty), tz)
    pz = tz == m
    py = (ty == m) & ~pz
    axis_new = np.where(pz, 2, np.where(py, 1, 0))
    assert np.array_equal(axis_ref, axis_new)
    assert np.array_equal(tcur_ref, m)                       # equal as values ...
    chosen = np.take_along_axis(t, axis_new[:, None], 1)[:, 0]
    assert np.array_equal(np.signbit(tcur_ref), np.signbit(chosen))   # ... and the reference's tCur carries the chosen axis' sign of zero
