"""VoxelEngine::performRayTraversal (voxelengine/VoxelEngine.cu:1040-1166) — the block picker behind the
scripted edit tests of mainOffline (--test-remove20 / --test-remove-circle). CPU: the oracle's restatement against the
generalised DDA (itself pinned against brute-force ray/AABB) and hand-checkable cases; gpu: vpt_pick_voxel == oracle, bit for bit."""
import numpy as np
import pytest

import common


def _terrain(ctx, chunks=(2, 1, 2)):
    inp = common.scene_inputs(chunks)
    common.setup(ctx, inp)
    return ctx


def _rays(n, dims, seed):
    rng = np.random.default_rng(seed)
    o = rng.random((n, 3)).astype(np.float32) * np.array(dims, np.float32)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d[: n // 8, 1] = -np.abs(d[: n // 8, 1]) * 4        # a good share looking down at the terrain
    d[n // 8: n // 6, 0] = 0.0                           # axis-parallel planes: tDelta = FLT_MAX
    d[n // 6: n // 5] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 5 - n // 6)] * rng.choice([-1.0, 1.0], (n // 5 - n // 6, 1)).astype(np.float32)
    d *= rng.random((n, 1)).astype(np.float32) * 3 + 0.1 # unnormalised like camera.dir may be
    return o, d


def test_oracle_picker_agrees_with_generalised_dda(oracle_lib):
    O = oracle_lib
    o = _terrain(O.Oracle(32, 32))
    W, H, D = 64, 32, 64
    origins, dirs = _rays(600, (W, H, D), 3)
    hits = 0
    for og, dr in zip(origins, dirs):
        pk = o.pick_voxel(og, dr)
        ref = o.dda(og, dr / np.float32(np.linalg.norm(dr)))
        if ref["hit"] and ref["steps"] < 1000:
            assert pk["hitSurface"] == 1 and pk["deletePos"] == ref["voxel"] and pk["deleteBlockId"] == ref["id"], (og, dr, pk, ref)
            hits += 1
        elif not ref["hit"]:
            assert pk["hitSurface"] == 0 and pk["deleteBlockId"] == -1
        if pk["hitSurface"] and pk["hasSpaceToCreate"]:
            # the placement cell is empty and face-adjacent to the hit voxel
            c, h = np.array(pk["createPos"]), np.array(pk["deletePos"])
            assert np.abs(c - h).sum() == 1
    assert hits > 150


def test_oracle_picker_edge_cases(oracle_lib):
    O = oracle_lib
    o = O.Oracle(32, 32)
    ids = np.zeros(2 * 1 * 2 * 32768, np.uint8)
    o.set_grid(2, 1, 2, ids)
    o.set_voxel(10, 5, 10, 7)
    # straight down onto the block from above: first solid = the block, placement = the cell above it
    pk = o.pick_voxel((10.5, 20.5, 10.5), (0.0, -1.0, 0.0))
    assert pk == dict(hasSpaceToCreate=1, hitSurface=1, createPos=(10, 6, 10), deletePos=(10, 5, 10), deleteBlockId=7)
    # unnormalised direction gives the same answer; zero direction finds nothing
    assert o.pick_voxel((10.5, 20.5, 10.5), (0.0, -250.0, 0.0)) == pk
    none = dict(hasSpaceToCreate=0, hitSurface=0, createPos=(-1, -1, -1), deletePos=(-1, -1, -1), deleteBlockId=-1)
    assert o.pick_voxel((10.5, 20.5, 10.5), (0.0, 0.0, 0.0)) == none
    # origin outside the grid: the walk ends before it starts (VoxelEngine.cu:1087-1093)
    assert o.pick_voxel((10.5, 40.0, 10.5), (0.0, -1.0, 0.0)) == none
    # looking away: leaves the grid, remembers the last empty cell, no hit
    pk = o.pick_voxel((10.5, 20.5, 10.5), (0.0, 1.0, 0.0))
    assert pk["hitSurface"] == 0 and pk["hasSpaceToCreate"] == 1 and pk["createPos"] == (10, 31, 10)
    # origin inside a solid voxel: that voxel is the hit and there is nothing to create
    pk = o.pick_voxel((10.5, 5.5, 10.5), (1.0, 0.0, 0.0))
    assert pk["hitSurface"] == 1 and pk["deletePos"] == (10, 5, 10) and pk["hasSpaceToCreate"] == 0
    # exact diagonal through cell corners: all three tMax tie at every corner; x steps only when strictly smallest and y only
    # when strictly smaller than z, so the order at each corner is z, y, x and the last empty cell before (12,12,12) is (11,12,12)
    o.set_voxel(12, 12, 12, 3)
    pk = o.pick_voxel((8.5, 8.5, 8.5), (1.0, 1.0, 1.0))
    assert pk["hitSurface"] == 1 and pk["deletePos"] == (12, 12, 12)
    assert pk["createPos"] == (11, 12, 12)
    # delete + re-pick walks on to the next block (what --test-remove20 does frame after frame)
    o.set_voxel(10, 2, 10, 9)
    o.set_voxel(10, 5, 10, 0)
    pk = o.pick_voxel((10.5, 20.5, 10.5), (0.0, -1.0, 0.0))
    assert pk["deletePos"] == (10, 2, 10) and pk["deleteBlockId"] == 9


@pytest.mark.gpu
def test_gpu_picker_matches_oracle(oracle_lib):
    import vpt
    O = oracle_lib
    g = _terrain(vpt.Vpt(64, 64))
    o = _terrain(O.Oracle(64, 64))
    origins, dirs = _rays(400, (64, 32, 64), 9)
    extra_o = [(10.5, 40.0, 10.5), (10.5, 20.5, 10.5), (-3.0, 5.0, 5.0), (63.999, 31.999, 63.999), (35.6184, 11.8733, 42.0387)]
    extra_d = [(0.0, -1.0, 0.0), (0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 1.0), (-0.321564, -0.0129988, -0.946799)]
    n_hit = 0
    for og, dr in list(zip(origins, dirs)) + list(zip(extra_o, extra_d)):
        a, b = g.pick_voxel(og, dr), o.pick_voxel(og, dr)
        assert a == b, (og, dr, a, b)
        n_hit += a["hitSurface"]
    assert n_hit > 100
    # scripted removal: 20 deletions along the scene camera's view ray, both sides stay in lock step
    cam_o, cam_d = (35.6184, 11.8733, 42.0387), (-0.321564, -0.0129988, -0.946799)
    removed = []
    for _ in range(20):
        a, b = g.pick_voxel(cam_o, cam_d), o.pick_voxel(cam_o, cam_d)
        assert a == b
        if not a["hitSurface"]:
            break
        removed.append(a["deletePos"])
        g.set_voxel(*a["deletePos"], 0)
        o.set_voxel(*a["deletePos"], 0)
    assert len(removed) >= 5 and len(set(removed)) == len(removed)
    assert np.array_equal(g.get_grid(), o.get_grid())


@pytest.mark.gpu
@pytest.mark.parametrize("flag,frames,expect", [("--test-remove20", 24, "REMOVAL TEST: Frame 20 deleting block #20"),
                                                ("--test-remove-circle", 12, "CIRCULAR TEST: Switching to view direction #3"),
                                                ("--test-sequence", 10, "TEST FRAME 8: Placing light block")])
def test_vpt_offline_scripted_edits(tmp_path, flag, frames, expect):
    """mainOffline's scripted block-edit runs (mainOffline.cpp:168-188, 279-393) through the offline entry: clicks are applied by
    the next frame's update via the picker, the edited world is what the following frames render."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "real-time-path-tracing-voxel-blocks_b200")
    (tmp_path / "scene.yaml").write_text("camera:\n  position: [35.6184, 11.8733, 42.0387]\n  direction: [-0.321564, -0.0129988, -0.946799]\n  up: [0, 1, 0]\n  fov: 90\n")
    (tmp_path / "settings.yaml").write_text("denoising:\n  atrousIterationNum: 1\npostprocess:\n  manualExposure: 0.8\n")
    args = [os.path.join(pkg, "vpt_offline"), "--width", "160", "--height", "96", "--frames", str(frames), "--scene", str(tmp_path / "scene.yaml"),
            "--settings", str(tmp_path / "settings.yaml"), "--tables", os.path.join(pkg, "data", "bluenoise_tables.bin"),
            "--sky-tables", os.path.join(pkg, "data", "sky_tables.bin"), "--output", str(tmp_path / "o"), flag]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert expect in r.stdout, r.stdout[-3000:]
    m = re.search(r"Scripted edits: (\d+) blocks removed, (\d+) placed", r.stdout)
    assert m, r.stdout[-2000:]
    removed, placed = int(m.group(1)), int(m.group(2))
    if flag == "--test-remove20":
        assert removed == 20 and placed == 0
        hits = re.findall(r"Hit block at \((\d+),(\d+),(\d+)\)", r.stdout)
        assert len(hits) == 20 and len(set(hits)) == 20            # every click removes a different block along the view ray
    elif flag == "--test-remove-circle":
        assert removed == 11 and placed == 0                        # clicks after frames 1..11 are consumed by frames 2..12
    else:
        assert (removed, placed) == (1, 2)                          # place (frame 3), remove (frame 6), place (frame 9)
    assert os.path.exists(str(tmp_path / "o_0003.png"))


@pytest.mark.gpu
def test_vpt_offline_canonical_image_flow(tmp_path):
    """mainOffline's --update-canonical / --test-canonical (mainOffline.cpp:422-497): the last frame becomes the canonical image,
    a second identical run is IDENTICAL to it (the render is deterministic), a run with other settings is reported as different."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "real-time-path-tracing-voxel-blocks_b200")
    (tmp_path / "scene.yaml").write_text("camera:\n  position: [35.6184, 11.8733, 42.0387]\n  direction: [-0.321564, -0.0129988, -0.946799]\n  up: [0, 1, 0]\n  fov: 90\n")
    (tmp_path / "settings.yaml").write_text("denoising:\n  atrousIterationNum: 1\npostprocess:\n  manualExposure: 0.8\n")
    canon = str(tmp_path / "canonical.png")
    base = [os.path.join(pkg, "vpt_offline"), "--width", "192", "--height", "128", "--frames", "4", "--scene", str(tmp_path / "scene.yaml"),
            "--settings", str(tmp_path / "settings.yaml"), "--tables", os.path.join(pkg, "data", "bluenoise_tables.bin"),
            "--sky-tables", os.path.join(pkg, "data", "sky_tables.bin"), "--canonical-image", canon]
    r0 = subprocess.run(base + ["--output", str(tmp_path / "a"), "--test-canonical"], capture_output=True, text=True, timeout=300)
    assert r0.returncode == 0 and "Canonical image not found" in r0.stdout, r0.stdout[-2000:]
    rep = str(tmp_path / "perf.txt")
    r1 = subprocess.run(base + ["--output", str(tmp_path / "a"), "--update-canonical", "--perf-report", rep, "--comment", "unit test run"],
                        capture_output=True, text=True, timeout=300)
    lines = open(rep).read().splitlines()
    assert lines[0].startswith("# Performance Report") and len(lines) == 4 and lines[3].endswith("unit test run") and "192x128" in lines[3]
    assert lines[3].count("|") == 10
    assert r1.returncode == 0 and "Canonical image updated" in r1.stdout and os.path.exists(canon), r1.stdout[-2000:]
    r2 = subprocess.run(base + ["--output", str(tmp_path / "b"), "--test-canonical"], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0 and "Assessment: IDENTICAL" in r2.stdout, r2.stdout[-2000:]
    assert "Image matches canonical reference" in r2.stdout and os.path.exists(str(tmp_path / "b_diff.png"))
    r3 = subprocess.run(base + ["--output", str(tmp_path / "c"), "--test-canonical", "--exposure", "3.0"], capture_output=True, text=True, timeout=300)
    assert r3.returncode == 0 and "Assessment: IDENTICAL" not in r3.stdout and "Different pixels:" in r3.stdout, r3.stdout[-2000:]
