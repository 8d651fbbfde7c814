"""-m gpu: the shared-memory tile kernels of the denoiser (TMA + mbarrier halo tiles for the a-trous passes, column-walking warps
for HistoryClamping; csrc/vpt_dn_tiles.cu) against the per-thread gather kernels they replace (VPT_DN_GATHER=1), and both against
the oracle. The a-trous tile kernels run the gather kernels' arithmetic tap for tap from a different data path, so their planes
are expected to agree to the last bit when fed the same input; the column-walking clamp sums its 5x5 moments in another order
(vertical ring, then lanes), so it agrees to rounding."""
import os

import numpy as np
import pytest

import common
import vpt_scenes as S

pytestmark = pytest.mark.gpu


def _ctx(vpt, w, h, inp, gather):
    old = os.environ.get("VPT_DN_GATHER")
    os.environ["VPT_DN_GATHER"] = "1" if gather else "0"
    os.environ.setdefault("VPT_DN_TILE_MASK", "7")   # all three tile kernels, the opt-in column-walking clamp included (read once per process)
    try:
        return common.setup(vpt.Vpt(w, h), inp, spp=1, total=3, diffuse=1)
    finally:
        if old is None:
            del os.environ["VPT_DN_GATHER"]
        else:
            os.environ["VPT_DN_GATHER"] = old


@pytest.mark.parametrize("size", [(256, 160), (332, 203), (64, 48)])
def test_tile_chain_matches_gather_chain_and_oracle(oracle_lib, size):
    """Sizes: tiles divide the image / ragged right and bottom edges (W % 32 != 0, H % 16 != 0) / an image smaller than the step-8
    tile, where every tile is a border tile. Moving camera from frame 3 on; 7 frames so every history-length branch runs."""
    import vpt
    W, H = size
    inp = common.scene_inputs((2, 1, 2))
    t = _ctx(vpt, W, H, inp, gather=False)
    g = _ctx(vpt, W, H, inp, gather=True)
    o = common.setup(oracle_lib.Oracle(W, H), inp, spp=1, total=3, diffuse=1)
    p = S.default_denoising_params()
    cam = common.scene_camera(W, H)
    prev = cam
    for f in range(7):
        if f >= 3:
            cam = vpt.camera_set_yaw_pitch(prev, prev[15] + np.float32(0.5 * np.pi / 180.0), prev[16])
        o.render(cam, prev, f)
        for c in (t, g):
            c.render(cam, prev, f)
            c.write("Illumination", o.read("Illumination"))          # identical noisy input on all three
            c.write_reservoirs(f & 1, o.read_reservoirs(f & 1))
            c.denoise(p, cam, prev, f, f + 1)
        o.denoise(p, cam, prev, f, f + 1)
        assert np.array_equal(t.read("HistoryLength"), g.read("HistoryLength")), f
        assert np.array_equal(t.read("HistoryLength"), o.read("HistoryLength")), f
        for name in ("PrevIllumination", "PrevFastIllumination", "IlluminationPing", "IlluminationOutput"):
            a, b = t.read(name), g.read(name)
            mre, tail, dmax = common.rel_err_stats(a, b)
            assert mre <= 2e-6 and tail <= 2e-4, (f, name, mre, tail, dmax)   # rounding order of the 5x5 moments only
            if name == "IlluminationOutput":   # the other planes against the oracle: tests/test_gpu_parity.py::test_denoiser_chain_matches_oracle
                # young-history frames (hl = 1, 2 with a static camera) sit on the hl < 3 / hl < 5 branches of the spatial passes, where
                # the fast class flips more borderline weights than in steady state: the cfg2 bar (tests/test_cfg2_parity.py)
                mre, tail, dmax = common.rel_err_stats(a, o.read(name))
                assert mre <= 2e-3 and tail <= 5e-2, (f, name, mre, tail, dmax)
        prev = cam


def test_atrous_tile_passes_are_bit_identical_to_the_gather_passes(oracle_lib):
    """Spatial filtering alone (temporal passes off: no HistoryClamping in front), so the only difference between the two contexts
    is the data path of the four a-trous passes: the output planes must be equal bit for bit."""
    import vpt
    W, H = 320, 192
    inp = common.scene_inputs((2, 1, 2))
    t = _ctx(vpt, W, H, inp, gather=False)
    g = _ctx(vpt, W, H, inp, gather=True)
    p = S.default_denoising_params().copy()
    p["enableTemporalAccumulation"] = 0
    cam = common.scene_camera(W, H)
    for f in range(3):
        for c in (t, g):
            c.render(cam, cam, f)
            c.denoise(p, cam, cam, f, f + 1)
        for name in ("IlluminationPing", "IlluminationOutput"):
            assert np.array_equal(t.read(name), g.read(name)), (f, name)
