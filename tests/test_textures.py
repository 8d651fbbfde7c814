"""Textured materials (SURVEY 8f "next" row 3; closesthit.cu:166-254, TextureManager.cu:82-115, 216-241): world-grid UV,
ray-cone LOD and trilinear fetches over uncompressed RGBA8 mip chains. CPU tests pin the oracle's software sampler through
its identities and the host mip-chain builder against a numpy restatement; the gpu tests compare the CUDA path with the oracle."""
import numpy as np
import pytest

import common
import vpt_scenes as S


def procedural_texture(n, seed, kind="albedo"):
    """Deterministic RGBA8 test image (n x n uint32, r = low byte)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:n, 0:n]
    if kind == "normal":
        # gentle bumps around +z: (0.5 + 0.5 n) * 255
        nx = 0.25 * np.sin(2 * np.pi * x / n * 3) + 0.05 * rng.standard_normal((n, n))
        ny = 0.25 * np.cos(2 * np.pi * y / n * 2) + 0.05 * rng.standard_normal((n, n))
        nz = np.sqrt(np.clip(1 - nx * nx - ny * ny, 0.05, 1))
        rgb = [np.clip((0.5 + 0.5 * c) * 255, 0, 255).astype(np.uint32) for c in (nx, ny, nz)]
    elif kind == "mono":
        v = np.clip(128 + 90 * np.sin(2 * np.pi * (x + 2 * y) / n * 2) + 20 * rng.standard_normal((n, n)), 0, 255).astype(np.uint32)
        rgb = [v, v, v]
    else:
        checker = ((x // max(n // 8, 1) + y // max(n // 8, 1)) & 1).astype(np.float64)
        rgb = [np.clip(60 + 150 * checker * s + 30 * rng.standard_normal((n, n)), 0, 255).astype(np.uint32) for s in (1.0, 0.8, 0.6)]
    return (rgb[0] | (rgb[1] << 8) | (rgb[2] << 16) | (np.uint32(255) << 24)).astype(np.uint32)


def numpy_mip_chain(level0):
    """TextureManager.cu:82-115 restated: 2x2 box average per channel, truncated; levels down to 4x4."""
    n = level0.shape[0]
    levels = max(int(np.log2(n)) - 1, 1)
    chain = [level0.astype(np.uint32)]
    for _ in range(1, levels):
        src = chain[-1]
        out = np.zeros((src.shape[0] // 2, src.shape[1] // 2), np.uint32)
        for ch in range(0, 32, 8):
            c = ((src >> ch) & 0xff).astype(np.float32)
            avg = np.minimum((c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2]) * np.float32(0.25), np.float32(255.0))
            out |= avg.astype(np.uint32) << ch
        chain.append(out)
    return chain


def textured_scene(vpt, n_mats, size=64):
    """Every material gets an albedo, a normal, a roughness and a metallic map (shared between materials round robin)."""
    kinds = ["albedo", "albedo", "normal", "normal", "mono", "mono"]
    textures = [vpt.build_mip_chain(procedural_texture(size if i % 2 == 0 else size // 2, 100 + i, k)) for i, k in enumerate(kinds)]
    slots = np.full((n_mats, 4), -1, np.int32)
    for m in range(n_mats):
        slots[m] = [m % 2, 2 + m % 2, 4 + m % 2, 5 - m % 2]
    slots[1, 3] = -1   # partially textured materials
    if n_mats > 2:
        slots[2, 1] = -1
    tex_size = np.tile(np.array([1024.0, 1024.0], np.float32), (n_mats, 1))
    return textures, slots, tex_size


def test_mip_chain_builder_matches_reference_restatement():
    import vpt
    for n in (4, 8, 64, 256):
        img = procedural_texture(n, n)
        chain = vpt.build_mip_chain(img)
        ref = numpy_mip_chain(img)
        assert len(chain) == len(ref) == max(int(np.log2(n)) - 1, 1)
        assert chain[-1].shape[0] == (4 if n >= 4 else n)
        for a, b in zip(chain, ref):
            assert np.array_equal(a, b)
    L = vpt.lib()
    assert L.vpt_mip_chain_texels(48) == 0 and L.vpt_mip_chain_texels(0) == 0
    assert L.vpt_mip_chain_texels(16) == 256 + 64 + 16


def test_oracle_sampler_identities(oracle_lib):
    """Software tex2DLod: texel centres reproduce texels, wrap addressing, constant textures stay constant at every LOD,
    LOD clamps to [0, levels-1], trilinear blend is linear in the LOD fraction."""
    import vpt
    O = oracle_lib
    o = O.Oracle(16, 16)
    img = procedural_texture(32, 7)
    chain = vpt.build_mip_chain(img)
    const = [np.full((8, 8), 0xFF4080C0, np.uint32), np.full((4, 4), 0xFF4080C0, np.uint32)]
    o.set_textures([chain, const], np.array([[0, -1, -1, -1]], np.int32), np.array([[32.0, 32.0]], np.float32))

    def rgba(v):
        return np.array([v & 0xff, (v >> 8) & 0xff, (v >> 16) & 0xff, v >> 24], np.float32) / np.float32(255.0)
    for (x, y) in ((0, 0), (5, 9), (31, 31), (17, 2)):
        u, v = (x + 0.5) / 32, (y + 0.5) / 32
        assert np.allclose(o.tex_sample(0, u, v, 0.0), rgba(int(img[y, x])), atol=1e-6)
        assert np.allclose(o.tex_sample(0, u + 3.0, v - 2.0, 0.0), rgba(int(img[y, x])), atol=2e-5)   # wrap
        assert np.allclose(o.tex_sample(0, u, v, -5.0), o.tex_sample(0, u, v, 0.0))                  # clamp low
    top = len(chain) - 1
    assert np.array_equal(o.tex_sample(0, 0.3, 0.7, top + 4.0), o.tex_sample(0, 0.3, 0.7, float(top)))  # clamp high
    a, b, mid = o.tex_sample(0, 0.3, 0.7, 1.0), o.tex_sample(0, 0.3, 0.7, 2.0), o.tex_sample(0, 0.3, 0.7, 1.25)
    assert np.allclose(mid, a + 0.25 * (b - a), atol=1e-6)
    # halfway between two texel centres = their average
    h = o.tex_sample(0, 6.0 / 32, 4.5 / 32, 0.0)
    assert np.allclose(h, 0.5 * (rgba(int(img[4, 5])) + rgba(int(img[4, 6]))), atol=1e-6)
    for lod in (0.0, 0.5, 1.0, 3.0):
        assert np.allclose(o.tex_sample(1, 0.123, 0.877, lod), rgba(0xFF4080C0), atol=1e-6)


def test_oracle_white_albedo_textures_equal_untextured_render(oracle_lib):
    """albedo * 1.0 and no other map: the textured code path must reproduce the untextured image (up to the rounding of the four
    bilinear weights, whose fp32 sum is 1 +- 1 ulp)."""
    O = oracle_lib
    W, H = 96, 64
    inp = common.scene_inputs((2, 1, 2))
    cam = common.scene_camera(W, H)
    a = common.setup(O.Oracle(W, H), inp, spp=1, total=3, diffuse=1)
    b = common.setup(O.Oracle(W, H), inp, spp=1, total=3, diffuse=1)
    n = len(inp["materials"])
    white = [np.full((8, 8), 0xFFFFFFFF, np.uint32), np.full((4, 4), 0xFFFFFFFF, np.uint32)]
    slots = np.full((n, 4), -1, np.int32)
    slots[:, 0] = 0
    b.set_textures([white], slots, np.full((n, 2), 1024.0, np.float32))
    a.render(cam, cam, 0)
    b.render(cam, cam, 0)
    assert np.array_equal(a.read("PrimaryHits"), b.read("PrimaryHits"))
    assert np.allclose(a.read("Albedo"), b.read("Albedo"), rtol=3e-7, atol=0)
    m, outl, _ = common.rel_err_stats(a.read("Illumination")[..., :3], b.read("Illumination")[..., :3])
    assert m <= 1e-6 and outl == 0.0, (m, outl)


def test_oracle_textured_render_changes_the_image(oracle_lib):
    import vpt
    O = oracle_lib
    W, H = 96, 64
    inp = common.scene_inputs((2, 1, 2))
    cam = common.scene_camera(W, H)
    a = common.setup(O.Oracle(W, H), inp, spp=1, total=3, diffuse=1)
    b = common.setup(O.Oracle(W, H), inp, spp=1, total=3, diffuse=1)
    b.set_textures(*textured_scene(vpt, len(inp["materials"])))
    a.render(cam, cam, 0)
    b.render(cam, cam, 0)
    assert np.array_equal(a.read("PrimaryHits"), b.read("PrimaryHits"))          # geometry does not depend on textures
    hit = a.read("PrimaryHits")[..., 3] >= 0
    da = np.abs(a.read("Albedo")[hit] - b.read("Albedo")[hit]).mean()
    dn = np.abs(a.read("NormalRoughness")[hit] - b.read("NormalRoughness")[hit]).mean()
    assert da > 0.02 and dn > 0.005, (da, dn)
    n = b.read("NormalRoughness")[hit][:, :3]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-3)


@pytest.mark.gpu
def test_textured_render_matches_oracle(oracle_lib):
    """cfg1-shaped textured scene, 2 spp, limits 3/1, two frames (temporal ReSTIR): primary hits exact, G-buffer maps and
    radiance within the north_star tolerance (mean relative error <= 1e-3)."""
    import vpt
    O = oracle_lib
    W, H = 256, 160
    inp = common.scene_inputs((2, 1, 2))
    cam = common.scene_camera(W, H)
    g = common.setup(vpt.Vpt(W, H), inp, spp=2, total=3, diffuse=1)
    o = common.setup(O.Oracle(W, H), inp, spp=2, total=3, diffuse=1)
    tex = textured_scene(vpt, len(inp["materials"]))
    g.set_textures(*tex)
    o.set_textures(*tex)
    for f in range(2):
        g.render(cam, cam, f)
        o.render(cam, cam, f)
        assert np.array_equal(g.read("PrimaryHits"), o.read("PrimaryHits"))
        assert np.array_equal(g.read("Depth"), o.read("Depth"))
        for name, tol in (("Albedo", 2e-5), ("NormalRoughness", 2e-4), ("MaterialParameter", 2e-5)):
            a, b = g.read(name), o.read(name)
            # a metallic map sitting exactly on its 0.5 threshold may flip one pixel's flag: allow a handful of pixels
            bad = (np.abs(a - b) > tol).any(-1).mean()
            assert bad <= 1e-3, (name, f, bad, float(np.abs(a - b).max()))
        m, outl, _ = common.rel_err_stats(g.read("Illumination")[..., :3], o.read("Illumination")[..., :3])
        assert m <= 1e-3 and outl <= 1e-2, (f, m, outl)
    # removing the textures restores the untextured render exactly as a fresh context produces it
    g.set_textures([], np.zeros((0, 4), np.int32), np.zeros((0, 2), np.float32))
    g2 = common.setup(vpt.Vpt(W, H), inp, spp=2, total=3, diffuse=1)
    g.render(cam, cam, 0)
    g2.render(cam, cam, 0)
    assert np.array_equal(g.read("Albedo"), g2.read("Albedo"))


@pytest.mark.gpu
def test_textured_specular_chain_matches_oracle(oracle_lib):
    """Ray-cone width must travel along continued paths (mirror-like roughness map -> depth > 0 texture fetches)."""
    import vpt
    O = oracle_lib
    W, H = 192, 128
    inp = common.scene_inputs((2, 1, 2))
    cam = common.scene_camera(W, H)
    n = len(inp["materials"])
    g = common.setup(vpt.Vpt(W, H), inp, spp=1, total=4, diffuse=1)
    o = common.setup(O.Oracle(W, H), inp, spp=1, total=4, diffuse=1)
    black = vpt.build_mip_chain(np.full((16, 16), 0xFF000000, np.uint32))      # roughness 0 everywhere -> specular
    albedo = vpt.build_mip_chain(procedural_texture(64, 3))
    slots = np.full((n, 4), -1, np.int32)
    slots[:, 0] = 1
    slots[:, 2] = 0
    ts = np.full((n, 2), 512.0, np.float32)
    g.set_textures([black, albedo], slots, ts)
    o.set_textures([black, albedo], slots, ts)
    g.render(cam, cam, 0)
    o.render(cam, cam, 0)
    assert np.array_equal(g.read("PrimaryHits"), o.read("PrimaryHits"))
    m, outl, _ = common.rel_err_stats(g.read("Illumination")[..., :3], o.read("Illumination")[..., :3])
    assert m <= 1e-3 and outl <= 1e-2, (m, outl)
