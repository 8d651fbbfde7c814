"""Synthetic inputs of the named shapes (BASELINE.json configs / SURVEY §8d) — pure numpy, no device code.

Everything here is INPUT DATA fed identically to the CUDA path and to the oracle:
  * blue-noise tables (data/bluenoise_tables.bin, exported from the reference's RandGenData.h:15-39)
  * the terrain block -> material table (reference data/assets/blocks.yaml + materials.yaml, texture-less:
    roughness from the yaml, albedo = a flat stand-in colour because BC textures are out of scope)
  * an analytic stand-in sky/sun (the Hosek-Wilkie model is "next" row #1): equal-area sphere map
    1024x512 + 32x32 sun disk, sun direction as the reference computes it for timeOfDay 0.25 / axis 45 deg
    (renderer/sky/Sky.cu:362-366)
  * the scene camera of data/scene/scene_export.yaml
  * the synthetic G-buffer of config 3 (denoiser alone)
"""
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DATA_DIR = os.path.join(os.path.dirname(_HERE), "data")

MATERIAL_DTYPE = np.dtype([("albedo", "<f4", 3), ("roughness", "<f4"), ("translucency", "<f4"), ("uvScale", "<f4"),
                           ("metallic", "<i4"), ("materialId", "<i4"), ("useWorldGridUV", "<i4"), ("isEmissive", "<i4"),
                           ("isThinfilm", "<i4"), ("pad", "<i4")])
DENOISE_DTYPE = np.dtype([("enableHitDistanceReconstruction", "<i4"), ("enablePrePass", "<i4"), ("enableTemporalAccumulation", "<i4"),
                          ("enableHistoryFix", "<i4"), ("enableHistoryClamping", "<i4"), ("enableSpatialFiltering", "<i4"),
                          ("enableFireflyFilter", "<i4"), ("maxAccumulatedFrameNum", "<f4"), ("maxFastAccumulatedFrameNum", "<f4"),
                          ("phiLuminance", "<f4"), ("lobeAngleFraction", "<f4"), ("roughnessFraction", "<f4"), ("depthThreshold", "<f4"),
                          ("atrousIterationNum", "<i4"), ("disocclusionThreshold", "<f4"), ("disocclusionThresholdAlternate", "<f4"),
                          ("denoisingRange", "<f4")])

# data/scene/scene_export.yaml
SCENE_CAMERA = dict(position=(35.6184, 11.8733, 42.0387), direction=(-0.321564, -0.0129988, -0.946799), fov=90.0)


def load_tables():
    t = np.fromfile(os.path.join(DATA_DIR, "bluenoise_tables.bin"), dtype=np.uint8)
    assert t.size == 327680, "bluenoise_tables.bin is corrupt"
    return t


def load_sky_tables():
    """Hosek-Wilkie coefficient datasets (data/sky_tables.bin, exported from the reference's SkyData.h by tools/export_sky_tables.py)."""
    t = np.fromfile(os.path.join(DATA_DIR, "sky_tables.bin"), dtype="<f4")
    assert t.size == 2460, "sky_tables.bin is corrupt"
    return t


DEFAULT_SKY_PARAMS = (0.25, 45.0, 0.0, 1.0)  # SkyParams defaults: timeOfDay, sunAxisAngle, sunAxisRotate, skyBrightness (GlobalSettings.h:200-203)


def default_materials():
    """12 terrain materials (ids 0..11 = materials.yaml order) and the block id -> material index map
    (block id i -> material i-1, blocks.yaml ids 1..12)."""
    rough = [0.8, 0.9, 0.85, 0.9, 0.8, 0.7, 0.85, 0.6, 0.7, 0.65, 0.75, 0.75]
    albedo = [(0.76, 0.70, 0.50), (0.45, 0.33, 0.22), (0.55, 0.52, 0.48), (0.40, 0.30, 0.20), (0.76, 0.70, 0.50), (0.80, 0.75, 0.60),
              (0.50, 0.50, 0.52), (0.60, 0.60, 0.58), (0.62, 0.58, 0.52), (0.78, 0.74, 0.66), (0.55, 0.40, 0.25), (0.55, 0.40, 0.25)]
    m = np.zeros(12, MATERIAL_DTYPE)
    for i in range(12):
        m[i]["albedo"] = albedo[i]
        m[i]["roughness"] = rough[i]
        m[i]["uvScale"] = 2.5
        m[i]["useWorldGridUV"] = 1
        m[i]["materialId"] = i
    b2m = np.zeros(256, np.uint16)
    for block in range(1, 13):
        b2m[block] = block - 1
    return m, b2m


def default_denoising_params(yaml_overrides=True):
    """C++ defaults of DenoisingParams (GlobalSettings.h:124-140); yaml_overrides applies the shipped
    data/settings/global_settings.yaml (atrousIterationNum: 1)."""
    p = np.zeros(1, DENOISE_DTYPE)
    p["enableTemporalAccumulation"] = 1
    p["enableHistoryFix"] = 1
    p["enableHistoryClamping"] = 1
    p["enableSpatialFiltering"] = 1
    p["enableFireflyFilter"] = 1
    p["maxAccumulatedFrameNum"] = 30.0
    p["maxFastAccumulatedFrameNum"] = 6.0
    p["phiLuminance"] = 2.0
    p["lobeAngleFraction"] = 0.5
    p["roughnessFraction"] = 0.15
    p["depthThreshold"] = 0.003
    p["atrousIterationNum"] = 1 if yaml_overrides else 5
    p["disocclusionThreshold"] = 0.01
    p["disocclusionThresholdAlternate"] = 0.05
    p["denoisingRange"] = 500000.0
    return p


def reference_sun_dir():
    v = np.array([0.70710678, 0.5, -0.5], np.float64)
    return (v / np.linalg.norm(v)).astype(np.float32)


def synthetic_sky(sky_w=1024, sky_h=512, sun_w=32, sun_h=32, constant=None):
    """Returns (sky[h,w,4], sun[h,w,4], sun_dir). `constant` -> uniform sky of that radiance and a black sun
    (furnace / bit-exact gates)."""
    sun_dir = reference_sun_dir()
    sky = np.zeros((sky_h, sky_w, 4), np.float32)
    sun = np.zeros((sun_h, sun_w, 4), np.float32)
    if constant is not None:
        sky[..., :3] = np.float32(constant)
        return sky, sun, sun_dir
    # equal-area sphere map: y = 2v-1, phi = 2*pi*u  (LinearMath.h:1858-1864)
    v = (np.arange(sky_h, dtype=np.float64) + 0.5) / sky_h
    u = (np.arange(sky_w, dtype=np.float64) + 0.5) / sky_w
    y = (2.0 * v - 1.0)[:, None]
    r = np.sqrt(np.maximum(0.0, 1.0 - y * y))
    phi = 2.0 * np.pi * u[None, :]
    d = np.stack([r * np.cos(phi), np.broadcast_to(y, (sky_h, sky_w)), r * np.sin(phi)], -1)
    zen = np.array([0.20, 0.38, 0.90])
    hor = np.array([0.75, 0.82, 0.95])
    gnd = np.array([0.22, 0.21, 0.20])
    t = np.clip(d[..., 1], 0.0, 1.0)[..., None] ** 0.45
    up = hor * (1.0 - t) + zen * t
    cosg = np.clip((d * sun_dir.astype(np.float64)).sum(-1), -1.0, 1.0)[..., None]
    glow = 0.9 * np.array([1.0, 0.85, 0.6]) * np.exp((cosg - 1.0) * 24.0)
    col = np.where(d[..., 1:2] >= 0.0, up + glow, gnd * (1.0 + 0.5 * np.clip(d[..., 1:2] + 0.2, 0.0, 1.0)))
    sky[..., :3] = col.astype(np.float32)
    # sun disk texels: mild limb darkening across u (u=0 is the disk centre, LinearMath.h:1871-1885)
    uu = (np.arange(sun_w, dtype=np.float64) + 0.5) / sun_w
    limb = (1.0 - 0.4 * uu)[None, :, None]
    sun[..., :3] = (np.array([52000.0, 47000.0, 40000.0]) * limb).astype(np.float32)
    return sky, sun, sun_dir


def luminance(rgb):
    return 0.2126 * rgb[..., 0] + 0.7152 * rgb[..., 1] + 0.0722 * rgb[..., 2]


def sky_pdf_weights(tex):
    """Alias-table weights = luminance of each texel (renderer/sky/Sky.cu:279,302,326), float32."""
    t = tex.astype(np.float32)
    w = np.float32(0.2126) * t[..., 0] + np.float32(0.7152) * t[..., 1] + np.float32(0.0722) * t[..., 2]
    return np.ascontiguousarray(w, np.float32).ravel()


def synthetic_gbuffer(width, height, frame, cam_pos=(0.0, 6.0, 0.0)):
    """Config 3: analytic depth = ground plane + 2 boxes seen from a pinhole at cam_pos looking
    along +z, piecewise-constant normals, materialId in {1,2,3,7}, albedo 0.5, noisy radiance
    base*(1+0.5*(u-0.5)) with u = SequenceHash(px + W*py + 0x9E3779B9*frame)/2^32. Returns dict of arrays."""
    ys, xs = np.mgrid[0:height, 0:width]
    aspect = height / width
    fovx = np.pi / 2
    tx = np.tan(fovx / 2)
    ty = np.tan(fovx * aspect / 2)
    dx = -(2.0 * (xs + 0.5) / width - 1.0) * tx      # camera "left" convention of the reference (x flips)
    dy = (2.0 * (ys + 0.5) / height - 1.0) * ty
    dz = np.ones_like(dx)
    n = np.sqrt(dx * dx + dy * dy + dz * dz)
    dx, dy, dz = dx / n, dy / n, dz / n
    ox, oy, oz = cam_pos
    t_best = np.full((height, width), 1.0e27)
    normal = np.zeros((height, width, 3))
    normal[..., 1] = -1.0
    mat = np.full((height, width), 65535.0)
    # ground plane y = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(dy < -1e-6, (0.0 - oy) / dy, np.inf)
    hit = tg < t_best
    t_best = np.where(hit, tg, t_best)
    normal[hit] = (0.0, 1.0, 0.0)
    tgf = np.where(np.isfinite(tg), tg, 0.0)
    px = ox + dx * tgf
    pz = oz + dz * tgf
    checker = ((np.floor(px / 4.0) + np.floor(pz / 4.0)) % 2 == 0)
    mat = np.where(hit, np.where(checker, 1.0, 2.0), mat)
    # two axis-aligned boxes
    for (bmin, bmax, mid) in (((-9.0, 0.0, 14.0), (-3.0, 7.0, 20.0), 3.0), ((2.0, 0.0, 10.0), (8.0, 4.0, 16.0), 7.0)):
        tmin = np.full((height, width), -np.inf)
        tmax = np.full((height, width), np.inf)
        axis_n = np.zeros((height, width), np.int64)
        for a, (o, dcomp) in enumerate(((ox, dx), (oy, dy), (oz, dz))):
            with np.errstate(divide="ignore", invalid="ignore"):
                t0 = (bmin[a] - o) / dcomp
                t1 = (bmax[a] - o) / dcomp
            tn, tf = np.minimum(t0, t1), np.maximum(t0, t1)
            axis_n = np.where(tn > tmin, a, axis_n)
            tmin = np.maximum(tmin, tn)
            tmax = np.minimum(tmax, tf)
        bh = (tmin <= tmax) & (tmin > 0) & (tmin < t_best)
        t_best = np.where(bh, tmin, t_best)
        dcomps = np.stack([dx, dy, dz], -1)
        nn = np.zeros((height, width, 3))
        sign = -np.sign(np.take_along_axis(dcomps, axis_n[..., None], -1))[..., 0]
        for a in range(3):
            nn[..., a] = np.where(axis_n == a, sign, 0.0)
        normal[bh] = nn[bh]
        mat = np.where(bh, mid, mat)
    sky = t_best > 5.0e5
    depth = np.where(sky, 1.0e27, t_best).astype(np.float32)
    nr = np.zeros((height, width, 4), np.float32)
    nr[..., :3] = normal
    nr[..., 3] = np.where(sky, 0.0, 0.8)
    albedo = np.zeros((height, width, 4), np.float32)
    albedo[..., :3] = np.where(sky[..., None], 1.0, 0.5)
    albedo[..., 3] = 1.0
    # hash noise
    x = (xs + width * ys + 0x9E3779B9 * frame).astype(np.uint64) & 0xFFFFFFFF
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x.astype(np.uint64) * 0x7FEB352D & 0xFFFFFFFF).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x.astype(np.uint64) * 0x846CA68B & 0xFFFFFFFF).astype(np.uint32)
    x ^= x >> np.uint32(16)
    u = x.astype(np.float64) / 4294967296.0
    base = np.where(mat == 1.0, 0.9, np.where(mat == 2.0, 0.35, np.where(mat == 3.0, 0.6, 0.5)))
    shade = np.clip(normal[..., 1] * 0.6 + 0.5, 0.1, 1.2)
    rad = base * shade * (1.0 + 0.5 * (u - 0.5))
    illum = np.zeros((height, width, 4), np.float32)
    illum[..., 0] = rad
    illum[..., 1] = rad * 0.95
    illum[..., 2] = rad * 0.85
    illum[sky] = (0.5, 0.6, 0.9, 0.0)
    illum[..., 3] = depth
    gnt = nr.copy()
    gnt[..., 3] = 0.0
    mp = np.zeros((height, width, 4), np.float32)
    return dict(Depth=depth, NormalRoughness=nr, Material=mat.astype(np.float32), Albedo=albedo, Illumination=illum,
                GeoNormalThinfilm=gnt, MaterialParameter=mp)
