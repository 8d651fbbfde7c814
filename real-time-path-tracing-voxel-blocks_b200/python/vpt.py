"""ctypes binding of libvpt.so (include/vpt.h) — the product path. No CPU fallback: if the shared library is
missing or no sm_100-class GPU is present every compute call raises VptError.

The class mirrors the reference's call sequence on this path (mainOffline.cpp:142-345):
    ctx = Vpt(width, height)                      # OfflineBackend::init + BufferManager::init + OptixRenderer::init
    ctx.set_tables(...); ctx.generate_terrain(...) / ctx.set_grid(...); ctx.set_materials(...); ctx.set_sky(...)
    ctx.render(cam, prev_cam, iteration_index)     # OptixRenderer::render
    ctx.denoise(params, cam, prev_cam, frame_num, iteration_index + 1)   # Denoiser::run
    ctx.read("IlluminationOutput")
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VPT_LIB") or os.path.join(os.path.dirname(_HERE), "libvpt.so")  # VPT_LIB: A/B builds while tuning
_LIB = None

BUF = dict(
    Illumination=0, IlluminationOutput=1, IlluminationPing=2, IlluminationPong=3, NormalRoughness=4, Depth=5,
    Material=6, Albedo=7, HistoryLength=8, PrevDepth=9, PrevMaterial=10, PrevIllumination=11,
    PrevFastIllumination=12, PrevHistoryLength=13, PrevNormalRoughness=14, GeoNormalThinfilm=15,
    MaterialParameter=16, PrevMaterialParameter=17, PrevGeoNormalThinfilm=18, PrevAlbedo=19, PrimaryHits=22)
_F1 = {"Depth", "Material", "HistoryLength", "PrevDepth", "PrevMaterial", "PrevHistoryLength"}

RESERVOIR_DTYPE = np.dtype([("lightData", "<u4"), ("uvData", "<u4"), ("weightSum", "<f4"), ("targetPdf", "<f4"), ("M", "<f4")])
ALIAS_DTYPE = np.dtype([("q", "<f4"), ("p", "<f4"), ("alias", "<i4")])
TIMINGS_DTYPE = np.dtype([("trace_ms", "<f4"), ("resolve_ms", "<f4"), ("firefly_ms", "<f4"), ("temporal_ms", "<f4"),
                          ("history_fix_ms", "<f4"), ("history_clamp_ms", "<f4"), ("atrous_smem_ms", "<f4"), ("atrous_ms", "<f4"),
                          ("composite_ms", "<f4"), ("denoise_total_ms", "<f4"), ("atrous_passes", "<i4"), ("kernel_launches", "<i4"),
                          ("trace_dda_ms", "<f4"), ("trace_shade_ms", "<f4"), ("trace_dda_launches", "<i4"), ("trace_shade_launches", "<i4")])
LIGHT_DTYPE = np.dtype([("center", "<f4", 3), ("scalars", "<u4"), ("radiance", "<u4", 2), ("direction1", "<u4"), ("direction2", "<u4")])
CAMERA_FLOATS = 53

# every symbol include/vpt.h declares (checked by the CPU test-suite against the built library)
EXPORTS = [
    "vpt_create", "vpt_destroy", "vpt_last_error", "vpt_sync", "vpt_stream", "vpt_set_tables", "vpt_set_grid", "vpt_get_grid",
    "vpt_set_voxel", "vpt_generate_terrain", "vpt_set_materials", "vpt_set_sky", "vpt_set_trace_params", "vpt_render",
    "vpt_render_shard", "vpt_resolve", "vpt_denoise", "vpt_begin_external_frame", "vpt_denoise_external", "vpt_read_buffer",
    "vpt_write_buffer", "vpt_read_reservoirs", "vpt_write_reservoirs", "vpt_device_ptr", "vpt_get_counters", "vpt_get_timings",
    "vpt_set_profiling", "vpt_comm_unique_id", "vpt_comm_init", "vpt_comm_allreduce_illumination", "vpt_comm_broadcast_gbuffer",
    "vpt_denoise_band", "vpt_camera_init", "vpt_camera_update", "vpt_camera_from_scene", "vpt_perlin_noise_chunks",
    "vpt_build_alias_table", "vpt_load_denoising_settings", "vpt_default_denoising_params", "vpt_load_scene_config",
    "vpt_debug_fastdiv", "vpt_generate_sky", "vpt_read_sky", "vpt_sky_size", "vpt_sky_state",
    "vpt_chunk_hash", "vpt_save_world", "vpt_load_world", "vpt_set_wave_budget", "vpt_read_buffer_async", "vpt_read_wait", "vpt_tonemap", "vpt_default_tonemapping_params", "vpt_load_tonemapping_settings", "vpt_load_sky_settings",
    "vpt_set_textures", "vpt_mip_chain_texels", "vpt_build_mip_chain", "vpt_load_materials", "vpt_load_png_rgba8", "vpt_pick_voxel", "vpt_render_shard_local", "vpt_image_diff", "vpt_image_diff_files", "vpt_get_total_rays", "vpt_debug_tma_timeouts", "vpt_debug_read_wave", "vpt_get_lights", "vpt_band_rows", "vpt_band_input_halo", "vpt_comm_gather_output", "vpt_write_buffer_device", "vpt_render_range"]


def pack_textures(textures, slots, tex_size):
    """Flatten a list of mip chains into the arrays vpt_set_textures takes (shared with the oracle binding)."""
    widths = np.array([t[0].shape[0] for t in textures], np.int32)
    levels = np.array([len(t) for t in textures], np.int32)
    texels = np.concatenate([np.ascontiguousarray(l, np.uint32).ravel() for t in textures for l in t]) if textures else np.zeros(1, np.uint32)
    slots = np.ascontiguousarray(slots, np.int32).reshape(-1, 4)
    tex_size = np.ascontiguousarray(tex_size, np.float32).reshape(-1, 2)
    assert slots.shape[0] == tex_size.shape[0]
    return widths, levels, np.ascontiguousarray(texels), slots, tex_size


def load_png_rgba8(path):
    """vpt_load_png_rgba8 -> ((h,w) uint32 RGBA8, channels)."""
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    rc = lib().vpt_load_png_rgba8(path.encode(), None, C.c_size_t(0), C.byref(w), C.byref(h), C.byref(ch))
    if rc != 0:
        raise VptError("vpt_load_png_rgba8(%s) failed (%d)" % (path, rc))
    out = np.zeros((h.value, w.value), np.uint32)
    rc = lib().vpt_load_png_rgba8(path.encode(), _p(out), C.c_size_t(out.size), C.byref(w), C.byref(h), C.byref(ch))
    if rc != 0:
        raise VptError("vpt_load_png_rgba8(%s) failed (%d)" % (path, rc))
    return out, ch.value


def load_materials(materials_yaml, blocks_yaml=None, max_materials=256):
    """vpt_load_materials -> (materials structured array, block->material uint16[256], list of 4-tuples of texture paths)."""
    import vpt_scenes as S
    mats = np.zeros(max_materials, S.MATERIAL_DTYPE)
    paths = np.zeros((max_materials, 4, 256), np.uint8)
    b2m = np.zeros(256, np.uint16)
    n = C.c_int()
    rc = lib().vpt_load_materials(materials_yaml.encode(), blocks_yaml.encode() if blocks_yaml else None, _p(mats), _p(paths), max_materials,
                                  C.byref(n), _p(b2m))
    if rc != 0:
        raise VptError("vpt_load_materials failed (%d)" % rc)
    names = [tuple(bytes(paths[i, k]).split(b"\0")[0].decode() for k in range(4)) for i in range(n.value)]
    return mats[:n.value].copy(), b2m, names


IMAGE_DIFF_DTYPE = np.dtype([("differentPixels", "<i4"), ("totalPixels", "<i4"), ("pixelDifferenceRatio", "<f4"), ("rmse", "<f4"), ("ssim", "<f4"),
                             ("isIdentical", "<i4"), ("isVeryClose", "<i4"), ("isClose", "<i4")])


def image_diff(a, b, channels=3):
    """vpt_image_diff on two (h,w) uint32 RGBA8 images -> dict of ImageDiffResult fields."""
    a = np.ascontiguousarray(a, np.uint32); b = np.ascontiguousarray(b, np.uint32)
    r = np.zeros(1, IMAGE_DIFF_DTYPE)
    rc = lib().vpt_image_diff(_p(a), _p(b), a.shape[1], a.shape[0], channels, _p(r))
    if rc != 0:
        raise VptError("vpt_image_diff failed (%d)" % rc)
    return {k: r[k][0].item() for k in IMAGE_DIFF_DTYPE.names}


def image_diff_files(path_a, path_b, diff_png=None):
    r = np.zeros(1, IMAGE_DIFF_DTYPE)
    rc = lib().vpt_image_diff_files(path_a.encode(), path_b.encode(), _p(r), diff_png.encode() if diff_png else None)
    if rc != 0:
        raise VptError("vpt_image_diff_files failed (%d)" % rc)
    return {k: r[k][0].item() for k in IMAGE_DIFF_DTYPE.names}


def build_mip_chain(level0):
    """vpt_build_mip_chain on one (n,n) uint32 RGBA8 image -> list of levels (host only)."""
    level0 = np.ascontiguousarray(level0, np.uint32)
    n = level0.shape[0]
    total = lib().vpt_mip_chain_texels(n)
    if total <= 0:
        raise VptError("vpt_mip_chain_texels(%d) failed: textures must be square powers of two" % n)
    out = np.zeros(total, np.uint32)
    nl = lib().vpt_build_mip_chain(_p(level0), n, _p(out))
    if nl <= 0:
        raise VptError("vpt_build_mip_chain failed")
    chain, off = [], 0
    for l in range(nl):
        w = n >> l
        chain.append(out[off:off + w * w].reshape(w, w).copy())
        off += w * w
    return chain


class VptError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise VptError("libvpt.so is not built (%s); run __graft_entry__.build() — there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.vpt_last_error.restype = C.c_char_p
        L.vpt_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.vpt_destroy.argtypes = [C.c_void_p]
        L.vpt_stream.restype = C.c_void_p
        L.vpt_stream.argtypes = [C.c_void_p]
        L.vpt_device_ptr.restype = C.c_void_p
        L.vpt_device_ptr.argtypes = [C.c_void_p, C.c_int]
        for n in ("vpt_read_buffer", "vpt_write_buffer"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        for n in ("vpt_read_reservoirs", "vpt_write_reservoirs"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.vpt_get_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.vpt_camera_from_scene.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float]
        L.vpt_load_denoising_settings.argtypes = [C.c_char_p, C.c_void_p]
        L.vpt_load_scene_config.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vpt_denoise_external.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.vpt_denoise.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.vpt_denoise_band.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vpt_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.vpt_render_shard.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.vpt_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vpt_comm_broadcast_gbuffer.argtypes = [C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc, what):
    if rc != 0:
        raise VptError("%s failed (%d): %s" % (what, rc, lib().vpt_last_error().decode()))


# ---- host helpers (no device) ----------------------------------------------------------------
def camera_init(width, height):
    cam = np.zeros(CAMERA_FLOATS, np.float32)
    lib().vpt_camera_init(_p(cam), width, height)
    return cam


def camera_update(cam):
    lib().vpt_camera_update(_p(cam))
    return cam


def camera_from_scene(width, height, position, direction, fov_deg):
    cam = np.zeros(CAMERA_FLOATS, np.float32)
    pos = np.asarray(position, np.float32)
    d = np.asarray(direction, np.float32)
    lib().vpt_camera_from_scene(_p(cam), width, height, _p(pos), _p(d), float(fov_deg))
    return cam


def camera_set_yaw_pitch(cam, yaw, pitch):
    cam = cam.copy()
    cam[15] = np.float32(yaw)
    cam[16] = np.float32(pitch)
    return camera_update(cam)


def perlin_noise_chunks(cx, cy, cz, seed=124):
    out = np.zeros((cx * cy * cz, 32, 32), np.float32)
    lib().vpt_perlin_noise_chunks(cx, cy, cz, seed, _p(out))
    return out


def build_alias_table(weights):
    w = np.ascontiguousarray(weights, np.float32).ravel()
    bins = np.zeros(w.size, ALIAS_DTYPE)
    lib().vpt_build_alias_table(_p(w), w.size, _p(bins))
    return bins


TONEMAP_DTYPE = np.dtype([("manualExposure", "<f4"), ("curve", "<i4"), ("highlightDesaturation", "<f4"), ("whitePoint", "<f4"),
                          ("contrast", "<f4"), ("saturation", "<f4"), ("lift", "<f4"), ("gain", "<f4")])
SKYPARAMS_DTYPE = np.dtype([("timeOfDay", "<f4"), ("sunAxisAngle", "<f4"), ("sunAxisRotate", "<f4"), ("skyBrightness", "<f4")])


def default_tonemapping_params():
    p = np.zeros(1, TONEMAP_DTYPE)
    lib().vpt_default_tonemapping_params(_p(p))
    return p


def load_tonemapping_settings(path, params=None):
    p = default_tonemapping_params() if params is None else params
    rc = lib().vpt_load_tonemapping_settings(path.encode(), _p(p))
    return p, rc


def load_sky_settings(path):
    p = np.zeros(1, SKYPARAMS_DTYPE)
    p["timeOfDay"], p["sunAxisAngle"], p["sunAxisRotate"], p["skyBrightness"] = 0.25, 45.0, 0.0, 1.0
    rc = lib().vpt_load_sky_settings(path.encode(), _p(p))
    return p, rc


def chunk_hash(chunk):
    """FNV-1a-64 of a 32768-byte chunk as 16 hex digits (WorldSceneManager.cpp:240-258)."""
    c = np.ascontiguousarray(chunk, np.uint8).ravel()
    assert c.size == 32768
    out = C.create_string_buffer(17)
    lib().vpt_chunk_hash(_p(c), out)
    return out.value.decode()


def save_world(scene_path, chunk_dir, chunks, ids, cam9, fov=90.0):
    ids = np.ascontiguousarray(ids, np.uint8).ravel()
    cam9 = np.ascontiguousarray(cam9, np.float32)
    return lib().vpt_save_world(scene_path.encode(), chunk_dir.encode(), chunks[0], chunks[1], chunks[2], _p(ids), _p(cam9), C.c_float(fov))


def load_world(scene_path, chunk_dir, chunks, ids):
    """Overwrites `ids` (runtime grid, chunk-major) with the chunk files the scene lists; returns (rc, loaded, failed)."""
    assert ids.dtype == np.uint8 and ids.flags["C_CONTIGUOUS"] and ids.size == chunks[0] * chunks[1] * chunks[2] * 32768
    ok, bad = C.c_int(), C.c_int()
    rc = lib().vpt_load_world(scene_path.encode(), chunk_dir.encode(), chunks[0], chunks[1], chunks[2], _p(ids), C.byref(ok), C.byref(bad))
    return rc, ok.value, bad.value


def sky_state(params, tables):
    """Host part of SkyModel::update: (configs[90], radiances[10], sunDir[3]) for params = (timeOfDay, sunAxisAngle,
    sunAxisRotate, skyBrightness)."""
    pr = np.asarray(params, np.float32)
    tb = np.ascontiguousarray(tables, np.float32)
    cfg = np.zeros(90, np.float32); rad = np.zeros(10, np.float32); sd = np.zeros(3, np.float32)
    lib().vpt_sky_state(_p(pr), _p(tb), _p(cfg), _p(rad), _p(sd))
    return cfg, rad, sd


def default_denoising_params():
    from vpt_scenes import DENOISE_DTYPE
    p = np.zeros(1, DENOISE_DTYPE)
    lib().vpt_default_denoising_params(_p(p))
    return p


def load_denoising_settings(path, params=None):
    p = default_denoising_params() if params is None else params
    rc = lib().vpt_load_denoising_settings(path.encode(), _p(p))
    return p, rc == 0


def load_scene_config(path):
    out9 = np.zeros(9, np.float32)
    fov = C.c_float()
    chunks = np.zeros(3, np.uint32)
    rc = lib().vpt_load_scene_config(path.encode(), _p(out9), C.byref(fov), _p(chunks))
    return dict(position=out9[0:3].copy(), direction=out9[3:6].copy(), up=out9[6:9].copy(), fov=fov.value,
                chunks=tuple(int(v) for v in chunks), loaded=(rc == 0))


def band_rows(height, nranks, rank):
    a, b = C.c_int(), C.c_int()
    lib().vpt_band_rows(height, nranks, rank, C.byref(a), C.byref(b))
    return a.value, b.value


def band_input_halo(params):
    p = np.ascontiguousarray(params)
    return int(lib().vpt_band_input_halo(_p(p)))


def comm_unique_id():
    buf = np.zeros(128, np.uint8)
    _check(lib().vpt_comm_unique_id(_p(buf)), "vpt_comm_unique_id")
    return buf


# ---- device context ----------------------------------------------------------------------------
class Vpt:
    def __init__(self, width, height, device=0):
        self.L = lib()
        self.w, self.h = width, height
        ctx = C.c_void_p()
        _check(self.L.vpt_create(device, width, height, C.byref(ctx)), "vpt_create")
        self.ctx = ctx
        self.chunks = None

    def close(self):
        if getattr(self, "ctx", None):
            self.L.vpt_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(self.L.vpt_sync(self.ctx), "vpt_sync")

    def stream(self):
        return self.L.vpt_stream(self.ctx)

    def set_profiling(self, on):
        self.L.vpt_set_profiling(self.ctx, int(on))

    def set_tables(self, tables):
        t = np.ascontiguousarray(tables, np.uint8)
        assert t.size == 327680
        _check(self.L.vpt_set_tables(self.ctx, _p(t[:65536]), _p(t[65536:196608]), _p(t[196608:])), "vpt_set_tables")

    def set_grid(self, cx, cy, cz, ids):
        ids = np.ascontiguousarray(ids, np.uint8)
        assert ids.size == cx * cy * cz * 32768
        _check(self.L.vpt_set_grid(self.ctx, cx, cy, cz, _p(ids)), "vpt_set_grid")
        self.chunks = (cx, cy, cz)

    def generate_terrain(self, cx, cy, cz, noise):
        noise = np.ascontiguousarray(noise, np.float32)
        assert noise.size == cx * cy * cz * 1024
        _check(self.L.vpt_generate_terrain(self.ctx, cx, cy, cz, _p(noise)), "vpt_generate_terrain")
        self.chunks = (cx, cy, cz)

    def get_grid(self):
        cx, cy, cz = self.chunks
        out = np.zeros(cx * cy * cz * 32768, np.uint8)
        _check(self.L.vpt_get_grid(self.ctx, _p(out), out.size), "vpt_get_grid")
        return out

    def set_voxel(self, x, y, z, block_id):
        _check(self.L.vpt_set_voxel(self.ctx, x, y, z, block_id), "vpt_set_voxel")

    def pick_voxel(self, origin, direction):
        """vpt_pick_voxel -> dict(hasSpaceToCreate, hitSurface, createPos, deletePos, deleteBlockId)."""
        o = np.ascontiguousarray(origin, np.float32); d = np.ascontiguousarray(direction, np.float32)
        out = np.zeros(9, np.int32)
        _check(self.L.vpt_pick_voxel(self.ctx, _p(o), _p(d), _p(out)), "vpt_pick_voxel")
        return dict(hasSpaceToCreate=int(out[0]), hitSurface=int(out[1]), createPos=tuple(int(v) for v in out[2:5]),
                    deletePos=tuple(int(v) for v in out[5:8]), deleteBlockId=int(out[8]))

    def set_materials(self, materials, block_to_material):
        m = np.ascontiguousarray(materials)
        b = np.ascontiguousarray(block_to_material, np.uint16)
        assert m.dtype.itemsize == 48 and b.size == 256
        _check(self.L.vpt_set_materials(self.ctx, _p(m), m.size, _p(b)), "vpt_set_materials")

    def set_textures(self, textures, slots, tex_size):
        """textures: list of (levels-list of (n,n) uint32 RGBA8 arrays); slots: (nMaterials,4) int32 albedo/normal/roughness/metallic
        texture index or -1; tex_size: (nMaterials,2) float32 MaterialParameter::texSize. An empty list removes the textures."""
        widths, levels, texels, slots, tex_size = pack_textures(textures, slots, tex_size)
        _check(self.L.vpt_set_textures(self.ctx, len(textures), _p(widths), _p(levels), _p(texels), slots.shape[0], _p(slots), _p(tex_size)),
               "vpt_set_textures")

    def set_sky(self, sky, sun, sky_alias, sun_alias, sun_dir):
        sky = np.ascontiguousarray(sky, np.float32)
        sun = np.ascontiguousarray(sun, np.float32)
        sa = np.ascontiguousarray(sky_alias, ALIAS_DTYPE)
        su = np.ascontiguousarray(sun_alias, ALIAS_DTYPE)
        sd = np.asarray(sun_dir, np.float32)
        _check(self.L.vpt_set_sky(self.ctx, _p(sky), sky.shape[1], sky.shape[0], _p(sun), sun.shape[1], sun.shape[0], _p(sa), _p(su), _p(sd)),
               "vpt_set_sky")

    def generate_sky(self, params, tables):
        pr = np.asarray(params, np.float32)
        tb = np.ascontiguousarray(tables, np.float32)
        assert pr.size == 4 and tb.size == 2460
        _check(self.L.vpt_generate_sky(self.ctx, _p(pr), _p(tb)), "vpt_generate_sky")

    def read_sky(self):
        w, h, sw, sh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(self.L.vpt_sky_size(self.ctx, C.byref(w), C.byref(h), C.byref(sw), C.byref(sh)), "vpt_sky_size")
        sky = np.zeros((h.value, w.value, 4), np.float32); sun = np.zeros((sh.value, sw.value, 4), np.float32); sd = np.zeros(3, np.float32)
        _check(self.L.vpt_read_sky(self.ctx, _p(sky), _p(sun), _p(sd)), "vpt_read_sky")
        return sky, sun, sd

    def read_async(self, name, out):
        """Pipelined read-back into `out` (pinned host array); completes at read_wait() / sync()."""
        _check(self.L.vpt_read_buffer_async(self.ctx, BUF[name], _p(out), C.c_size_t(out.nbytes)), "vpt_read_buffer_async")

    def read_wait(self):
        _check(self.L.vpt_read_wait(self.ctx), "vpt_read_wait")

    def set_wave_budget(self, max_paths):
        _check(self.L.vpt_set_wave_budget(self.ctx, C.c_size_t(int(max_paths))), "vpt_set_wave_budget")

    def tonemap(self, params):
        """vpt_tonemap on IlluminationOutput: (rgb8[h,w,3] top row first, ldr[h,w,4])."""
        pr = np.ascontiguousarray(params, TONEMAP_DTYPE)
        rgb8 = np.zeros((self.h, self.w, 3), np.uint8); ldr = np.zeros((self.h, self.w, 4), np.float32)
        _check(self.L.vpt_tonemap(self.ctx, _p(pr), _p(rgb8), _p(ldr)), "vpt_tonemap")
        return rgb8, ldr

    def set_trace_params(self, spp=1, total_bounce_limit=3, diffuse_bounce_limit=1, enable_restir=1):
        _check(self.L.vpt_set_trace_params(self.ctx, spp, total_bounce_limit, diffuse_bounce_limit, enable_restir), "vpt_set_trace_params")

    def render(self, cam, prev_cam, iteration_index):
        _check(self.L.vpt_render(self.ctx, _p(cam), _p(prev_cam), iteration_index), "vpt_render")

    def render_shard(self, cam, prev_cam, iteration_index, sample_begin, sample_step):
        _check(self.L.vpt_render_shard(self.ctx, _p(cam), _p(prev_cam), iteration_index, sample_begin, sample_step), "vpt_render_shard")

    def render_range(self, cam, prev_cam, iteration_index, sample_begin, sample_count):
        _check(self.L.vpt_render_range(self.ctx, _p(cam), _p(prev_cam), iteration_index, sample_begin, sample_count), "vpt_render_range")

    def render_shard_local(self, cam, prev_cam, iteration_index, sample_begin, sample_step):
        _check(self.L.vpt_render_shard_local(self.ctx, _p(cam), _p(prev_cam), iteration_index, sample_begin, sample_step), "vpt_render_shard_local")

    def resolve(self):
        _check(self.L.vpt_resolve(self.ctx), "vpt_resolve")

    def begin_external_frame(self):
        _check(self.L.vpt_begin_external_frame(self.ctx), "vpt_begin_external_frame")

    def denoise(self, params, cam, prev_cam, frame_num, iteration_index):
        p = np.ascontiguousarray(params)
        assert p.dtype.itemsize == 68
        _check(self.L.vpt_denoise(self.ctx, _p(p), _p(cam), _p(prev_cam), frame_num, iteration_index), "vpt_denoise")

    def denoise_band(self, params, cam, prev_cam, frame_num, iteration_index, row_begin, row_end):
        p = np.ascontiguousarray(params)
        _check(self.L.vpt_denoise_band(self.ctx, _p(p), _p(cam), _p(prev_cam), frame_num, iteration_index, row_begin, row_end), "vpt_denoise_band")

    def denoise_external(self, params, cam, prev_cam, frame_num, iteration_index, gbuf, out=None):
        """gbuf: dict with Illumination, Depth, NormalRoughness, Material, Albedo host arrays (pinned or not)."""
        p = np.ascontiguousarray(params)
        if out is None:
            out = np.zeros((self.h, self.w, 4), np.float32)
        _check(self.L.vpt_denoise_external(self.ctx, _p(p), _p(cam), _p(prev_cam), frame_num, iteration_index,
                                           _p(gbuf["Illumination"]), _p(gbuf["Depth"]), _p(gbuf["NormalRoughness"]),
                                           _p(gbuf["Material"]), _p(gbuf["Albedo"]), _p(out)), "vpt_denoise_external")
        return out

    def read(self, name, out=None):
        if out is None:
            if name == "PrimaryHits":
                out = np.zeros((self.h, self.w, 4), np.int32)
            elif name in _F1:
                out = np.zeros((self.h, self.w), np.float32)
            else:
                out = np.zeros((self.h, self.w, 4), np.float32)
        _check(self.L.vpt_read_buffer(self.ctx, BUF[name], _p(out), out.nbytes), "vpt_read_buffer(%s)" % name)
        return out

    def write(self, name, arr):
        a = np.ascontiguousarray(arr, np.int32 if name == "PrimaryHits" else np.float32)
        _check(self.L.vpt_write_buffer(self.ctx, BUF[name], _p(a), a.nbytes), "vpt_write_buffer(%s)" % name)

    def device_ptr(self, name):
        return self.L.vpt_device_ptr(self.ctx, BUF[name])

    def read_reservoirs(self, parity):
        out = np.zeros((self.h, self.w), RESERVOIR_DTYPE)
        _check(self.L.vpt_read_reservoirs(self.ctx, parity, _p(out), out.nbytes), "vpt_read_reservoirs")
        return out

    def write_reservoirs(self, parity, arr):
        a = np.ascontiguousarray(arr, RESERVOIR_DTYPE)
        _check(self.L.vpt_write_reservoirs(self.ctx, parity, _p(a), a.nbytes), "vpt_write_reservoirs")

    def counters(self):
        r, s = C.c_uint64(), C.c_uint64()
        _check(self.L.vpt_get_counters(self.ctx, C.byref(r), C.byref(s)), "vpt_get_counters")
        return r.value, s.value

    def lights(self):
        """(VptLightInfo[n], VptAliasBin[n], faceKeys[n/2]): the local emissive lights as the next render will see them."""
        n = int(self.L.vpt_get_lights(self.ctx, None, None, None, 0))
        if n < 0:
            raise VptError("vpt_get_lights failed: %s" % lib().vpt_last_error().decode())
        li = np.zeros(n, LIGHT_DTYPE); al = np.zeros(n, ALIAS_DTYPE); keys = np.zeros(n // 2, np.uint32)
        if n and int(self.L.vpt_get_lights(self.ctx, _p(li), _p(al), _p(keys), n)) != n:
            raise VptError("vpt_get_lights failed: %s" % lib().vpt_last_error().decode())
        return li, al, keys

    def total_rays(self, reset=False):
        r = C.c_uint64()
        _check(self.L.vpt_get_total_rays(self.ctx, C.byref(r), 1 if reset else 0), "vpt_get_total_rays")
        return r.value

    def timings(self):
        t = np.zeros(1, TIMINGS_DTYPE)
        _check(self.L.vpt_get_timings(self.ctx, _p(t)), "vpt_get_timings")
        return {k: (float(t[k][0]) if k.endswith("_ms") else int(t[k][0])) for k in TIMINGS_DTYPE.names}

    def comm_init(self, rank, nranks, unique_id):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        _check(self.L.vpt_comm_init(self.ctx, rank, nranks, _p(uid)), "vpt_comm_init")

    def comm_allreduce_illumination(self):
        _check(self.L.vpt_comm_allreduce_illumination(self.ctx), "vpt_comm_allreduce_illumination")

    def write_device(self, name, device_ptr, nbytes):
        _check(self.L.vpt_write_buffer_device(self.ctx, BUF[name], C.c_void_p(int(device_ptr)), C.c_size_t(int(nbytes))), "vpt_write_buffer_device(%s)" % name)

    def comm_gather_output(self, root=0):
        _check(self.L.vpt_comm_gather_output(self.ctx, root), "vpt_comm_gather_output")

    def comm_broadcast_gbuffer(self, iteration_index):
        _check(self.L.vpt_comm_broadcast_gbuffer(self.ctx, iteration_index), "vpt_comm_broadcast_gbuffer")
