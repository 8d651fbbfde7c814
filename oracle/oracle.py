"""ORACLE — test infrastructure only.

ctypes binding of oracle/liboracle.so (the CPU restatement of the reference's traversal, shading and
denoiser code; see the headers of oracle/orc_*.h for the reference file:line each function follows).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg import this
module; the product path (libvpt.so) never does.

The object model mirrors the product binding (python/vpt.py) on purpose so parity tests drive both with
the same bytes: Oracle(width, height).set_tables/.set_grid/.set_materials/.set_sky/.render/.denoise/.read.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# buffer names shared with include/vpt.h (VptBufferName)
BUF = dict(
    Illumination=0, IlluminationOutput=1, IlluminationPing=2, IlluminationPong=3, NormalRoughness=4, Depth=5,
    Material=6, Albedo=7, HistoryLength=8, PrevDepth=9, PrevMaterial=10, PrevIllumination=11,
    PrevFastIllumination=12, PrevHistoryLength=13, PrevNormalRoughness=14, GeoNormalThinfilm=15,
    MaterialParameter=16, PrevMaterialParameter=17, PrevGeoNormalThinfilm=18, PrevAlbedo=19, PrimaryHits=22)
_F1 = {"Depth", "Material", "HistoryLength", "PrevDepth", "PrevMaterial", "PrevHistoryLength"}

RESERVOIR_DTYPE = np.dtype([("lightData", "<u4"), ("uvData", "<u4"), ("weightSum", "<f4"), ("targetPdf", "<f4"), ("M", "<f4")])
ALIAS_DTYPE = np.dtype([("q", "<f4"), ("p", "<f4"), ("alias", "<i4")])
MATERIAL_DTYPE = np.dtype([("albedo", "<f4", 3), ("roughness", "<f4"), ("translucency", "<f4"), ("uvScale", "<f4"),
                           ("metallic", "<i4"), ("materialId", "<i4"), ("useWorldGridUV", "<i4"), ("isEmissive", "<i4"),
                           ("isThinfilm", "<i4"), ("pad", "<i4")])
DENOISE_DTYPE = np.dtype([("enableHitDistanceReconstruction", "<i4"), ("enablePrePass", "<i4"), ("enableTemporalAccumulation", "<i4"),
                          ("enableHistoryFix", "<i4"), ("enableHistoryClamping", "<i4"), ("enableSpatialFiltering", "<i4"),
                          ("enableFireflyFilter", "<i4"), ("maxAccumulatedFrameNum", "<f4"), ("maxFastAccumulatedFrameNum", "<f4"),
                          ("phiLuminance", "<f4"), ("lobeAngleFraction", "<f4"), ("roughnessFraction", "<f4"), ("depthThreshold", "<f4"),
                          ("atrousIterationNum", "<i4"), ("disocclusionThreshold", "<f4"), ("disocclusionThresholdAlternate", "<f4"),
                          ("denoisingRange", "<f4")])
assert RESERVOIR_DTYPE.itemsize == 20 and ALIAS_DTYPE.itemsize == 12 and MATERIAL_DTYPE.itemsize == 48 and DENOISE_DTYPE.itemsize == 68
CAMERA_FLOATS = 53  # 212-byte POD (Camera.h:6-28)


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.startswith("orc_")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/voxelengine"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_perlin_noise.restype = C.c_float
        L.orc_perlin_noise.argtypes = [C.c_uint, C.c_int, C.c_float, C.c_float]
        L.orc_rand.restype = C.c_float
        L.orc_rand.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_uv_to_world_direction.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.orc_camera_from_scene.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float]
        L.orc_dda.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_disney_evaluate.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_disney_sample.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("orc_read_buffer", "orc_write_buffer"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        for name in ("orc_read_reservoirs", "orc_write_reservoirs"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_get_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def camera_init(width, height):
    cam = np.zeros(CAMERA_FLOATS, np.float32)
    lib().orc_camera_init(_p(cam), width, height)
    return cam


def camera_update(cam):
    lib().orc_camera_update(_p(cam))
    return cam


def camera_set_yaw_pitch(cam, yaw, pitch):
    cam = cam.copy()
    cam[15] = np.float32(yaw)
    cam[16] = np.float32(pitch)
    return camera_update(cam)


def camera_from_scene(width, height, pos, direction, fov_deg):
    cam = np.zeros(CAMERA_FLOATS, np.float32)
    pos = np.asarray(pos, np.float32)
    direction = np.asarray(direction, np.float32)
    lib().orc_camera_from_scene(_p(cam), width, height, _p(pos), _p(direction), float(fov_deg))
    return cam


def uv_to_world_direction(cam, u, v):
    out = np.zeros(3, np.float32)
    lib().orc_uv_to_world_direction(_p(cam), float(u), float(v), _p(out))
    return out


def world_direction_to_uv(cam, d):
    out = np.zeros(2, np.float32)
    d = np.asarray(d, np.float32)
    lib().orc_world_direction_to_uv(_p(cam), _p(d), _p(out))
    return out


def perlin_noise_chunks(cx, cy, cz, seed=124):
    out = np.zeros((cx * cy * cz, 32, 32), np.float32)
    lib().orc_perlin_noise_chunks(cx, cy, cz, seed, _p(out))
    return out


def build_alias_table(weights):
    w = np.ascontiguousarray(weights, np.float32).ravel()
    bins = np.zeros(w.size, ALIAS_DTYPE)
    lib().orc_build_alias_table(_p(w), w.size, _p(bins))
    return bins


def generate_sky(params, tables, sky_w=1024, sky_h=512, sun_w=32, sun_h=32):
    """SkyModel::update restatement: params = (timeOfDay, sunAxisAngle, sunAxisRotate, skyBrightness), tables = the 2460
    floats of data/sky_tables.bin. Returns sky[h,w,4], sun[h,w,4], skyPdf, sunPdf, sunDir."""
    pr = np.asarray(params, np.float32)
    tb = np.ascontiguousarray(tables, np.float32)
    sky = np.zeros((sky_h, sky_w, 4), np.float32); sun = np.zeros((sun_h, sun_w, 4), np.float32)
    sky_pdf = np.zeros(sky_h * sky_w, np.float32); sun_pdf = np.zeros(sun_h * sun_w, np.float32)
    sd = np.zeros(3, np.float32)
    lib().orc_generate_sky(_p(pr), _p(tb), sky_w, sky_h, sun_w, sun_h, _p(sky), _p(sun), _p(sky_pdf), _p(sun_pdf), _p(sd))
    return sky, sun, sky_pdf, sun_pdf, sd


def sky_state(params, tables):
    pr = np.asarray(params, np.float32)
    tb = np.ascontiguousarray(tables, np.float32)
    cfg = np.zeros(90, np.float32); rad = np.zeros(10, np.float32); sd = np.zeros(3, np.float32)
    lib().orc_sky_state(_p(pr), _p(tb), _p(cfg), _p(rad), _p(sd))
    return cfg, rad, sd


TONEMAP_DTYPE = np.dtype([("manualExposure", "<f4"), ("curve", "<i4"), ("highlightDesaturation", "<f4"), ("whitePoint", "<f4"),
                          ("contrast", "<f4"), ("saturation", "<f4"), ("lift", "<f4"), ("gain", "<f4")])


def tonemap(hdr_rgba, params):
    """FilmicToneMapping (manual exposure) + PNG conversion restatement: returns (rgb8[h,w,3] top row first, ldr[h,w,4])."""
    hdr = np.ascontiguousarray(hdr_rgba, np.float32)
    h, w = hdr.shape[:2]
    pr = np.ascontiguousarray(params, TONEMAP_DTYPE)
    rgb8 = np.zeros((h, w, 3), np.uint8); ldr = np.zeros((h, w, 4), np.float32)
    lib().orc_tonemap(_p(hdr), w, h, _p(pr), _p(rgb8), _p(ldr))
    return rgb8, ldr


LIGHT_DTYPE = np.dtype([("center", "<f4", 3), ("scalars", "<u4"), ("radiance", "<u4", 2), ("direction1", "<u4"), ("direction2", "<u4")])


def f32_to_f16_bits(f):
    lib().orc_f32_to_f16_bits.argtypes = [C.c_float]; lib().orc_f32_to_f16_bits.restype = C.c_uint32
    return int(lib().orc_f32_to_f16_bits(float(f)))


def f16_bits_to_f32(h):
    lib().orc_f16_bits_to_f32.argtypes = [C.c_uint32]; lib().orc_f16_bits_to_f32.restype = C.c_float
    return float(lib().orc_f16_bits_to_f32(int(h)))


def set_threads(n):
    lib().orc_set_threads(int(n))


def max_threads():
    return int(lib().orc_max_threads())


def prepass_rotator(frame_index):
    out = np.zeros(4, np.float32)
    lib().orc_prepass_rotator(int(frame_index), _p(out))
    return out


class Oracle:
    def __init__(self, width, height):
        self.L = lib()
        self.w, self.h = width, height
        self.ctx = C.c_void_p(self.L.orc_create(width, height))

    def close(self):
        if self.ctx:
            self.L.orc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tables(self, tables):
        t = np.ascontiguousarray(tables, np.uint8)
        assert t.size == 327680
        self.L.orc_set_tables(self.ctx, _p(t[:65536]), _p(t[65536:196608]), _p(t[196608:]))

    def set_grid(self, cx, cy, cz, ids):
        ids = np.ascontiguousarray(ids, np.uint8)
        assert ids.size == cx * cy * cz * 32768
        self.L.orc_set_grid(self.ctx, cx, cy, cz, _p(ids))
        self.chunks = (cx, cy, cz)

    def generate_terrain(self, cx, cy, cz, noise):
        noise = np.ascontiguousarray(noise, np.float32)
        self.L.orc_generate_terrain(self.ctx, cx, cy, cz, _p(noise))
        self.chunks = (cx, cy, cz)

    def get_grid(self):
        cx, cy, cz = self.chunks
        out = np.zeros(cx * cy * cz * 32768, np.uint8)
        assert self.L.orc_get_grid(self.ctx, _p(out), out.size) == 0
        return out

    def set_voxel(self, x, y, z, block_id):
        return self.L.orc_set_voxel(self.ctx, x, y, z, block_id)

    def pick_voxel(self, origin, direction):
        o = np.ascontiguousarray(origin, np.float32); d = np.ascontiguousarray(direction, np.float32)
        out = np.zeros(9, np.int32)
        self.L.orc_pick_voxel(self.ctx, _p(o), _p(d), _p(out))
        return dict(hasSpaceToCreate=int(out[0]), hitSurface=int(out[1]), createPos=tuple(int(v) for v in out[2:5]),
                    deletePos=tuple(int(v) for v in out[5:8]), deleteBlockId=int(out[8]))

    def set_materials(self, materials, block_to_material):
        m = np.ascontiguousarray(materials, MATERIAL_DTYPE)
        b = np.ascontiguousarray(block_to_material, np.uint16)
        assert b.size == 256
        self.L.orc_set_materials(self.ctx, _p(m), m.size, _p(b))

    def set_textures(self, textures, slots, tex_size):
        widths = np.array([t[0].shape[0] for t in textures], np.int32)
        levels = np.array([len(t) for t in textures], np.int32)
        texels = np.ascontiguousarray(np.concatenate([np.ascontiguousarray(l, np.uint32).ravel() for t in textures for l in t])
                                      if textures else np.zeros(1, np.uint32))
        slots = np.ascontiguousarray(slots, np.int32).reshape(-1, 4)
        tex_size = np.ascontiguousarray(tex_size, np.float32).reshape(-1, 2)
        self.L.orc_set_textures(self.ctx, len(textures), _p(widths), _p(levels), _p(texels), slots.shape[0], _p(slots), _p(tex_size))

    def tex_sample(self, tex, u, v, lod):
        out = np.zeros(4, np.float32)
        self.L.orc_tex_sample(self.ctx, int(tex), C.c_float(u), C.c_float(v), C.c_float(lod), _p(out))
        return out

    def set_sky(self, sky, sun, sky_alias, sun_alias, sun_dir):
        sky = np.ascontiguousarray(sky, np.float32)
        sun = np.ascontiguousarray(sun, np.float32)
        sd = np.asarray(sun_dir, np.float32)
        sa = np.ascontiguousarray(sky_alias, ALIAS_DTYPE)
        su = np.ascontiguousarray(sun_alias, ALIAS_DTYPE)
        self.L.orc_set_sky(self.ctx, _p(sky), sky.shape[1], sky.shape[0], _p(sun), sun.shape[1], sun.shape[0], _p(sa), _p(su), _p(sd))

    def set_trace_params(self, spp=1, total_bounce_limit=3, diffuse_bounce_limit=1, enable_restir=1):
        self.L.orc_set_trace_params(self.ctx, spp, total_bounce_limit, diffuse_bounce_limit, enable_restir)

    def render(self, cam, prev_cam, iteration_index):
        self.L.orc_render(self.ctx, _p(cam), _p(prev_cam), iteration_index)

    def render_shard(self, cam, prev_cam, iteration_index, sample_begin, sample_step):
        self.L.orc_render_shard(self.ctx, _p(cam), _p(prev_cam), iteration_index, sample_begin, sample_step)

    def render_range(self, cam, prev_cam, iteration_index, sample_begin, sample_count):
        assert self.L.orc_render_range(self.ctx, _p(cam), _p(prev_cam), iteration_index, sample_begin, sample_count) == 0

    def render_shard_local(self, cam, prev_cam, iteration_index, sample_begin, sample_step):
        self.L.orc_render_shard_local(self.ctx, _p(cam), _p(prev_cam), iteration_index, sample_begin, sample_step)

    def resolve(self):
        self.L.orc_resolve(self.ctx)

    def begin_external_frame(self):
        self.L.orc_begin_external_frame(self.ctx)

    def denoise(self, params, cam, prev_cam, frame_num, iteration_index):
        p = np.ascontiguousarray(params, DENOISE_DTYPE)
        self.L.orc_denoise(self.ctx, _p(p), _p(cam), _p(prev_cam), frame_num, iteration_index)

    def read(self, name):
        n = self.w * self.h
        if name == "PrimaryHits":
            out = np.zeros((self.h, self.w, 4), np.int32)
        elif name in _F1:
            out = np.zeros((self.h, self.w), np.float32)
        else:
            out = np.zeros((self.h, self.w, 4), np.float32)
        rc = self.L.orc_read_buffer(self.ctx, BUF[name], _p(out), out.nbytes)
        assert rc == 0, name
        return out

    def write(self, name, arr):
        dt = np.int32 if name == "PrimaryHits" else np.float32
        a = np.ascontiguousarray(arr, dt)
        rc = self.L.orc_write_buffer(self.ctx, BUF[name], _p(a), a.nbytes)
        assert rc == 0, name

    def read_reservoirs(self, parity):
        out = np.zeros((self.h, self.w), RESERVOIR_DTYPE)
        assert self.L.orc_read_reservoirs(self.ctx, parity, _p(out), out.nbytes) == 0
        return out

    def write_reservoirs(self, parity, arr):
        a = np.ascontiguousarray(arr, RESERVOIR_DTYPE)
        assert self.L.orc_write_reservoirs(self.ctx, parity, _p(a), a.nbytes) == 0

    def counters(self):
        r, s = C.c_uint64(), C.c_uint64()
        self.L.orc_get_counters(self.ctx, C.byref(r), C.byref(s))
        return r.value, s.value

    def lights(self):
        """(LightInfo[n], AliasBin[n], faceKeys[n/2]) of the local emissive lights (exposed faces of emissive voxels)."""
        n = int(self.L.orc_light_count(self.ctx))
        li = np.zeros(n, LIGHT_DTYPE); al = np.zeros(n, ALIAS_DTYPE); keys = np.zeros(n // 2, np.uint32)
        if n:
            self.L.orc_get_lights(self.ctx, _p(li), _p(al), _p(keys))
        return li, al, keys

    def dda(self, origin, direction, tmin=0.0, tmax=1.0e27):
        o = np.asarray(origin, np.float32)
        d = np.asarray(direction, np.float32)
        out = np.zeros(7, np.int32)
        t = C.c_float()
        self.L.orc_dda(self.ctx, _p(o), _p(d), float(tmin), float(tmax), _p(out), C.byref(t))
        return dict(hit=int(out[0]), voxel=(int(out[1]), int(out[2]), int(out[3])), face=int(out[4]), id=int(out[5]),
                    steps=int(out[6]), t=t.value)
