// ORACLE — test infrastructure only (see orc_math.h header). C entry points for the Python tests
// (ctypes), __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
// Mirrors the shape of include/vpt.h so parity tests drive both sides with the same bytes.
#include "orc_denoise.h"
#include "orc_sky.h"
#include <omp.h>
#include <cfloat>
#include <cstdio>
#include <cstring>

using namespace orc;

struct orc_ctx
{
    Scene sc;
    DenoiseState ds;
};

// Buffer names: numeric values are shared with include/vpt.h (VptBufferName)
enum
{
    BUF_Illumination = 0, BUF_IlluminationOutput = 1, BUF_IlluminationPing = 2, BUF_IlluminationPong = 3,
    BUF_NormalRoughness = 4, BUF_Depth = 5, BUF_Material = 6, BUF_Albedo = 7, BUF_HistoryLength = 8,
    BUF_PrevDepth = 9, BUF_PrevMaterial = 10, BUF_PrevIllumination = 11, BUF_PrevFastIllumination = 12,
    BUF_PrevHistoryLength = 13, BUF_PrevNormalRoughness = 14, BUF_GeoNormalThinfilm = 15, BUF_MaterialParameter = 16,
    BUF_PrevMaterialParameter = 17, BUF_PrevGeoNormalThinfilm = 18, BUF_PrevAlbedo = 19,
    BUF_ReservoirCur = 20, BUF_ReservoirPrev = 21, BUF_PrimaryHits = 22
};

extern "C" {

orc_ctx *orc_create(int width, int height)
{
    orc_ctx *c = new orc_ctx();
    c->sc.width = width; c->sc.height = height;
    size_t n = (size_t)width * height;
    c->sc.gb[0].resize(n); c->sc.gb[1].resize(n);
    c->sc.illumination.assign(n, F4(0.0f));
    c->sc.reservoirs.assign(2 * n, emptyReservoir());
    c->sc.primaryHits.assign(4 * n, -1);
    c->ds.resize(n);
    return c;
}
void orc_destroy(orc_ctx *c) { delete c; }
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_max_threads() { return omp_get_max_threads(); }

int orc_set_tables(orc_ctx *c, const uint8_t *sobol, const uint8_t *scrambling, const uint8_t *ranking)
{
    c->sc.tables.sobol.assign(sobol, sobol + 65536);
    c->sc.tables.scrambling.assign(scrambling, scrambling + 131072);
    c->sc.tables.ranking.assign(ranking, ranking + 131072);
    c->sc.tables.ranking.resize(131072 + 256, 0);
    return 0;
}
int orc_set_grid(orc_ctx *c, int cx, int cy, int cz, const uint8_t *ids)
{
    c->sc.grid.cx = cx; c->sc.grid.cy = cy; c->sc.grid.cz = cz;
    c->sc.grid.ids.assign(ids, ids + (size_t)cx * cy * cz * 32768);
    c->sc.havePrevGrid = false; // a new world has no previous state
    c->sc.lightsStale = true;
    return 0;
}
int orc_generate_terrain(orc_ctx *c, int cx, int cy, int cz, const float *noise)
{
    generateTerrain(c->sc.grid, cx, cy, cz, noise);
    c->sc.havePrevGrid = false;
    c->sc.lightsStale = true;
    return 0;
}
int orc_get_grid(orc_ctx *c, uint8_t *out, size_t bytes)
{
    if (bytes != c->sc.grid.ids.size()) return 1;
    std::memcpy(out, c->sc.grid.ids.data(), bytes);
    return 0;
}
int orc_set_voxel(orc_ctx *c, int x, int y, int z, int id)
{
    Grid &g = c->sc.grid;
    if (x < 0 || y < 0 || z < 0 || x >= g.W() || y >= g.H() || z >= g.D()) return 1;
    if (!c->sc.havePrevGrid) { c->sc.prevGrid = g; c->sc.havePrevGrid = true; } // the world the previous render saw
    g.ids[g.index(x, y, z)] = (uint8_t)id;
    c->sc.lightsStale = true;
    return 0;
}
int orc_set_materials(orc_ctx *c, const Material *m, int n, const uint16_t *blockToMaterial)
{
    c->sc.materials.assign(m, m + n);
    std::memcpy(c->sc.blockToMaterial, blockToMaterial, 256 * sizeof(uint16_t));
    c->sc.lightsStale = true;
    return 0;
}
int orc_set_sky(orc_ctx *c, const float *sky, int skyW, int skyH, const float *sun, int sunW, int sunH,
                const AliasBin *skyAlias, const AliasBin *sunAlias, const float *sunDir)
{
    Sky &s = c->sc.sky;
    s.skyW = skyW; s.skyH = skyH; s.sunW = sunW; s.sunH = sunH;
    s.sky.assign((const f4 *)sky, (const f4 *)sky + (size_t)skyW * skyH);
    s.sun.assign((const f4 *)sun, (const f4 *)sun + (size_t)sunW * sunH);
    s.skyAlias.assign(skyAlias, skyAlias + (size_t)skyW * skyH);
    s.sunAlias.assign(sunAlias, sunAlias + (size_t)sunW * sunH);
    s.sunDir = {sunDir[0], sunDir[1], sunDir[2]};
    return 0;
}
// Uncompressed RGBA8 mip chains + per-material texture slots (albedo, normal, roughness, metallic; -1 = none) and texSize.
int orc_set_textures(orc_ctx *c, int nTextures, const int *widths, const int *levels, const uint32_t *texels, int nMaterials,
                     const int32_t *slots4, const float *texSize2)
{
    c->sc.textures.assign((size_t)nTextures, Texture());
    size_t off = 0;
    for (int t = 0; t < nTextures; ++t)
    {
        Texture &tx = c->sc.textures[(size_t)t];
        tx.width = widths[t]; tx.levels = levels[t];
        size_t n = 0;
        for (int l = 0; l < tx.levels; ++l) { tx.levelOffset.push_back(n); n += (size_t)(tx.width >> l) * (tx.width >> l); }
        tx.texels.assign(texels + off, texels + off + n);
        off += n;
    }
    c->sc.matTex.assign((size_t)nMaterials, MaterialTextures());
    for (int m = 0; m < nMaterials; ++m)
    {
        MaterialTextures &mt = c->sc.matTex[(size_t)m];
        mt.albedo = slots4[m * 4]; mt.normal = slots4[m * 4 + 1]; mt.roughness = slots4[m * 4 + 2]; mt.metallic = slots4[m * 4 + 3];
        mt.texSizeX = texSize2[m * 2]; mt.texSizeY = texSize2[m * 2 + 1];
    }
    return 0;
}
// test hook: one trilinear fetch (tex2DLod restatement)
void orc_tex_sample(orc_ctx *c, int tex, float u, float v, float lod, float *out4)
{
    const f4 r = tex2DLod(c->sc.textures[(size_t)tex], u, v, lod);
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
// The block picker, restated from VoxelEngine::performRayTraversal (voxelengine/VoxelEngine.cu:1040-1166). out9 = hasSpaceToCreate,
// hitSurface, createPos xyz, deletePos xyz, deleteBlockId.
int orc_pick_voxel(orc_ctx *c, const float *origin, const float *direction, int32_t *out9)
{
    const Grid &g = c->sc.grid;
    int32_t res[9] = {0, 0, -1, -1, -1, -1, -1, -1, -1};
    const float len = std::sqrt(direction[0] * direction[0] + direction[1] * direction[1] + direction[2] * direction[2]);
    if (!(len <= 1e-8f))
    {
        const int dims[3] = {g.W(), g.H(), g.D()};
        float dir[3], tDelta[3], tNext[3];
        int cell[3], inc[3];
        for (int k = 0; k < 3; ++k)
        {
            dir[k] = direction[k] / len;
            cell[k] = (int)std::floor(origin[k]);
            inc[k] = dir[k] > 0.0f ? 1 : -1;
            const bool flat = std::fabs(dir[k]) < 1e-8f;
            tDelta[k] = flat ? FLT_MAX : 1.0f / std::fabs(dir[k]);
            const float wall = inc[k] > 0 ? (float)(cell[k] + 1) : (float)cell[k];
            tNext[k] = flat ? FLT_MAX : (wall - origin[k]) / dir[k];
        }
        for (int visited = 0; visited < 1000; ++visited)
        {
            bool inside = true;
            for (int k = 0; k < 3; ++k) inside = inside && cell[k] >= 0 && cell[k] < dims[k];
            if (!inside) break;
            const int id = g.at(cell[0], cell[1], cell[2]);
            if (id != 0)
            {
                res[1] = 1; res[5] = cell[0]; res[6] = cell[1]; res[7] = cell[2]; res[8] = id;
                break;
            }
            res[0] = 1; res[2] = cell[0]; res[3] = cell[1]; res[4] = cell[2];
            int k;
            if (tNext[0] < tNext[1]) k = tNext[0] < tNext[2] ? 0 : 2;
            else k = tNext[1] < tNext[2] ? 1 : 2;
            cell[k] += inc[k];
            tNext[k] += tDelta[k];
        }
    }
    std::memcpy(out9, res, sizeof res);
    return 0;
}
int orc_set_trace_params(orc_ctx *c, int spp, int totalBounceLimit, int diffuseBounceLimit, int enableRestir)
{
    c->sc.tp = {spp, totalBounceLimit, diffuseBounceLimit, enableRestir};
    return 0;
}

// OptixRenderer::render for the sample shard {k = sampleBegin, sampleBegin+sampleStep, ...}.
// Leaves the un-normalised radiance SUM in Illumination (w = primary distance on the shard owning k=0).
static int renderShard(orc_ctx *c, const Camera *cam, const Camera *prevCam, int iterationIndex, int sampleBegin, int sampleStep, int ownerSample, int sampleLimit = 0);
int orc_render_shard(orc_ctx *c, const Camera *cam, const Camera *prevCam, int iterationIndex, int sampleBegin, int sampleStep)
{
    return renderShard(c, cam, prevCam, iterationIndex, sampleBegin, sampleStep, 0);
}
// rank-local owner: the shard's first sample owns the G-buffer, the reservoir and the ReSTIR pass
int orc_render_shard_local(orc_ctx *c, const Camera *cam, const Camera *prevCam, int iterationIndex, int sampleBegin, int sampleStep)
{
    return renderShard(c, cam, prevCam, iterationIndex, sampleBegin, sampleStep, sampleBegin);
}
// contiguous form (vpt_render_range): samples sampleBegin .. sampleBegin + sampleCount - 1; 0 samples = an empty shard
int orc_render_range(orc_ctx *c, const Camera *cam, const Camera *prevCam, int iterationIndex, int sampleBegin, int sampleCount)
{
    if (sampleCount == 0) return renderShard(c, cam, prevCam, iterationIndex, c->sc.tp.spp, 1, 0);
    return renderShard(c, cam, prevCam, iterationIndex, sampleBegin, 1, 0, sampleCount);
}
static int renderShard(orc_ctx *c, const Camera *cam, const Camera *prevCam, int iterationIndex, int sampleBegin, int sampleStep, int ownerSample, int sampleLimit)
{
    Scene &sc = c->sc;
    sc.cur ^= 1;
    prepareLightRemap(sc);
    uint64_t rays = 0, steps = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : rays, steps)
    for (int y = 0; y < sc.height; ++y)
        for (int x = 0; x < sc.width; ++x)
        {
            f4 acc;
            renderPixel(sc, *cam, *prevCam, iterationIndex, x, y, sampleBegin, sampleStep, &acc, rays, steps, ownerSample, sampleLimit);
            sc.illumination[(size_t)y * sc.width + x] = acc;
        }
    sc.rayCount = rays; sc.stepCount = steps;
    sc.havePrevGrid = false; // the next frame's previous world is this one unless an edit takes a new snapshot
    return 0;
}
// Divide the (possibly all-reduced) radiance sum by spp.
int orc_resolve(orc_ctx *c)
{
    Scene &sc = c->sc;
    const float spp = (float)sc.tp.spp;
    for (auto &v : sc.illumination) { v.x = v.x / spp; v.y = v.y / spp; v.z = v.z / spp; }
    return 0;
}
int orc_render(orc_ctx *c, const Camera *cam, const Camera *prevCam, int iterationIndex)
{
    orc_render_shard(c, cam, prevCam, iterationIndex, 0, 1);
    return orc_resolve(c);
}
// Advance the G-buffer ping-pong without rendering (denoiser-only use: upload G-buffer, then denoise).
int orc_begin_external_frame(orc_ctx *c) { c->sc.cur ^= 1; return 0; }

int orc_denoise(orc_ctx *c, const DenoisingParams *p, const Camera *cam, const Camera *prevCam, int frameNum, int iterationIndex)
{
    denoiseRun(c->sc, c->ds, *cam, *prevCam, *p, frameNum, iterationIndex);
    return 0;
}
void orc_prepass_rotator(int frameIndex, float *rot4) { prePassRotator(frameIndex, rot4); }
// Single passes, for per-pass parity tests (same argument meaning as the kernels in Denoiser.cu)
int orc_pass_temporal(orc_ctx *c, const DenoisingParams *p, const Camera *cam, const Camera *prevCam) { temporalAccumulation(c->sc, c->ds, *cam, *prevCam, *p); return 0; }
int orc_pass_history_fix(orc_ctx *c, const Camera *cam) { historyFix(c->sc, c->ds, *cam); return 0; }
int orc_pass_history_clamping(orc_ctx *c) { historyClamping(c->sc, c->ds); return 0; }
int orc_pass_firefly(orc_ctx *c, const DenoisingParams *p, const Camera *cam, int parity) { fireflyFilter(c->sc, *cam, parity, 80.0f, 5.0f, 0.8f, 0.02f, p->phiLuminance); return 0; }
int orc_pass_atrous_smem(orc_ctx *c, const DenoisingParams *p, const Camera *cam) { atrousSmem(c->sc, c->ds.prevIllum, c->ds.ping, c->ds.historyLength, *cam, *p); return 0; }
int orc_pass_atrous(orc_ctx *c, const DenoisingParams *p, const Camera *cam, int pingToPong, unsigned frameIndex, unsigned step)
{
    if (pingToPong) atrous(c->sc, c->ds.ping, c->ds.pong, c->ds.historyLength, *cam, frameIndex, step, *p);
    else atrous(c->sc, c->ds.pong, c->ds.ping, c->ds.historyLength, *cam, frameIndex, step, *p);
    return 0;
}

static void *bufferPtr(orc_ctx *c, int name, size_t &bytes)
{
    Scene &sc = c->sc; DenoiseState &ds = c->ds;
    size_t n = (size_t)sc.width * sc.height;
    GBufferSet &g = sc.gb[sc.cur], &pg = sc.gb[sc.cur ^ 1];
    auto F4B = [&](std::vector<f4> &v) { bytes = n * 16; return (void *)v.data(); };
    auto F1B = [&](std::vector<float> &v) { bytes = n * 4; return (void *)v.data(); };
    switch (name)
    {
    case BUF_Illumination: return F4B(sc.illumination);
    case BUF_IlluminationOutput: return F4B(ds.illumOutput);
    case BUF_IlluminationPing: return F4B(ds.ping);
    case BUF_IlluminationPong: return F4B(ds.pong);
    case BUF_NormalRoughness: return F4B(g.normalRoughness);
    case BUF_Depth: return F1B(g.depth);
    case BUF_Material: return F1B(g.material);
    case BUF_Albedo: return F4B(g.albedo);
    case BUF_HistoryLength: return F1B(ds.historyLength);
    case BUF_PrevDepth: return F1B(pg.depth);
    case BUF_PrevMaterial: return F1B(pg.material);
    case BUF_PrevIllumination: return F4B(ds.prevIllum);
    case BUF_PrevFastIllumination: return F4B(ds.prevFastIllum);
    case BUF_PrevHistoryLength: return F1B(ds.prevHistoryLength);
    case BUF_PrevNormalRoughness: return F4B(pg.normalRoughness);
    case BUF_GeoNormalThinfilm: return F4B(g.geoNormalThinfilm);
    case BUF_MaterialParameter: return F4B(g.materialParameter);
    case BUF_PrevMaterialParameter: return F4B(pg.materialParameter);
    case BUF_PrevGeoNormalThinfilm: return F4B(pg.geoNormalThinfilm);
    case BUF_PrevAlbedo: return F4B(pg.albedo);
    case BUF_PrimaryHits: bytes = n * 16; return (void *)sc.primaryHits.data();
    default: bytes = 0; return nullptr;
    }
}
// Reservoir planes are addressed by parity: BUF_ReservoirCur/Prev need the iteration index.
int orc_read_buffer(orc_ctx *c, int name, void *out, size_t bytes)
{
    size_t have; void *p = bufferPtr(c, name, have);
    if (!p || have != bytes) return 1;
    std::memcpy(out, p, bytes);
    return 0;
}
int orc_write_buffer(orc_ctx *c, int name, const void *in, size_t bytes)
{
    size_t have; void *p = bufferPtr(c, name, have);
    if (!p || have != bytes) return 1;
    std::memcpy(p, in, bytes);
    return 0;
}
int orc_read_reservoirs(orc_ctx *c, int parity, void *out, size_t bytes)
{
    size_t n = (size_t)c->sc.width * c->sc.height;
    if (bytes != n * sizeof(Reservoir)) return 1;
    std::memcpy(out, c->sc.reservoirs.data() + (size_t)(parity & 1) * n, bytes);
    return 0;
}
int orc_write_reservoirs(orc_ctx *c, int parity, const void *in, size_t bytes)
{
    size_t n = (size_t)c->sc.width * c->sc.height;
    if (bytes != n * sizeof(Reservoir)) return 1;
    std::memcpy(c->sc.reservoirs.data() + (size_t)(parity & 1) * n, in, bytes);
    return 0;
}
// local emissive lights (orc_lights.h): count, then the list / alias table / face keys as the next render will see them
int orc_light_count(orc_ctx *c)
{
    refreshLights(c->sc);
    return (int)c->sc.lights.lights.size();
}
int orc_get_lights(orc_ctx *c, LightInfo *lights, AliasBin *alias, uint32_t *faceKeys)
{
    const int n = orc_light_count(c);
    if (lights) std::memcpy(lights, c->sc.lights.lights.data(), (size_t)n * sizeof(LightInfo));
    if (alias) std::memcpy(alias, c->sc.lights.alias.data(), (size_t)n * sizeof(AliasBin));
    if (faceKeys) std::memcpy(faceKeys, c->sc.lights.faceKeys.data(), (size_t)(n / 2) * sizeof(uint32_t));
    return n;
}
// fp16 / octahedral helpers, exposed for the known-answer tests against numpy
uint32_t orc_f32_to_f16_bits(float f) { return f32ToF16Bits(f); }
float orc_f16_bits_to_f32(uint32_t h) { return f16BitsToF32(h); }
uint32_t orc_ndir_to_oct(const float *n) { return ndirToOctUnorm32(F3(n[0], n[1], n[2])); }
void orc_oct_to_ndir(uint32_t u, float *out) { const f3 d = octToNdirUnorm32(u); out[0] = d.x; out[1] = d.y; out[2] = d.z; }
void orc_get_counters(orc_ctx *c, uint64_t *rays, uint64_t *steps) { *rays = c->sc.rayCount; *steps = c->sc.stepCount; }

// ---- stand-alone helpers
void orc_camera_init(Camera *cam, int w, int h) { cameraInit(*cam, w, h); }
void orc_camera_update(Camera *cam) { cameraUpdate(*cam); }
// mainOffline.cpp:227-247: camera from a scene config (position, direction, fov degrees)
void orc_camera_from_scene(Camera *cam, int w, int h, const float *pos, const float *dirIn, float fovDeg)
{
    cameraInit(*cam, w, h);
    cam->pos = {pos[0], pos[1], pos[2]};
    f3 dir = normalize(normalize(F3(dirIn[0], dirIn[1], dirIn[2]))); // SceneConfig normalises, then mainOffline again
    f2 yp = dirToYawPitch(dir);
    cam->yaw = yp.x; cam->pitch = yp.y;
    float fovX = fovDeg * kPiOver180;
    float fovY = fovX * (cam->resolution.y / cam->resolution.x);
    cam->tanHalfFov = {tanf(fovX * 0.5f), tanf(fovY * 0.5f)};
    cameraUpdate(*cam);
}
void orc_uv_to_world_direction(const Camera *cam, float u, float v, float *out)
{
    f3 d = uvToWorldDirection(*cam, {u, v});
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
}
void orc_world_direction_to_uv(const Camera *cam, const float *d, float *out)
{
    f2 uv = worldDirectionToUV(*cam, {d[0], d[1], d[2]});
    out[0] = uv.x; out[1] = uv.y;
}
void orc_perlin_noise_chunks(int cx, int cy, int cz, unsigned seed, float *out)
{
    Perlin p(seed);
    for (int c = 0; c < cx * cy * cz; ++c)
        chunkNoiseMap(p, c % cx, (c / cx) % cz, cx * 32, out + (size_t)c * 1024);
}
float orc_perlin_noise(unsigned seed, int octaves, float x, float y)
{
    Perlin p(seed);
    return p.octave2D_01(x, y, octaves);
}
// SkyModel::update restatement (orc_sky.h). params: timeOfDay, sunAxisAngle, sunAxisRotate, skyBrightness.
void orc_generate_sky(const float *params, const float *tables, int skyW, int skyH, int sunW, int sunH, float *sky, float *sun, float *skyPdf,
                      float *sunPdf, float *sunDir)
{
    f3 sd;
    generateSky(params, tables, skyW, skyH, sunW, sunH, (f4 *)sky, (f4 *)sun, skyPdf, sunPdf, &sd);
    sunDir[0] = sd.x; sunDir[1] = sd.y; sunDir[2] = sd.z;
}
void orc_sky_state(const float *params, const float *tables, float *configs90, float *radiances10, float *sunDir)
{
    const f3 sd = skySunDir(params[0], params[1], params[2]);
    const SkyState st = skyUpdateState(skyTablesFrom(tables), sd);
    std::memcpy(configs90, st.configs, sizeof st.configs); std::memcpy(radiances10, st.radiances, sizeof st.radiances);
    sunDir[0] = sd.x; sunDir[1] = sd.y; sunDir[2] = sd.z;
}
void orc_tonemap(const float *hdrRGBA, int W, int H, const ToneMappingParams *p, uint8_t *rgb8, float *ldrRGBA)
{
    tonemap((const f4 *)hdrRGBA, W, H, *p, rgb8, (f4 *)ldrRGBA);
}
void orc_build_alias_table(const float *weights, unsigned n, AliasBin *bins) { buildAliasTable(weights, n, bins); }
float orc_rand(orc_ctx *c, int px, int py, int sampleIndex, int dim) { return blueNoiseRand(c->sc.tables, px, py, sampleIndex, dim); }

// out: hit,x,y,z,face,id,steps ; t
int orc_dda(orc_ctx *c, const float *o, const float *d, float tmin, float tmax, int *out, float *t)
{
    Hit h = ddaTrace(c->sc.grid, {o[0], o[1], o[2]}, {d[0], d[1], d[2]}, tmin, tmax);
    out[0] = h.hit; out[1] = h.x; out[2] = h.y; out[3] = h.z; out[4] = h.face; out[5] = h.id; out[6] = h.steps;
    *t = h.t;
    return 0;
}
// Disney BSDF probes (unit tests: furnace / reciprocity checks)
void orc_disney_evaluate(const float *n, const float *wi, const float *wo, const float *albedo, int metallic, float roughness, float *bsdfOut, float *pdfOut)
{
    f3 bsdf; float pdf;
    disneyEvaluate({n[0], n[1], n[2]}, {n[0], n[1], n[2]}, {wi[0], wi[1], wi[2]}, {wo[0], wo[1], wo[2]},
                   {albedo[0], albedo[1], albedo[2]}, metallic != 0, 0.0f, roughness, bsdf, pdf);
    bsdfOut[0] = bsdf.x; bsdfOut[1] = bsdf.y; bsdfOut[2] = bsdf.z; *pdfOut = pdf;
}
void orc_disney_sample(const float *u4, const float *n, const float *wo, const float *albedo, int metallic, float translucency, float roughness,
                       float *wiOut, float *bsdfOverPdfOut, float *pdfOut)
{
    f3 wi, bop; float pdf; bool tr;
    disneySample({u4[0], u4[1], u4[2], u4[3]}, {n[0], n[1], n[2]}, {n[0], n[1], n[2]}, {wo[0], wo[1], wo[2]},
                 {albedo[0], albedo[1], albedo[2]}, metallic != 0, translucency, roughness, wi, bop, pdf, tr);
    wiOut[0] = wi.x; wiOut[1] = wi.y; wiOut[2] = wi.z;
    bsdfOverPdfOut[0] = bop.x; bsdfOverPdfOut[1] = bop.y; bsdfOverPdfOut[2] = bop.z; *pdfOut = pdf;
}

} // extern "C"
