// libvpt.so — context, device memory and the C ABI (include/vpt.h) over the sm_100a kernels.
//
// Replaces the reference's singleton stack on this path: OfflineBackend (renderer/core/OfflineBackend.cpp:26-131),
// BufferManager (renderer/core/BufferManager.cpp:107-241: 34 cudaArray surfaces -> dense linear planes in HBM),
// OptixRenderer::render (renderer/core/OptixRenderer.cpp:411-485) and Denoiser::run (renderer/denoising/Denoiser.cu:24-408).
// Memory plan per context (fp32, W*H pixels): 2 G-buffer sets x (4+4+16+16+16+16) B, 8 float4 + 2 float denoiser
// planes, 2 reservoir planes x 20 B, int4 primary hits  ->  ~356 B/px (0.74 GB at 1080p, 2.9 GB at 4K): trivially
// resident in 180 GB HBM3e. The Prev* G-buffer copies of the reference (Denoiser.cu:394-407, OptixRenderer.cpp:476-478)
// are pointer ping-pong here: the "current" set flips at every render.
#include "vpt_kernels.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace vpt;

static thread_local std::string g_lastError;
static int fail(int code, const std::string &msg) { g_lastError = msg; return code; }
#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(VPT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct GSet
{
    float *depth = nullptr, *material = nullptr;
    float4 *normalRoughness = nullptr, *geoNormalThinfilm = nullptr, *materialParameter = nullptr, *albedo = nullptr;
    GBufferPtrs ptrs() const { return {depth, material, normalRoughness, geoNormalThinfilm, materialParameter, albedo}; }
};

enum { EV_TRACE0, EV_TRACE1, EV_RESOLVE1, EV_DN0, EV_FIREFLY, EV_SKY, EV_TEMPORAL, EV_HFIX, EV_HCLAMP, EV_ASMEM, EV_ATROUS, EV_COMP, EV_COUNT };

struct vpt_ctx
{
    int device = 0, width = 0, height = 0, smCount = 0;
    size_t smemOptIn = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;       // pipelined read-backs (vpt_read_buffer_async)
    cudaEvent_t copyReady = nullptr, copyDone = nullptr;
    bool copyPending = false, copyPendingTraceWritten = false;
    // tables
    uint8_t *sobol = nullptr, *scrambling = nullptr, *ranking = nullptr;
    // grid
    int cx = 0, cy = 0, cz = 0;
    uint8_t *idsChunk = nullptr, *idsLinear = nullptr;
    uint32_t *occ = nullptr;
    int *upHDev = nullptr; int upH = 0; // highest solid y + 1 (GridView::upH)
    // the world as the PREVIOUS render saw it (the reference's prevTopObject, closesthit.cu:736-755): a snapshot of the masks
    // taken by the first edit after a render, walked by the bias-correction rays of the temporal ReSTIR pass
    uint32_t *occPrev = nullptr; size_t occPrevWords = 0; int upHPrev = 0; bool prevSnapshot = false;
    // materials / sky
    VptMaterial *materials = nullptr; int nMaterials = 0;
    uint16_t *blockToMaterial = nullptr;
    VptPickResult *pickDev = nullptr;
    uint32_t *texels = nullptr; int4 *texDescs = nullptr, *matTexSlots = nullptr; float *matTexMip0Size = nullptr; int nTextures = 0;
    float4 *sky = nullptr, *sun = nullptr;      // four views into skyArena
    VptAliasBin *skyAlias = nullptr, *sunAlias = nullptr;
    void *skyArena = nullptr;
    size_t skyArenaBytes = 0;
    const void *l2Base = nullptr; size_t l2Bytes = 0; // what the stream's L2 access-policy window covers now
    int skyW = 0, skyH = 0, sunW = 0, sunH = 0;
    float sunDir[3] = {0, 1, 0};
    // trace params
    int spp = 1, totalBounceLimit = 3, diffuseBounceLimit = 1, enableRestir = 1;
    // buffers
    GSet gb[2];
    int cur = 1;
    float4 *illumination = nullptr, *illumOutput = nullptr, *ping = nullptr, *pong = nullptr, *prevIllum = nullptr, *prevFastIllum = nullptr;
    float *historyLength = nullptr, *prevHistoryLength = nullptr;
    VptReservoir *reservoirs = nullptr; // 2 planes
    int4 *primaryHits = nullptr;
    unsigned long long *counters = nullptr;
    FireflyPatch *patches = nullptr; int *patchCount = nullptr; int maxPatches = 0;
    float4 *dnG = nullptr; uint32_t *dnMQ = nullptr; // packed denoiser G-buffer (vpt_denoise.cu)
    unsigned *dnCounters = nullptr; int4 *fireflyList = nullptr; int *fixList = nullptr;
    // wavefront workspace (vpt_wave.cu), sized at the first render for (pixel slots, samples per wave)
    WaveWorkspace wave;
    TraceProfile traceProf;
    TraceStreams traceStreams;
    bool overlapParts = false; // two-stream part overlap: measured, no gain (see vpt_wave.cu launchTrace)
    bool texSpecular = false; // a roughness map exists: any texel may be specular
    bool anySpecular = false; // a non-diffuse, non-emissive material exists: paths may continue past their first hit
    int countSteps = 1;
    size_t waveBudget = (size_t)16u << 20;
    // staging for vpt_denoise_external
    void *pinned = nullptr; size_t pinnedBytes = 0;
    uint8_t *rgb8 = nullptr; // vpt_tonemap's 8-bit plane
    // local emissive lights (host/vpt_lights.cpp): the list is rebuilt by the first render after the grid or the materials changed
    std::vector<VptMaterial> hostMaterials; uint16_t hostB2m[256] = {};
    LightList lights; std::vector<int> prevLightToCur; int prevNumLights = 0;
    std::vector<uint32_t> renderedKeys; // face keys of the list the LAST RENDER used: what the stored reservoirs' light ids refer to
    bool lightsStale = true, lightsStateDirty = false;
    VptLightInfo *dLights = nullptr; VptAliasBin *dLightAlias = nullptr; uint32_t *dFaceKeys = nullptr; int *dPrevToCur = nullptr;
    size_t dLightCap = 0, dRemapCap = 0;
    bool dnGather = false;   // VPT_DN_GATHER=1: the per-thread gather kernels instead of the shared-memory tile kernels (A/B runs)
    // profiling
    bool profiling = true;
    cudaEvent_t ev[EV_COUNT] = {};
    bool haveTrace = false, haveDenoise = false, ranFirefly = false, ranTemporal = false, ranFix = false, ranClamp = false, ranSpatial = false, ranResolve = false;
    int atrousPasses = 0, launchesRender = 0, launchesDenoise = 0;
    float atrousMs = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> atrousEv;
    // multi-GPU
    void *ncclComm = nullptr; int rank = 0, nranks = 1;
    size_t npix() const { return (size_t)width * height; }
};

static int createImpl(vpt_ctx *c);
static void destroyComm(vpt_ctx *c);
// a pipelined read-back (vpt_read_buffer_async) still in flight: the stream waits for it on the device before anything overwrites a
// plane it may be reading
static cudaError_t waitPendingCopy(vpt_ctx *c)
{
    if (!c->copyPending) return cudaSuccess;
    c->copyPending = false; c->copyPendingTraceWritten = false;
    return cudaStreamWaitEvent(c->stream, c->copyDone, 0);
}

extern "C" {

const char *vpt_last_error(void) { return g_lastError.c_str(); }

int vpt_create(int device, int width, int height, vpt_ctx **out)
{
    if (!out || width <= 0 || height <= 0) return fail(VPT_ERR_ARG, "vpt_create: bad arguments");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(VPT_ERR_CUDA, std::string("vpt_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(VPT_ERR_ARG, "vpt_create: device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(VPT_ERR_CUDA, "vpt_create: device is not sm_100-class; libvpt is built for sm_100a only");
    vpt_ctx *c = new vpt_ctx();
    c->device = device; c->width = width; c->height = height; c->smCount = prop.multiProcessorCount;
    c->smemOptIn = prop.sharedMemPerBlockOptin;
    { const char *e = std::getenv("VPT_DN_GATHER"); c->dnGather = e && e[0] == '1'; }
    const int rc = createImpl(c);
    if (rc != VPT_OK) { const std::string keep = g_lastError; vpt_destroy(c); g_lastError = keep; return rc; } // nothing leaks on a partial create
    *out = c;
    return VPT_OK;
}

} // extern "C"

static int createImpl(vpt_ctx *c)
{
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->copyReady, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->copyDone, cudaEventDisableTiming));
    const size_t n = c->npix();
    auto alloc = [&](void **p, size_t bytes) -> cudaError_t {
        cudaError_t r = cudaMalloc(p, bytes);
        if (r == cudaSuccess) r = cudaMemsetAsync(*p, 0, bytes, c->stream);
        return r;
    };
    for (int s = 0; s < 2; ++s)
    {
        CU(alloc((void **)&c->gb[s].depth, n * 4)); CU(alloc((void **)&c->gb[s].material, n * 4));
        CU(alloc((void **)&c->gb[s].normalRoughness, n * 16)); CU(alloc((void **)&c->gb[s].geoNormalThinfilm, n * 16));
        CU(alloc((void **)&c->gb[s].materialParameter, n * 16)); CU(alloc((void **)&c->gb[s].albedo, n * 16));
    }
    CU(alloc((void **)&c->illumination, n * 16)); CU(alloc((void **)&c->illumOutput, n * 16));
    CU(alloc((void **)&c->ping, n * 16)); CU(alloc((void **)&c->pong, n * 16));
    CU(alloc((void **)&c->prevIllum, n * 16)); CU(alloc((void **)&c->prevFastIllum, n * 16));
    CU(alloc((void **)&c->historyLength, n * 4)); CU(alloc((void **)&c->prevHistoryLength, n * 4));
    CU(alloc((void **)&c->reservoirs, 2 * n * sizeof(VptReservoir)));
    CU(alloc((void **)&c->primaryHits, n * sizeof(int4)));
    CU(alloc((void **)&c->counters, 4 * sizeof(unsigned long long)));
    c->maxPatches = (int)(n / 4 + 64);
    CU(alloc((void **)&c->patches, (size_t)c->maxPatches * sizeof(FireflyPatch)));
    CU(alloc((void **)&c->patchCount, sizeof(int)));
    CU(alloc((void **)&c->dnG, n * 16)); CU(alloc((void **)&c->dnMQ, n * 4)); CU(alloc((void **)&c->dnCounters, 4 * sizeof(unsigned)));
    CU(alloc((void **)&c->fireflyList, (size_t)c->maxPatches * sizeof(int4))); CU(alloc((void **)&c->fixList, n * sizeof(int)));
    CU(alloc((void **)&c->sobol, 65536)); CU(alloc((void **)&c->scrambling, 131072)); CU(alloc((void **)&c->ranking, 131072 + 256));
    CU(alloc((void **)&c->blockToMaterial, 256 * sizeof(uint16_t)));
    for (int i = 0; i < EV_COUNT; ++i) CU(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i <= TraceProfile::kMax; ++i) CU(cudaEventCreate(&c->traceProf.ev[i]));
    for (int i = 0; i < 2; ++i)
    {
        CU(cudaStreamCreateWithFlags(&c->traceStreams.part[i], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->traceStreams.join[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->traceStreams.fork, cudaEventDisableTiming));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}

extern "C" {

void vpt_destroy(vpt_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copyStream) cudaStreamSynchronize(c->copyStream);
    for (int i = 0; i < 2; ++i) if (c->traceStreams.part[i]) cudaStreamSynchronize(c->traceStreams.part[i]);
    destroyComm(c);
    if (c->l2Base) { cudaCtxResetPersistingL2Cache(); cudaGetLastError(); } // hand the lines pinned by the L2 window back to the normal L2
    void *ptrs[] = {c->sobol, c->scrambling, c->ranking, c->idsChunk, c->idsLinear, c->occ, c->materials, c->blockToMaterial, c->skyArena,
                    c->illumination, c->illumOutput, c->ping, c->pong, c->prevIllum, c->prevFastIllum,
                    c->historyLength, c->prevHistoryLength, c->reservoirs, c->primaryHits, c->counters, c->patches, c->patchCount, c->wave.arena, c->upHDev, c->occPrev, c->pickDev, c->texels, c->texDescs, c->matTexSlots, c->matTexMip0Size, c->dnG, c->dnMQ, c->dnCounters, c->fireflyList, c->fixList, c->rgb8, c->dLights, c->dLightAlias, c->dFaceKeys, c->dPrevToCur};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int s = 0; s < 2; ++s)
    {
        void *g[] = {c->gb[s].depth, c->gb[s].material, c->gb[s].normalRoughness, c->gb[s].geoNormalThinfilm, c->gb[s].materialParameter, c->gb[s].albedo};
        for (void *p : g) if (p) cudaFree(p);
    }
    if (c->pinned) cudaFreeHost(c->pinned);
    for (int i = 0; i < EV_COUNT; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i <= TraceProfile::kMax; ++i) if (c->traceProf.ev[i]) cudaEventDestroy(c->traceProf.ev[i]);
    for (auto &p : c->atrousEv) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (int i = 0; i < 2; ++i)
    {
        if (c->traceStreams.join[i]) cudaEventDestroy(c->traceStreams.join[i]);
        if (c->traceStreams.part[i]) cudaStreamDestroy(c->traceStreams.part[i]);
    }
    if (c->traceStreams.fork) cudaEventDestroy(c->traceStreams.fork);
    if (c->copyReady) cudaEventDestroy(c->copyReady);
    if (c->copyDone) cudaEventDestroy(c->copyDone);
    if (c->copyStream) cudaStreamDestroy(c->copyStream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int vpt_sync(vpt_ctx *c)
{
    if (c && c->copyStream) { cudaSetDevice(c->device); cudaStreamSynchronize(c->copyStream); c->copyPending = false; c->copyPendingTraceWritten = false; }
    if (!c) return fail(VPT_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}
void *vpt_stream(vpt_ctx *c) { return c ? (void *)c->stream : nullptr; }
int vpt_set_profiling(vpt_ctx *c, int enabled) { if (!c) return VPT_ERR_ARG; c->profiling = enabled != 0; c->countSteps = enabled != 0; return VPT_OK; }

int vpt_set_tables(vpt_ctx *c, const uint8_t *sobol, const uint8_t *scrambling, const uint8_t *ranking)
{
    if (!c || !sobol || !scrambling || !ranking) return fail(VPT_ERR_ARG, "vpt_set_tables: null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->sobol, sobol, 65536, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->scrambling, scrambling, 131072, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->ranking, ranking, 131072, cudaMemcpyHostToDevice, c->stream)); // +256 zero bytes of padding stay 0
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}

static int allocGrid(vpt_ctx *c, int cx, int cy, int cz)
{
    if (cx <= 0 || cy <= 0 || cz <= 0) return fail(VPT_ERR_ARG, "grid: chunk counts must be positive");
    if (cx > 32 || cz > 32 || cy > 16) return fail(VPT_ERR_ARG, "grid: at most 32 x 16 x 32 chunks (1024 x 512 x 1024 voxels)");
    if (c->cx != cx || c->cy != cy || c->cz != cz)
    {
        if (c->idsChunk) cudaFree(c->idsChunk);
        if (c->idsLinear) cudaFree(c->idsLinear);
        if (c->occ) cudaFree(c->occ);
        c->idsChunk = c->idsLinear = nullptr; c->occ = nullptr;
        const size_t vox = (size_t)cx * cy * cz * 32768;
        CU(cudaMalloc((void **)&c->idsChunk, vox));
        CU(cudaMalloc((void **)&c->idsLinear, vox));
        CU(cudaMalloc((void **)&c->occ, paddedOccWords(cx * 32, cy * 32, cz * 32) * 4));
        if (!c->upHDev) CU(cudaMalloc((void **)&c->upHDev, sizeof(int)));
        c->cx = cx; c->cy = cy; c->cz = cz;
    }
    return VPT_OK;
}

int vpt_set_grid(vpt_ctx *c, int cx, int cy, int cz, const uint8_t *ids)
{
    if (!c || !ids) return fail(VPT_ERR_ARG, "vpt_set_grid: null argument");
    CU(cudaSetDevice(c->device));
    int rc = allocGrid(c, cx, cy, cz);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->idsChunk, ids, (size_t)cx * cy * cz * 32768, cudaMemcpyHostToDevice, c->stream));
    c->prevSnapshot = false; // a new world has no previous state
    c->lightsStale = true;
    CU(launchRepackGrid(c->idsChunk, c->idsLinear, c->occ, c->upHDev, &c->upH, cx, cy, cz, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}
int vpt_generate_terrain(vpt_ctx *c, int cx, int cy, int cz, const float *noise)
{
    if (!c || !noise) return fail(VPT_ERR_ARG, "vpt_generate_terrain: null argument");
    CU(cudaSetDevice(c->device));
    int rc = allocGrid(c, cx, cy, cz);
    if (rc) return rc;
    float *dNoise = nullptr;
    const size_t nb = (size_t)cx * cy * cz * 1024 * sizeof(float);
    CU(cudaMalloc((void **)&dNoise, nb));
    CU(cudaMemcpyAsync(dNoise, noise, nb, cudaMemcpyHostToDevice, c->stream));
    CU(launchGenerateTerrain(dNoise, c->idsChunk, cx, cy, cz, c->stream));
    c->prevSnapshot = false; // a new world has no previous state
    c->lightsStale = true;
    CU(launchRepackGrid(c->idsChunk, c->idsLinear, c->occ, c->upHDev, &c->upH, cx, cy, cz, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(dNoise));
    return VPT_OK;
}
int vpt_get_grid(vpt_ctx *c, uint8_t *out, size_t bytes)
{
    if (!c || !out) return fail(VPT_ERR_ARG, "vpt_get_grid: null argument");
    if (!c->idsChunk) return fail(VPT_ERR_STATE, "vpt_get_grid: no grid set");
    if (bytes != (size_t)c->cx * c->cy * c->cz * 32768) return fail(VPT_ERR_ARG, "vpt_get_grid: size mismatch");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(out, c->idsChunk, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}
int vpt_set_voxel(vpt_ctx *c, int x, int y, int z, int blockId)
{
    if (!c || !c->idsChunk) return fail(VPT_ERR_STATE, "vpt_set_voxel: no grid set");
    if (x < 0 || y < 0 || z < 0 || x >= c->cx * 32 || y >= c->cy * 32 || z >= c->cz * 32) return VPT_OK; // reference ignores out-of-range edits
    CU(cudaSetDevice(c->device));
    if (!c->prevSnapshot)
    {
        const size_t words = paddedOccWords(c->cx * 32, c->cy * 32, c->cz * 32);
        if (c->occPrevWords != words) { if (c->occPrev) cudaFree(c->occPrev); c->occPrev = nullptr; CU(cudaMalloc((void **)&c->occPrev, words * 4)); c->occPrevWords = words; }
        CU(cudaMemcpyAsync(c->occPrev, c->occ, words * 4, cudaMemcpyDeviceToDevice, c->stream));
        c->upHPrev = c->upH; c->prevSnapshot = true;
    }
    CU(launchSetVoxel(c->idsChunk, c->idsLinear, c->occ, &c->upH, c->cx, c->cy, c->cz, x, y, z, blockId, c->stream));
    c->lightsStale = true;
    return VPT_OK;
}

int vpt_pick_voxel(vpt_ctx *c, const float *origin, const float *direction, VptPickResult *out)
{
    if (!c || !origin || !direction || !out) return fail(VPT_ERR_ARG, "vpt_pick_voxel: null argument");
    if (!c->idsLinear) return fail(VPT_ERR_STATE, "vpt_pick_voxel: no grid set");
    CU(cudaSetDevice(c->device));
    if (!c->pickDev) CU(cudaMalloc((void **)&c->pickDev, sizeof(VptPickResult)));
    CU(launchPick(c->idsLinear, c->cx, c->cy, c->cz, origin, direction, c->pickDev, c->stream));
    CU(cudaMemcpyAsync(out, c->pickDev, sizeof(VptPickResult), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}

int vpt_set_materials(vpt_ctx *c, const VptMaterial *m, int count, const uint16_t *b2m)
{
    if (!c || !m || !b2m || count <= 0) return fail(VPT_ERR_ARG, "vpt_set_materials: bad argument");
    for (int i = 0; i < 256; ++i) if (b2m[i] >= count) return fail(VPT_ERR_ARG, "vpt_set_materials: blockToMaterial index out of range");
    CU(cudaSetDevice(c->device));
    if (c->materials) { CU(cudaStreamSynchronize(c->stream)); cudaFree(c->materials); c->materials = nullptr; }
    CU(cudaMalloc((void **)&c->materials, (size_t)count * sizeof(VptMaterial)));
    c->nMaterials = count;
    c->nTextures = 0; c->texSpecular = false; // texture slots are per material: vpt_set_textures must follow
    c->anySpecular = false;
    for (int i = 0; i < count; ++i)
        if (!m[i].isEmissive && !(m[i].roughness > 0.00001f)) c->anySpecular = true; // isDiffuse = roughness > 1e-5 (Bsdf.h:5)
    CU(cudaMemcpyAsync(c->materials, m, (size_t)count * sizeof(VptMaterial), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->blockToMaterial, b2m, 256 * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->hostMaterials.assign(m, m + count); std::memcpy(c->hostB2m, b2m, sizeof c->hostB2m);
    c->lightsStale = true;
    return VPT_OK;
}

// One L2 access-policy window per context, on its stream (hits persisting, misses streaming). A hint: failures are ignored.
static void setL2Window(vpt_ctx *c, const void *base, size_t bytes)
{
#ifndef VPT_NO_L2_WINDOW
    if (!base || bytes == 0 || (c->l2Base == base && c->l2Bytes == bytes)) return;
    int maxPersist = 0, maxWindow = 0;
    cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
    cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
    if (maxPersist > 0 && maxWindow > 0)
    {
        const size_t win = bytes < (size_t)maxWindow ? bytes : (size_t)maxWindow;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, win < (size_t)maxPersist ? win : (size_t)maxPersist);
        cudaStreamAttrValue v = {};
        v.accessPolicyWindow.base_ptr = const_cast<void *>(base);
        v.accessPolicyWindow.num_bytes = win;
        v.accessPolicyWindow.hitRatio = win <= (size_t)maxPersist ? 1.0f : (float)maxPersist / (float)win;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &v);
        cudaGetLastError();
    }
    c->l2Base = base; c->l2Bytes = bytes;
#else
    (void)c; (void)base; (void)bytes;
#endif
}

// Sky / sun maps and their alias tables in ONE allocation, pinned in L2 as far as the device allows: the stages sample them at random
// (alias bin -> texel, twice per path) while 6 GB of wavefront state stream through the same 126 MB L2 every frame; an access-policy
// window keeps the 12 MB of tables resident (hit = persisting) instead of letting the stream evict them. A hint: failures are ignored.
static int allocSkyArena(vpt_ctx *c, size_t ns, size_t nu)
{
    if (c->skyArena) cudaFree(c->skyArena);
    c->skyArena = nullptr; c->sky = c->sun = nullptr; c->skyAlias = c->sunAlias = nullptr;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t oSky = 0, oSkyAlias = oSky + up(ns * 16), oSun = oSkyAlias + up(ns * sizeof(VptAliasBin)), oSunAlias = oSun + up(nu * 16),
                 total = oSunAlias + up(nu * sizeof(VptAliasBin));
    CU(cudaMalloc(&c->skyArena, total));
    char *b = static_cast<char *>(c->skyArena);
    c->sky = reinterpret_cast<float4 *>(b + oSky); c->skyAlias = reinterpret_cast<VptAliasBin *>(b + oSkyAlias);
    c->sun = reinterpret_cast<float4 *>(b + oSun); c->sunAlias = reinterpret_cast<VptAliasBin *>(b + oSunAlias);
    c->skyArenaBytes = total;
    c->l2Base = nullptr; // the old window (if any) pointed into the freed arena: the next render sets it again
    return VPT_OK;
}

int vpt_set_sky(vpt_ctx *c, const float *sky, int skyW, int skyH, const float *sun, int sunW, int sunH,
                const VptAliasBin *skyAlias, const VptAliasBin *sunAlias, const float *sunDir)
{
    if (!c || !sky || !sun || !skyAlias || !sunAlias || !sunDir || skyW <= 0 || skyH <= 0 || sunW <= 0 || sunH <= 0)
        return fail(VPT_ERR_ARG, "vpt_set_sky: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    const size_t ns = (size_t)skyW * skyH, nu = (size_t)sunW * sunH;
    if (int rc = allocSkyArena(c, ns, nu)) return rc;
    CU(cudaMemcpyAsync(c->sky, sky, ns * 16, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->sun, sun, nu * 16, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->skyAlias, skyAlias, ns * sizeof(VptAliasBin), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->sunAlias, sunAlias, nu * sizeof(VptAliasBin), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->skyW = skyW; c->skyH = skyH; c->sunW = sunW; c->sunH = sunH;
    c->sunDir[0] = sunDir[0]; c->sunDir[1] = sunDir[1]; c->sunDir[2] = sunDir[2];
    return VPT_OK;
}

int vpt_set_textures(vpt_ctx *c, int nTextures, const int32_t *widths, const int32_t *levels, const uint32_t *texels, int nMaterials,
                     const int32_t *slots4, const float *texSize2)
{
    if (!c || nTextures < 0) return fail(VPT_ERR_ARG, "vpt_set_textures: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (void *p : {(void *)c->texels, (void *)c->texDescs, (void *)c->matTexSlots, (void *)c->matTexMip0Size}) if (p) cudaFree(p);
    c->texels = nullptr; c->texDescs = nullptr; c->matTexSlots = nullptr; c->matTexMip0Size = nullptr; c->nTextures = 0; c->texSpecular = false;
    if (nTextures == 0) return VPT_OK;
    if (!widths || !levels || !texels || !slots4 || !texSize2) return fail(VPT_ERR_ARG, "vpt_set_textures: null argument");
    if (nMaterials != c->nMaterials) return fail(VPT_ERR_ARG, "vpt_set_textures: material count differs from vpt_set_materials");
    std::vector<int4> descs((size_t)nTextures);
    size_t total = 0;
    for (int t = 0; t < nTextures; ++t)
    {
        const int w = widths[t], l = levels[t];
        if (w <= 0 || (w & (w - 1)) || l < 1 || (w >> (l - 1)) < 1) return fail(VPT_ERR_ARG, "vpt_set_textures: textures must be square powers of two with 1..log2(w)+1 levels");
        if (total > 0x7fffffffu) return fail(VPT_ERR_ARG, "vpt_set_textures: more than 2^31 texels");
        descs[(size_t)t] = make_int4((int)total, w, l, 0);
        for (int k = 0; k < l; ++k) total += (size_t)(w >> k) * (w >> k);
    }
    std::vector<int4> slots((size_t)nMaterials);
    std::vector<float> mip0((size_t)nMaterials);
    for (int m = 0; m < nMaterials; ++m)
    {
        for (int k = 0; k < 4; ++k) if (slots4[m * 4 + k] >= nTextures) return fail(VPT_ERR_ARG, "vpt_set_textures: texture index out of range");
        slots[(size_t)m] = make_int4(slots4[m * 4], slots4[m * 4 + 1], slots4[m * 4 + 2], slots4[m * 4 + 3]);
        mip0[(size_t)m] = sqrtf(texSize2[m * 2] * texSize2[m * 2] + texSize2[m * 2 + 1] * texSize2[m * 2 + 1]); // texSize.length(), host IEEE like the oracle
    }
    CU(cudaMalloc((void **)&c->texels, total * 4)); CU(cudaMalloc((void **)&c->texDescs, descs.size() * sizeof(int4)));
    CU(cudaMalloc((void **)&c->matTexSlots, slots.size() * sizeof(int4))); CU(cudaMalloc((void **)&c->matTexMip0Size, mip0.size() * sizeof(float)));
    CU(cudaMemcpyAsync(c->texels, texels, total * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->texDescs, descs.data(), descs.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->matTexSlots, slots.data(), slots.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->matTexMip0Size, mip0.data(), mip0.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->nTextures = nTextures;
    for (int m = 0; m < nMaterials; ++m) if (slots4[m * 4 + 2] >= 0) c->texSpecular = true;
    return VPT_OK;
}
int vpt_set_wave_budget(vpt_ctx *c, size_t maxPaths)
{
    if (!c || maxPaths == 0) return fail(VPT_ERR_ARG, "vpt_set_wave_budget: bad argument");
    c->waveBudget = maxPaths;
    return VPT_OK;
}
int vpt_generate_sky(vpt_ctx *c, const VptSkyParams *params, const float *tables)
{
    if (!c || !params || !tables) return fail(VPT_ERR_ARG, "vpt_generate_sky: null argument");
    CU(cudaSetDevice(c->device));
    const int skyW = 1024, skyH = 512, sunW = 32, sunH = 32; // SkyModel::skyRes / sunRes (renderer/sky/Sky.h:51-52)
    const size_t ns = (size_t)skyW * skyH, nu = (size_t)sunW * sunH;
    float configs[90], radiances[10], sunDir[3];
    vpt_sky_state(params, tables, configs, radiances, sunDir);
    if (int rc = allocSkyArena(c, ns, nu)) return rc;
    float *dPdf = nullptr, *dTab = nullptr;
    CU(cudaMalloc((void **)&dPdf, (ns + nu) * sizeof(float)));
    CU(cudaMalloc((void **)&dTab, 1860 * sizeof(float)));
    CU(cudaMemcpyAsync(dTab, tables + 600, 1860 * sizeof(float), cudaMemcpyHostToDevice, c->stream)); // solar[1800], limb[60]
    std::vector<float> pdf(ns + nu);
    CU(launchSkyUpper(configs, radiances, sunDir, params->skyBrightness, c->sky, dPdf, skyW, skyH, c->stream));
    // thrust::reduce over the upper hemisphere (Sky.cu:378): summed on the host in double, rounded once
    CU(cudaMemcpyAsync(pdf.data() + ns / 2, dPdf + ns / 2, (ns / 2) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    double sum = 0.0;
    for (size_t i = ns / 2; i < ns; ++i) sum += (double)pdf[i];
    CU(launchSkyLower(c->sky, dPdf, skyW, skyH, (float)sum, c->stream));
    CU(launchSkySun(sunDir, params->skyBrightness, dTab, dTab + 1800, c->sun, dPdf + ns, sunW, sunH, c->stream));
    CU(cudaMemcpyAsync(pdf.data(), dPdf, (ns + nu) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<VptAliasBin> skyBins(ns), sunBins(nu);
    vpt_build_alias_table(pdf.data(), (unsigned)ns, skyBins.data());
    vpt_build_alias_table(pdf.data() + ns, (unsigned)nu, sunBins.data());
    CU(cudaMemcpyAsync(c->skyAlias, skyBins.data(), ns * sizeof(VptAliasBin), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->sunAlias, sunBins.data(), nu * sizeof(VptAliasBin), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(dPdf)); CU(cudaFree(dTab));
    c->skyW = skyW; c->skyH = skyH; c->sunW = sunW; c->sunH = sunH;
    c->sunDir[0] = sunDir[0]; c->sunDir[1] = sunDir[1]; c->sunDir[2] = sunDir[2];
    return VPT_OK;
}
int vpt_sky_size(vpt_ctx *c, int *skyW, int *skyH, int *sunW, int *sunH)
{
    if (!c || !c->sky) return fail(VPT_ERR_STATE, "vpt_sky_size: no sky set");
    if (skyW) *skyW = c->skyW; if (skyH) *skyH = c->skyH; if (sunW) *sunW = c->sunW; if (sunH) *sunH = c->sunH;
    return VPT_OK;
}
int vpt_read_sky(vpt_ctx *c, float *skyRGBA, float *sunRGBA, float *sunDir3)
{
    if (!c || !c->sky) return fail(VPT_ERR_STATE, "vpt_read_sky: no sky set");
    CU(cudaSetDevice(c->device));
    if (skyRGBA) CU(cudaMemcpyAsync(skyRGBA, c->sky, (size_t)c->skyW * c->skyH * 16, cudaMemcpyDeviceToHost, c->stream));
    if (sunRGBA) CU(cudaMemcpyAsync(sunRGBA, c->sun, (size_t)c->sunW * c->sunH * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (sunDir3) { sunDir3[0] = c->sunDir[0]; sunDir3[1] = c->sunDir[1]; sunDir3[2] = c->sunDir[2]; }
    return VPT_OK;
}

int vpt_set_trace_params(vpt_ctx *c, int spp, int totalBounceLimit, int diffuseBounceLimit, int enableRestir)
{
    if (!c || spp < 1 || totalBounceLimit < 1 || diffuseBounceLimit < 1) return fail(VPT_ERR_ARG, "vpt_set_trace_params: bad argument");
    // every depth round owns three DDA count/cursor pairs and one active-list counter of the per-part counter block (vpt_wave.cu:
    // kCntList, kCntWords); 16 is also the reference UI's range for the bounce sliders
    if (totalBounceLimit > 16 || diffuseBounceLimit > 16) return fail(VPT_ERR_ARG, "vpt_set_trace_params: bounce limits above 16 are not supported");
    c->spp = spp; c->totalBounceLimit = totalBounceLimit; c->diffuseBounceLimit = diffuseBounceLimit; c->enableRestir = enableRestir ? 1 : 0;
    return VPT_OK;
}

// Bring the local-light list up to date (the reference rebuilds its LightInfo buffer + alias table on every scene update,
// VoxelEngine.cu:53-192, synchronously). Only scenes with an emissive material pay for the grid read-back.
static int refreshLights(vpt_ctx *c)
{
    if (!c->lightsStale) return VPT_OK;
    c->lightsStale = false;
    bool anyEmissive = false;
    for (const VptMaterial &m : c->hostMaterials) anyEmissive = anyEmissive || m.isEmissive != 0;
    c->lights = LightList();
    if (anyEmissive && c->idsChunk)
    {
        std::vector<uint8_t> ids((size_t)c->cx * c->cy * c->cz * 32768);
        CU(cudaMemcpyAsync(ids.data(), c->idsChunk, ids.size(), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        buildLightList(ids.data(), c->cx, c->cy, c->cz, c->hostMaterials.data(), (int)c->hostMaterials.size(), c->hostB2m, c->lights);
    }
    const size_t n = c->lights.lights.size();
    if (n > c->dLightCap)
    {
        CU(cudaStreamSynchronize(c->stream));
        for (void *p : {(void *)c->dLights, (void *)c->dLightAlias, (void *)c->dFaceKeys}) if (p) cudaFree(p);
        c->dLights = nullptr; c->dLightAlias = nullptr; c->dFaceKeys = nullptr; c->dLightCap = 0;
        const size_t cap = n + n / 2 + 64;
        CU(cudaMalloc((void **)&c->dLights, cap * sizeof(VptLightInfo)));
        CU(cudaMalloc((void **)&c->dLightAlias, cap * sizeof(VptAliasBin)));
        CU(cudaMalloc((void **)&c->dFaceKeys, cap / 2 * sizeof(uint32_t) + 4));
        c->dLightCap = cap;
    }
    if (n)
    {
        CU(cudaMemcpyAsync(c->dLights, c->lights.lights.data(), n * sizeof(VptLightInfo), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->dLightAlias, c->lights.alias.data(), n * sizeof(VptAliasBin), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->dFaceKeys, c->lights.faceKeys.data(), c->lights.faceKeys.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream)); // pageable sources
    }
    return VPT_OK;
}
// Before a render: if the list differs from the one the previous render used, the stored reservoirs hold stale light ids — upload
// the previous -> current id table (Restir.h:60-75) and raise lightsStateDirty for this frame.
static int prepareLightRemap(vpt_ctx *c)
{
    c->lightsStateDirty = c->renderedKeys != c->lights.faceKeys;
    if (!c->lightsStateDirty) return VPT_OK;
    buildLightRemap(c->renderedKeys, c->lights.faceKeys, c->prevLightToCur);
    c->prevNumLights = (int)c->renderedKeys.size() * 2;
    if (c->prevLightToCur.size() > c->dRemapCap)
    {
        CU(cudaStreamSynchronize(c->stream));
        if (c->dPrevToCur) cudaFree(c->dPrevToCur);
        c->dPrevToCur = nullptr; c->dRemapCap = 0;
        const size_t cap = c->prevLightToCur.size() * 2 + 64;
        CU(cudaMalloc((void **)&c->dPrevToCur, cap * sizeof(int)));
        c->dRemapCap = cap;
    }
    if (!c->prevLightToCur.empty())
    {
        CU(cudaMemcpyAsync(c->dPrevToCur, c->prevLightToCur.data(), c->prevLightToCur.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    c->renderedKeys = c->lights.faceKeys; // from this render on the reservoirs speak the new ids
    return VPT_OK;
}

static int renderImpl(vpt_ctx *c, const VptCamera *cam, const VptCamera *prevCam, int iterationIndex, int sampleBegin, int sampleStep, bool resolve, bool localOwner = false,
                      int sampleLimit = 0)
{
    if (!c || !cam || !prevCam || sampleBegin < 0 || sampleStep < 1) return fail(VPT_ERR_ARG, "vpt_render: bad argument");
    if (!c->occ) return fail(VPT_ERR_STATE, "vpt_render: no voxel grid (vpt_set_grid / vpt_generate_terrain)");
    if (!c->materials) return fail(VPT_ERR_STATE, "vpt_render: no materials (vpt_set_materials)");
    if (!c->sky) return fail(VPT_ERR_STATE, "vpt_render: no sky (vpt_set_sky)");
    if ((int)cam->resolution[0] != c->width || (int)cam->resolution[1] != c->height) return fail(VPT_ERR_ARG, "vpt_render: camera resolution != context size");
    CU(cudaSetDevice(c->device));
    if (c->copyPending && c->copyPendingTraceWritten) CU(waitPendingCopy(c));
    { int rcl = refreshLights(c); if (!rcl) rcl = prepareLightRemap(c); if (rcl) return rcl; }
    c->cur ^= 1;
    TraceArgs a;
    std::memset(&a, 0, sizeof a);
    a.cam = *cam; a.prevCam = *prevCam;
    a.width = c->width; a.height = c->height;
    a.iterationIndex = iterationIndex; a.spp = c->spp; a.totalBounceLimit = c->totalBounceLimit; a.diffuseBounceLimit = c->diffuseBounceLimit;
    a.enableRestir = c->enableRestir; a.sampleBegin = sampleBegin; a.sampleStep = sampleStep;
    a.ownerSample = localOwner ? sampleBegin : 0;
    a.grid.W = c->cx * 32; a.grid.H = c->cy * 32; a.grid.D = c->cz * 32;
    a.grid.Wp = paddedW(a.grid.W); a.grid.Hp = a.grid.H + 2; a.grid.Dp = a.grid.D + 2;
    a.grid.maskWords = (int)paddedMaskWords(a.grid.W, a.grid.H, a.grid.D);
    a.grid.occWords = 2 * a.grid.maskWords + 4; a.grid.parkLin = 2 * a.grid.maskWords * 32;
    a.grid.upH = c->upH;
    a.occPrev = c->prevSnapshot ? c->occPrev : nullptr; a.upHPrev = c->upHPrev;
    a.grid.occ = c->occ; a.grid.idsLinear = c->idsLinear;
    a.grid.divW = makeFastDiv(a.grid.W); a.grid.divD = makeFastDiv(a.grid.D);
    a.grid.divWp = makeFastDiv(a.grid.Wp); a.grid.divDp = makeFastDiv(a.grid.Dp);
    a.tilesX = (c->width + 7) / 8;
    a.nSlots = a.tilesX * ((c->height + 3) / 4) * 32;
    a.divSlots = makeFastDiv(a.nSlots); a.divTilesX = makeFastDiv(a.tilesX);
    a.divSkyW = makeFastDiv(c->skyW); a.divSunW = makeFastDiv(c->sunW);
    // a path continues past its first hit only through a specular surface or with a diffuse limit above 1
    a.depthRounds = (c->anySpecular || c->texSpecular || c->diffuseBounceLimit > 1) ? c->totalBounceLimit : 1;
    a.countSteps = c->countSteps;
    a.resolveSpp = (resolve && c->spp > 1) ? (float)c->spp : 0.0f;
    // wave size: as many samples per wave as fit a 16 M-path budget
    int shardSamples = sampleBegin < c->spp ? (c->spp - sampleBegin + sampleStep - 1) / sampleStep : 0;
    if (sampleLimit > 0 && shardSamples > sampleLimit) shardSamples = sampleLimit;
    a.sampleLimit = sampleLimit;
    int samplesPerWave = (int)(c->waveBudget / (size_t)a.nSlots);
    if (samplesPerWave < 1) samplesPerWave = 1;
    if (samplesPerWave > shardSamples) samplesPerWave = shardSamples > 0 ? shardSamples : 1;
    if (c->wave.nSlots != a.nSlots || c->wave.maxSamplesInWave < samplesPerWave)
    {
        if (c->wave.arena) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->wave.arena)); c->wave.arena = nullptr; }
        c->wave.arenaBytes = waveWorkspaceBytes(a.nSlots, samplesPerWave);
        CU(cudaMalloc(&c->wave.arena, c->wave.arenaBytes));
        waveCarve(c->wave, a.nSlots, samplesPerWave);
    }
    a.wb = c->wave.wb;
    a.sobol = c->sobol; a.scrambling = c->scrambling; a.ranking = c->ranking;
    a.materials = c->materials; a.blockToMaterial = c->blockToMaterial;
    a.texels = c->texels; a.texDescs = c->texDescs; a.matTexSlots = c->matTexSlots; a.matTexMip0Size = c->matTexMip0Size; a.nTextures = c->nTextures;
    a.sky = c->sky; a.sun = c->sun; a.skyAlias = c->skyAlias; a.sunAlias = c->sunAlias;
    a.skyW = c->skyW; a.skyH = c->skyH; a.sunW = c->sunW; a.sunH = c->sunH;
    a.sunDir[0] = c->sunDir[0]; a.sunDir[1] = c->sunDir[1]; a.sunDir[2] = c->sunDir[2];
    a.sunCosThetaMax = cosf(0.51f * 3.1415926535897932384626422832795028841971f / 180.0f / 2.0f); // miss.cu:46-47, host libm like the oracle
    a.lv.lights = c->dLights; a.lv.alias = c->dLightAlias; a.lv.faceKeys = c->dFaceKeys; a.lv.prevToCur = c->dPrevToCur;
    a.lv.numLights = (int)c->lights.lights.size(); a.lv.numFaces = (int)c->lights.faceKeys.size();
    a.lv.prevNumLights = c->prevNumLights; a.lv.stateDirty = c->lightsStateDirty ? 1 : 0;
    a.cur = c->gb[c->cur].ptrs(); a.prev = c->gb[c->cur ^ 1].ptrs();
    a.illumination = c->illumination;
    a.resCur = c->reservoirs + (size_t)(iterationIndex & 1) * c->npix();
    a.resPrev = c->reservoirs + (size_t)((iterationIndex + 1) & 1) * c->npix();
    a.primaryHits = c->primaryHits;
    a.counters = c->counters;
    CU(cudaMemsetAsync(c->counters, 0, 2 * sizeof(unsigned long long), c->stream)); // [0] rays, [1] steps of this frame; [2] = running total of rays
    if (c->profiling) CU(cudaEventRecord(c->ev[EV_TRACE0], c->stream));
    // more ranks than samples: this shard renders nothing, so its term of the cross-rank sum must be zero (no wave runs, so no
    // accumulate pass would otherwise overwrite the previous frame's image)
    if (shardSamples == 0) CU(cudaMemsetAsync(c->illumination, 0, c->npix() * sizeof(float4), c->stream));
    // The L2 window goes to the sky tables: the stages read them at random and the streaming wavefront state would evict them
    // (stages 1.311 -> 1.292 ms on cfg2). Not for worlds whose traversal masks are walked through L1/L2: those want the whole L2 —
    // pinning the masks themselves (2 x 33 MiB) measured 51.5 vs 48.3 ms per frame on the cfg5-shaped trace (tools/cfg5_quick.py).
    if ((size_t)a.grid.occWords * 4 + 1024 <= c->smemOptIn) setL2Window(c, c->skyArena, c->skyArenaBytes);
    int launches = 0;
    c->traceProf.enabled = c->profiling;
    CU(launchTrace(a, c->wave.maxSamplesInWave, c->stream, c->overlapParts ? &c->traceStreams : nullptr, c->smCount, c->smemOptIn, &launches, &c->traceProf));
    c->prevSnapshot = false; // the next frame's "previous world" is this one unless an edit takes a new snapshot
    if (c->profiling) CU(cudaEventRecord(c->ev[EV_TRACE1], c->stream));
    c->haveTrace = c->profiling; c->ranResolve = false; c->launchesRender = launches;
    return VPT_OK;
}
int vpt_resolve(vpt_ctx *c)
{
    if (!c) return fail(VPT_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (c->spp > 1)
    {
        CU(launchResolve(c->illumination, (int)c->npix(), (float)c->spp, c->stream));
        if (c->profiling) CU(cudaEventRecord(c->ev[EV_RESOLVE1], c->stream));
        c->ranResolve = c->profiling; c->launchesRender += 1;
    }
    return VPT_OK;
}
int vpt_render_shard(vpt_ctx *c, const VptCamera *cam, const VptCamera *prevCam, int iterationIndex, int sampleBegin, int sampleStep)
{
    return renderImpl(c, cam, prevCam, iterationIndex, sampleBegin, sampleStep, false);
}
int vpt_render_shard_local(vpt_ctx *c, const VptCamera *cam, const VptCamera *prevCam, int iterationIndex, int sampleBegin, int sampleStep)
{
    return renderImpl(c, cam, prevCam, iterationIndex, sampleBegin, sampleStep, false, true);
}
int vpt_render_range(vpt_ctx *c, const VptCamera *cam, const VptCamera *prevCam, int iterationIndex, int sampleBegin, int sampleCount)
{
    if (sampleCount < 0) return fail(VPT_ERR_ARG, "vpt_render_range: negative sample count");
    if (sampleCount == 0) return renderImpl(c, cam, prevCam, iterationIndex, c ? c->spp : 0, 1, false); // an empty shard: contributes zero
    return renderImpl(c, cam, prevCam, iterationIndex, sampleBegin, 1, false, false, sampleCount);
}
int vpt_render(vpt_ctx *c, const VptCamera *cam, const VptCamera *prevCam, int iterationIndex)
{
    // == vpt_render_shard(0, 1) + vpt_resolve, with the division fused into the last accumulate
    return renderImpl(c, cam, prevCam, iterationIndex, 0, 1, true);
}
int vpt_begin_external_frame(vpt_ctx *c) { if (!c) return VPT_ERR_ARG; c->cur ^= 1; return VPT_OK; }

// Denoiser::run (renderer/denoising/Denoiser.cu:24-408) over rows [rowBegin,rowEnd). band: the rows are one rank's EXTENDED band
// (vpt_denoise_band): the chain runs without any exchange and its result is exact on the rows whose whole dependency
// cone lies inside [rowBegin,rowEnd) — the band proper; the packed G-buffer is prepared on a guard around the rows.
static int denoiseChain(vpt_ctx *c, const VptDenoisingParams *p, const VptCamera *cam, const VptCamera *prevCam, int frameNum, int iterationIndex,
                        int rowBegin, int rowEnd, bool timing, bool band)
{
    DenoiseLaunch d;
    d.width = c->width; d.height = c->height; d.rowBegin = rowBegin; d.rowEnd = rowEnd;
    d.cam = *cam; d.prevCam = *prevCam; d.p = *p; d.stream = c->stream;
    d.b.cur = c->gb[c->cur].ptrs(); d.b.prev = c->gb[c->cur ^ 1].ptrs();
    d.b.illumination = c->illumination; d.b.illumOutput = c->illumOutput; d.b.ping = c->ping; d.b.pong = c->pong;
    d.b.prevIllum = c->prevIllum; d.b.prevFastIllum = c->prevFastIllum; d.b.historyLength = c->historyLength; d.b.prevHistoryLength = c->prevHistoryLength;
    d.view = makeDnView(*cam); d.G = c->dnG; d.MQ = c->dnMQ; d.counters = c->dnCounters; d.fireflyList = c->fireflyList; d.fixList = c->fixList;
    const int usedIter = iterationIndex > 0 ? iterationIndex - 1 : 0;
    d.b.reservoirs = c->reservoirs + (size_t)(usedIter & 1) * c->npix();
    auto rec = [&](int e) -> cudaError_t { return timing ? cudaEventRecord(c->ev[e], c->stream) : cudaSuccess; };
    int launches = 0, rc;
    c->ranFirefly = c->ranTemporal = c->ranFix = c->ranClamp = c->ranSpatial = false;
    c->atrousPasses = 0;
    CU(rec(EV_DN0));
    // packed G-buffer + sky copy (+ firefly): the stencil passes read G/MQ up to 18 rows outside the band
    const int guard = 32;
    const int prep0 = band ? std::max(0, rowBegin - guard) : rowBegin, prep1 = band ? std::min(c->height, rowEnd + guard) : rowEnd;
    CU(launchPrep(d, prep0, prep1, p->enableFireflyFilter != 0, c->patches, c->maxPatches));
    launches += p->enableFireflyFilter ? 3 : 1; c->ranFirefly = true;
    CU(rec(EV_FIREFLY));
    int finalBuf = 0;
    // HitDistReconstruction -> Ping, PrePass Ping -> Illumination (Denoiser.cu:86-119; off in the shipped settings). Their
    // times are booked under the sky-copy stage.
    if (p->enableHitDistanceReconstruction) { CU(launchHitDist(d)); launches++; finalBuf = 1; }
    if (p->enablePrePass) { CU(launchPrePass(d, iterationIndex)); launches++; }
    if (frameNum == 0)
    {
        CU(launchFrame0Init(d)); launches++;
    }
    CU(rec(EV_SKY));
    if (p->enableTemporalAccumulation && frameNum > 0)
    {
        CU(launchTemporal(d)); launches++; finalBuf = 1; c->ranTemporal = true;
        CU(rec(EV_TEMPORAL));
        if (p->enableHistoryFix)
        {
            CU(launchHistoryFix(d, c->smCount)); launches++; finalBuf = 2; c->ranFix = true;
        }
        CU(rec(EV_HFIX));
        if (p->enableHistoryClamping)
        {
            if (c->dnGather || !tileClampEnabled()) CU(launchHistoryClamping(d)); else CU(launchHistoryClampingCols(d));
            launches++; finalBuf = 3; c->ranClamp = true;
        }
        CU(rec(EV_HCLAMP));
    }
    else { CU(rec(EV_TEMPORAL)); CU(rec(EV_HFIX)); CU(rec(EV_HCLAMP)); }
    bool composited = false;
    if (p->enableSpatialFiltering)
    {
        {
            bool tiled = false;
            if (!c->dnGather) CU(launchAtrousSmemTiled(d, c->prevIllum, c->ping, &tiled));
            if (!tiled) CU(launchAtrousSmem(d, c->prevIllum, c->ping));
        }
        launches++; finalBuf = 1; c->ranSpatial = true;
        CU(rec(EV_ASMEM));
        if (p->atrousIterationNum > 0)
        {
            int idx = 1, step = 1 << idx;
            const int maxIt = p->atrousIterationNum * 2;
            auto pass = [&](float4 *in, float4 *out, bool last) -> int {
                // the last pass multiplies by albedo and writes IlluminationOutput (BufferCopyNonSky fused)
                bool tiled = false;
                cudaError_t e = c->dnGather ? cudaSuccess : launchAtrousTiled(d, in, last ? c->illumOutput : out, (unsigned)iterationIndex, (unsigned)step, last, &tiled);
                if (e == cudaSuccess && !tiled) e = launchAtrous(d, in, last ? c->illumOutput : out, (unsigned)iterationIndex, (unsigned)step, last);
                if (e != cudaSuccess) return fail(VPT_ERR_CUDA, cudaGetErrorString(e));
                launches++; c->atrousPasses++;
                return VPT_OK;
            };
            while (idx < maxIt)
            {
                if ((rc = pass(c->ping, c->pong, false))) return rc;
                ++idx; step = 1 << idx;
                if ((rc = pass(c->pong, c->ping, false))) return rc;
                ++idx; step = 1 << idx;
            }
            if ((rc = pass(c->ping, c->pong, true))) return rc;
            finalBuf = 2; composited = true;
        }
        CU(rec(EV_ATROUS));
    }
    else { CU(rec(EV_ASMEM)); CU(rec(EV_ATROUS)); }
    if (!composited)
    {
        const float4 *fin = finalBuf == 1 ? c->ping : finalBuf == 2 ? c->pong : finalBuf == 3 ? c->prevIllum : c->illumination;
        CU(launchCompositeNonSky(d, fin)); launches++;
    }
    CU(rec(EV_COMP));
    (void)rc;
    c->launchesDenoise = launches;
    c->haveDenoise = timing;
    return VPT_OK;
}

int vpt_denoise(vpt_ctx *c, const VptDenoisingParams *p, const VptCamera *cam, const VptCamera *prevCam, int frameNum, int iterationIndex)
{
    if (!c || !p || !cam || !prevCam) return fail(VPT_ERR_ARG, "vpt_denoise: null argument");
    CU(cudaSetDevice(c->device));
    // a pipelined read-back of the previous frame's planes must finish before this chain overwrites them (device-side wait)
    CU(waitPendingCopy(c));
    return denoiseChain(c, p, cam, prevCam, frameNum, iterationIndex, 0, c->height, c->profiling, false);
}

static int planeInfo(vpt_ctx *c, VptBufferName name, void **ptr, size_t *bytes)
{
    const size_t n = c->npix();
    GSet &g = c->gb[c->cur], &pg = c->gb[c->cur ^ 1];
    switch (name)
    {
    case VPT_BUF_Illumination: *ptr = c->illumination; *bytes = n * 16; break;
    case VPT_BUF_IlluminationOutput: *ptr = c->illumOutput; *bytes = n * 16; break;
    case VPT_BUF_IlluminationPing: *ptr = c->ping; *bytes = n * 16; break;
    case VPT_BUF_IlluminationPong: *ptr = c->pong; *bytes = n * 16; break;
    case VPT_BUF_NormalRoughness: *ptr = g.normalRoughness; *bytes = n * 16; break;
    case VPT_BUF_Depth: *ptr = g.depth; *bytes = n * 4; break;
    case VPT_BUF_Material: *ptr = g.material; *bytes = n * 4; break;
    case VPT_BUF_Albedo: *ptr = g.albedo; *bytes = n * 16; break;
    case VPT_BUF_HistoryLength: *ptr = c->historyLength; *bytes = n * 4; break;
    case VPT_BUF_PrevDepth: *ptr = pg.depth; *bytes = n * 4; break;
    case VPT_BUF_PrevMaterial: *ptr = pg.material; *bytes = n * 4; break;
    case VPT_BUF_PrevIllumination: *ptr = c->prevIllum; *bytes = n * 16; break;
    case VPT_BUF_PrevFastIllumination: *ptr = c->prevFastIllum; *bytes = n * 16; break;
    case VPT_BUF_PrevHistoryLength: *ptr = c->prevHistoryLength; *bytes = n * 4; break;
    case VPT_BUF_PrevNormalRoughness: *ptr = pg.normalRoughness; *bytes = n * 16; break;
    case VPT_BUF_GeoNormalThinfilm: *ptr = g.geoNormalThinfilm; *bytes = n * 16; break;
    case VPT_BUF_MaterialParameter: *ptr = g.materialParameter; *bytes = n * 16; break;
    case VPT_BUF_PrevMaterialParameter: *ptr = pg.materialParameter; *bytes = n * 16; break;
    case VPT_BUF_PrevGeoNormalThinfilm: *ptr = pg.geoNormalThinfilm; *bytes = n * 16; break;
    case VPT_BUF_PrevAlbedo: *ptr = pg.albedo; *bytes = n * 16; break;
    case VPT_BUF_PrimaryHits: *ptr = c->primaryHits; *bytes = n * 16; break;
    default: return fail(VPT_ERR_ARG, "unknown buffer name");
    }
    return VPT_OK;
}
void *vpt_device_ptr(vpt_ctx *c, VptBufferName name)
{
    void *p = nullptr; size_t b = 0;
    if (!c || planeInfo(c, name, &p, &b)) return nullptr;
    return p;
}
int vpt_read_buffer(vpt_ctx *c, VptBufferName name, void *host, size_t bytes)
{
    if (!c || !host) return fail(VPT_ERR_ARG, "vpt_read_buffer: null argument");
    void *p; size_t have;
    int rc = planeInfo(c, name, &p, &have);
    if (rc) return rc;
    if (have != bytes) return fail(VPT_ERR_ARG, "vpt_read_buffer: size mismatch");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}
int vpt_read_buffer_async(vpt_ctx *c, VptBufferName name, void *host, size_t bytes)
{
    if (!c || !host) return fail(VPT_ERR_ARG, "vpt_read_buffer_async: null argument");
    void *p; size_t have;
    int rc = planeInfo(c, name, &p, &have);
    if (rc) return rc;
    if (have != bytes) return fail(VPT_ERR_ARG, "vpt_read_buffer_async: size mismatch");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->copyReady, c->stream));
    CU(cudaStreamWaitEvent(c->copyStream, c->copyReady, 0));
    CU(cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, c->copyStream));
    CU(cudaEventRecord(c->copyDone, c->copyStream));
    c->copyPending = true;
    // planes the trace writes must also survive until the copy is done; IlluminationOutput and the history planes are
    // only touched by the denoiser, so their copy may overlap the whole next trace
    c->copyPendingTraceWritten |= !(name == VPT_BUF_IlluminationOutput || name == VPT_BUF_IlluminationPing || name == VPT_BUF_IlluminationPong ||
                                   name == VPT_BUF_PrevIllumination || name == VPT_BUF_PrevFastIllumination || name == VPT_BUF_HistoryLength ||
                                   name == VPT_BUF_PrevHistoryLength);
    return VPT_OK;
}
int vpt_read_wait(vpt_ctx *c)
{
    if (!c) return fail(VPT_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->copyStream));
    c->copyPending = false; c->copyPendingTraceWritten = false;
    return VPT_OK;
}
int vpt_write_buffer(vpt_ctx *c, VptBufferName name, const void *host, size_t bytes)
{
    if (!c || !host) return fail(VPT_ERR_ARG, "vpt_write_buffer: null argument");
    void *p; size_t have;
    int rc = planeInfo(c, name, &p, &have);
    if (rc) return rc;
    if (have != bytes) return fail(VPT_ERR_ARG, "vpt_write_buffer: size mismatch");
    CU(cudaSetDevice(c->device));
    CU(waitPendingCopy(c));
    CU(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}
int vpt_write_buffer_device(vpt_ctx *c, VptBufferName name, const void *device, size_t bytes)
{
    if (!c || !device) return fail(VPT_ERR_ARG, "vpt_write_buffer_device: null argument");
    void *p; size_t have;
    int rc = planeInfo(c, name, &p, &have);
    if (rc) return rc;
    if (have != bytes) return fail(VPT_ERR_ARG, "vpt_write_buffer_device: size mismatch");
    CU(cudaSetDevice(c->device));
    CU(waitPendingCopy(c));
    CU(cudaMemcpyAsync(p, device, bytes, cudaMemcpyDeviceToDevice, c->stream)); // stays asynchronous: ordered on the context's stream
    return VPT_OK;
}
int vpt_read_reservoirs(vpt_ctx *c, int parity, VptReservoir *host, size_t bytes)
{
    if (!c || !host || bytes != c->npix() * sizeof(VptReservoir)) return fail(VPT_ERR_ARG, "vpt_read_reservoirs: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(host, c->reservoirs + (size_t)(parity & 1) * c->npix(), bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}
int vpt_write_reservoirs(vpt_ctx *c, int parity, const VptReservoir *host, size_t bytes)
{
    if (!c || !host || bytes != c->npix() * sizeof(VptReservoir)) return fail(VPT_ERR_ARG, "vpt_write_reservoirs: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->reservoirs + (size_t)(parity & 1) * c->npix(), host, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}

int vpt_denoise_external(vpt_ctx *c, const VptDenoisingParams *p, const VptCamera *cam, const VptCamera *prevCam, int frameNum, int iterationIndex,
                         const float *illumination, const float *depth, const float *normalRoughness, const float *material, const float *albedo,
                         float *outputRGBA)
{
    if (!c || !p || !cam || !prevCam || !illumination || !depth || !normalRoughness || !material || !albedo || !outputRGBA)
        return fail(VPT_ERR_ARG, "vpt_denoise_external: null argument");
    CU(cudaSetDevice(c->device));
    CU(waitPendingCopy(c));
    c->cur ^= 1;
    const size_t n = c->npix();
    GSet &g = c->gb[c->cur];
    CU(cudaMemcpyAsync(c->illumination, illumination, n * 16, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(g.depth, depth, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(g.normalRoughness, normalRoughness, n * 16, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(g.material, material, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(g.albedo, albedo, n * 16, cudaMemcpyHostToDevice, c->stream));
    int rc = vpt_denoise(c, p, cam, prevCam, frameNum, iterationIndex);
    if (rc) return rc;
    CU(cudaMemcpyAsync(outputRGBA, c->illumOutput, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}

int vpt_tonemap(vpt_ctx *c, const VptToneMappingParams *p, uint8_t *rgb8, float *rgbaLDR)
{
    if (!c || !p) return fail(VPT_ERR_ARG, "vpt_tonemap: null argument");
    CU(cudaSetDevice(c->device));
    const size_t n = c->npix();
    // ping is free between frames (it is rewritten by the next denoise): LDR float plane; the bytes go after it
    CU(waitPendingCopy(c)); // the LDR plane lives in ping
    uint8_t *d8 = nullptr;
    if (rgb8)
    {
        if (!c->rgb8) CU(cudaMalloc((void **)&c->rgb8, n * 3)); // allocated once, at the first 8-bit read-out
        d8 = c->rgb8;
    }
    CU(launchTonemap(c->illumOutput, c->width, c->height, *p, d8, rgbaLDR ? c->ping : nullptr, c->stream));
    if (rgb8) CU(cudaMemcpyAsync(rgb8, d8, n * 3, cudaMemcpyDeviceToHost, c->stream));
    if (rgbaLDR) CU(cudaMemcpyAsync(rgbaLDR, c->ping, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}

int vpt_get_counters(vpt_ctx *c, uint64_t *rays, uint64_t *steps)
{
    if (!c || !rays || !steps) return fail(VPT_ERR_ARG, "vpt_get_counters: null argument");
    unsigned long long h[2];
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(h, c->counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *rays = h[0]; *steps = h[1];
    return VPT_OK;
}

/* debug: raw read of a wavefront-state plane of the last wave (which: 0 candC, 1 ris, 2 rstA, 3 rstB, 4 lightA, 5 light2A); 16 bytes per entry */
int vpt_debug_read_wave(vpt_ctx *c, int which, void *host, size_t entries)
{
    if (!c || !host || !c->wave.arena) return fail(VPT_ERR_STATE, "vpt_debug_read_wave: no wave state");
    const void *src[] = {c->wave.wb.candC, c->wave.wb.ris, c->wave.wb.rstA, c->wave.wb.rstB, c->wave.wb.lightA, c->wave.wb.light2A};
    if (which < 0 || which > 5) return fail(VPT_ERR_ARG, "vpt_debug_read_wave: unknown plane");
    if (entries > (size_t)c->wave.nSlots * (size_t)c->wave.maxSamplesInWave) return fail(VPT_ERR_ARG, "vpt_debug_read_wave: more entries than the wave has paths");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(host, src[which], entries * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VPT_OK;
}

/* debug: tile loads (TMA) of the denoiser whose completion barrier timed out since the process started; 0 in a healthy build */
int vpt_debug_tma_timeouts(void) { return (int)debugTmaTimeouts(); }

int vpt_get_lights(vpt_ctx *c, VptLightInfo *lights, VptAliasBin *alias, uint32_t *faceKeys, int capacity)
{
    if (!c) { fail(VPT_ERR_ARG, "vpt_get_lights: null context"); return -1; }
    if (cudaSetDevice(c->device) != cudaSuccess) { fail(VPT_ERR_CUDA, "vpt_get_lights: cudaSetDevice"); return -1; }
    if (refreshLights(c) != VPT_OK) return -1;
    const int n = (int)c->lights.lights.size();
    if (n > capacity && (lights || alias || faceKeys)) { fail(VPT_ERR_ARG, "vpt_get_lights: capacity too small"); return -1; }
    if (lights) std::memcpy(lights, c->lights.lights.data(), (size_t)n * sizeof(VptLightInfo));
    if (alias) std::memcpy(alias, c->lights.alias.data(), (size_t)n * sizeof(VptAliasBin));
    if (faceKeys) std::memcpy(faceKeys, c->lights.faceKeys.data(), c->lights.faceKeys.size() * sizeof(uint32_t));
    return n;
}

int vpt_get_total_rays(vpt_ctx *c, uint64_t *rays, int reset)
{
    if (!c || !rays) return fail(VPT_ERR_ARG, "vpt_get_total_rays: null argument");
    unsigned long long h = 0;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(&h, c->counters + 2, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    if (reset) CU(cudaMemsetAsync(c->counters + 2, 0, sizeof h, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *rays = h;
    return VPT_OK;
}

int vpt_get_timings(vpt_ctx *c, VptTimings *t)
{
    if (!c || !t) return fail(VPT_ERR_ARG, "vpt_get_timings: null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    std::memset(t, 0, sizeof *t);
    auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]); return ms; };
    if (c->haveTrace)
    {
        t->trace_ms = el(EV_TRACE0, EV_TRACE1);
        if (c->ranResolve) t->resolve_ms = el(EV_TRACE1, EV_RESOLVE1);
    }
    if (c->haveDenoise)
    {
        t->firefly_ms = el(EV_DN0, EV_FIREFLY);
        t->composite_ms = el(EV_FIREFLY, EV_SKY) + el(EV_ATROUS, EV_COMP);
        t->temporal_ms = el(EV_SKY, EV_TEMPORAL);
        t->history_fix_ms = el(EV_TEMPORAL, EV_HFIX);
        t->history_clamp_ms = el(EV_HFIX, EV_HCLAMP);
        t->atrous_smem_ms = el(EV_HCLAMP, EV_ASMEM);
        t->atrous_ms = el(EV_ASMEM, EV_ATROUS);
        t->denoise_total_ms = el(EV_DN0, EV_COMP);
        t->atrous_passes = c->atrousPasses;
    }
    if (c->haveTrace)
    {
        for (int i = 0; i < c->traceProf.n; ++i)
        {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, c->traceProf.ev[i], c->traceProf.ev[i + 1]);
            if (c->traceProf.kind[i] == 0) { t->trace_dda_ms += ms; t->trace_dda_launches++; }
            else { t->trace_shade_ms += ms; t->trace_shade_launches++; }
        }
    }
    t->kernel_launches = c->launchesRender + c->launchesDenoise;
    return VPT_OK;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------ multi-GPU (NCCL, dlopen'ed)
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi
{
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
bool loadNccl()
{
    if (g_nccl.handle) return true;
    // Prefer a libnccl already mapped into the process (torch's bundled 2.28) so that only one NCCL lives here.
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_lastError = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { g_lastError = std::string("libnccl: missing ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(AllReduce, "ncclAllReduce") SYM(Broadcast, "ncclBroadcast")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString") SYM(CommDestroy, "ncclCommAbort")
#undef SYM
    g_nccl.handle = h;
    return true;
}
} // namespace
static void destroyComm(vpt_ctx *c)
{
    // ncclCommAbort: frees the communicator without any handshake with the peers (the context's streams are already drained), so a rank
    // that tears down while its peers are still reporting — or have died — can never block here
    if (c->ncclComm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->ncclComm);
    c->ncclComm = nullptr;
}
#define NC(call)                                                                                                    \
    do {                                                                                                            \
        ncclResult_t r_ = (call);                                                                                   \
        if (r_ != ncclSuccess) return fail(VPT_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_));    \
    } while (0)

extern "C" {

int vpt_comm_unique_id(uint8_t *id128)
{
    if (!id128) return fail(VPT_ERR_ARG, "null id");
    if (!loadNccl()) return VPT_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, &id, 128);
    return VPT_OK;
}
int vpt_comm_init(vpt_ctx *c, int rank, int nranks, const uint8_t *id128)
{
    if (!c || !id128 || rank < 0 || rank >= nranks) return fail(VPT_ERR_ARG, "vpt_comm_init: bad argument");
    if (!loadNccl()) return VPT_ERR_NCCL;
    CU(cudaSetDevice(c->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t comm;
    NC(g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->ncclComm = comm; c->rank = rank; c->nranks = nranks;
    return VPT_OK;
}
int vpt_comm_allreduce_illumination(vpt_ctx *c)
{
    if (!c || !c->ncclComm) return fail(VPT_ERR_STATE, "vpt_comm_allreduce_illumination: communicator not initialised");
    CU(cudaSetDevice(c->device));
    NC(g_nccl.AllReduce(c->illumination, c->illumination, c->npix() * 4, ncclFloat, ncclSum, (ncclComm_t)c->ncclComm, c->stream));
    return VPT_OK;
}
int vpt_comm_broadcast_gbuffer(vpt_ctx *c, int iterationIndex)
{
    if (!c || !c->ncclComm) return fail(VPT_ERR_STATE, "vpt_comm_broadcast_gbuffer: communicator not initialised");
    CU(cudaSetDevice(c->device));
    const size_t n = c->npix();
    GSet &g = c->gb[c->cur];
    ncclComm_t comm = (ncclComm_t)c->ncclComm;
    NC(g_nccl.GroupStart());
    NC(g_nccl.Broadcast(g.depth, g.depth, n, ncclFloat, 0, comm, c->stream));
    NC(g_nccl.Broadcast(g.material, g.material, n, ncclFloat, 0, comm, c->stream));
    NC(g_nccl.Broadcast(g.normalRoughness, g.normalRoughness, n * 4, ncclFloat, 0, comm, c->stream));
    NC(g_nccl.Broadcast(g.geoNormalThinfilm, g.geoNormalThinfilm, n * 4, ncclFloat, 0, comm, c->stream));
    NC(g_nccl.Broadcast(g.materialParameter, g.materialParameter, n * 4, ncclFloat, 0, comm, c->stream));
    NC(g_nccl.Broadcast(g.albedo, g.albedo, n * 4, ncclFloat, 0, comm, c->stream));
    VptReservoir *r = c->reservoirs + (size_t)(iterationIndex & 1) * n;
    NC(g_nccl.Broadcast(r, r, n * 5, ncclFloat, 0, comm, c->stream));
    NC(g_nccl.GroupEnd());
    return VPT_OK;
}

} // extern "C"

// ---- row-band sharded Denoiser::run
// Every rank runs the WHOLE chain, without any exchange, on its band extended by E rows either side, E = the depth of the chain's
// dependency cone: HistoryFix 2 x 9 taps, HistoryClamping 5 x 5, the first a-trous pass (halo 2) and every a-trous step (step +
// hashed jitter), plus HitDistReconstruction (2) / PrePass (31) when enabled. Rows of the extension are computed with a truncated
// cone and are simply not used: the band proper is exact — bit-identical to the single-GPU chain. The only data a rank lacks for the
// NEXT frame is the history (PrevIllumination, PrevFastIllumination, PrevHistoryLength) outside the rows where its own copy is
// exact (band +- 20: HistoryFix + HistoryClamping): one grouped exchange per frame with the two neighbours brings in the rows
// [band - E - 32, band - 20) and [band + 20, band + E + 32) — the extension plus the 32-row guard of the temporal pass's reprojected
// taps. One NCCL launch per frame instead of the 13 per-pass halo exchanges of the first version (which were launch-latency bound:
// a 4K band at 8 GPUs is ~0.15 ms of kernels).
static int chainDepthRows(const VptDenoisingParams *p)
{
    int d = 0;
    if (p->enableHitDistanceReconstruction) d += 2;
    if (p->enablePrePass) d += 31;
    if (p->enableTemporalAccumulation) d += (p->enableHistoryFix ? 18 : 0) + (p->enableHistoryClamping ? 2 : 0);
    if (p->enableSpatialFiltering)
    {
        d += 2;
        // the chain runs steps 2^1 .. 2^(2n+1) for atrousIterationNum = n > 0 (Denoiser.cu:296-352)
        for (int idx = 1; p->atrousIterationNum > 0 && idx <= 2 * p->atrousIterationNum + 1; ++idx) { const int step = 1 << idx; d += step + (step > 4 ? step / 4 : 0); }
    }
    return (d + 3) & ~3; // band boundaries are multiples of 4 rows (firefly tiles): so is the extension
}
constexpr int kHistoryExact = 20; // a rank's own history planes are exact on its band +- 20 rows
constexpr int kTemporalGuard = 32;

extern "C" void vpt_band_rows(int height, int nranks, int rank, int *rowBegin, int *rowEnd)
{
    // python/vpt_shard.row_bands: boundaries at multiples of 4 rows
    auto bound = [&](int r) { return r >= nranks ? height : ((int)((long long)height * r / nranks) / 4) * 4; };
    if (rowBegin) *rowBegin = bound(rank);
    if (rowEnd) *rowEnd = bound(rank + 1);
}
extern "C" int vpt_band_input_halo(const VptDenoisingParams *p) { return p ? chainDepthRows(p) + kTemporalGuard : -1; }

// history planes of the next frame: rows [b0 - E - guard, b0 - exact) from the band above, [b1 + exact, b1 + E + guard) from below
static int historyExchange(vpt_ctx *c, int b0, int b1, int E)
{
    if (c->nranks == 1) return VPT_OK;
    ncclComm_t comm = (ncclComm_t)c->ncclComm;
    const int H = c->height, W = c->width;
    struct Plane { void *ptr; int floats; } planes[3] = {{c->prevIllum, 4}, {c->prevFastIllum, 4}, {c->prevHistoryLength, 1}};
    const int up = c->rank - 1, down = c->rank + 1;
    NC(g_nccl.GroupStart());
    for (const Plane &pl : planes)
    {
        float *base = (float *)pl.ptr;
        const size_t rowFloats = (size_t)W * pl.floats;
        if (up >= 0)
        {
            // the band above needs [b0 + exact, b0 + E + guard) of mine; I need [b0 - E - guard, b0 - exact) of its rows
            const int s0 = std::min(H, b0 + kHistoryExact), s1 = std::min(H, b0 + E + kTemporalGuard);
            const int r0 = std::max(0, b0 - E - kTemporalGuard), r1 = std::max(0, b0 - kHistoryExact);
            if (s1 > s0) NC(g_nccl.Send(base + (size_t)s0 * rowFloats, (size_t)(s1 - s0) * rowFloats, ncclFloat, up, comm, c->stream));
            if (r1 > r0) NC(g_nccl.Recv(base + (size_t)r0 * rowFloats, (size_t)(r1 - r0) * rowFloats, ncclFloat, up, comm, c->stream));
        }
        if (down < c->nranks)
        {
            const int s0 = std::max(0, b1 - E - kTemporalGuard), s1 = std::max(0, b1 - kHistoryExact);
            const int r0 = std::min(H, b1 + kHistoryExact), r1 = std::min(H, b1 + E + kTemporalGuard);
            if (s1 > s0) NC(g_nccl.Send(base + (size_t)s0 * rowFloats, (size_t)(s1 - s0) * rowFloats, ncclFloat, down, comm, c->stream));
            if (r1 > r0) NC(g_nccl.Recv(base + (size_t)r0 * rowFloats, (size_t)(r1 - r0) * rowFloats, ncclFloat, down, comm, c->stream));
        }
    }
    NC(g_nccl.GroupEnd());
    return VPT_OK;
}

// largest reprojection shift (pixel rows / columns) the camera pair can produce for geometry at >= 4 units: rotation + translation
static float reprojectionBound(const VptCamera *cam, const VptCamera *prev, int height)
{
    const float d = cam->dir[0] * prev->dir[0] + cam->dir[1] * prev->dir[1] + cam->dir[2] * prev->dir[2];
    const float angle = std::acos(std::fmin(1.0f, std::fmax(-1.0f, d)));
    const float dx = cam->pos[0] - prev->pos[0], dy = cam->pos[1] - prev->pos[1], dz = cam->pos[2] - prev->pos[2];
    const float move = std::sqrt(dx * dx + dy * dy + dz * dz);
    const float pixelsPerRadian = (float)height / (2.0f * std::atan(cam->tanHalfFov[1]));
    return (angle + std::atan(move / 4.0f)) * pixelsPerRadian;
}

extern "C" {

// Row-band sharded Denoiser::run: every rank holds full-size planes; rank r owns the rows [rowBegin,rowEnd) (vpt_band_rows)
// and leaves its band of IlluminationOutput exact. Inputs (Illumination + the current G-buffer) must be valid on the band +-
// vpt_band_input_halo(params) rows (the spp-sharded renderer leaves them valid everywhere).
int vpt_denoise_band(vpt_ctx *c, const VptDenoisingParams *p, const VptCamera *cam, const VptCamera *prevCam, int frameNum, int iterationIndex,
                     int rowBegin, int rowEnd)
{
    if (!c || !p || !cam || !prevCam) return fail(VPT_ERR_ARG, "vpt_denoise_band: null argument");
    if (rowBegin < 0 || rowEnd > c->height || rowBegin >= rowEnd || (rowBegin & 3)) return fail(VPT_ERR_ARG, "vpt_denoise_band: band must start on a multiple of 4 rows");
    if (c->nranks > 1 && !c->ncclComm) return fail(VPT_ERR_STATE, "vpt_denoise_band: communicator not initialised");
    const int E = chainDepthRows(p);
    if (c->nranks > 1)
    {
        // one-hop exchange: what a neighbour sends must lie in the rows where its own history is exact (its band +- 20)
        int smallest = c->height;
        for (int r = 0; r < c->nranks; ++r) { int a, b; vpt_band_rows(c->height, c->nranks, r, &a, &b); smallest = std::min(smallest, b - a); }
        if (E + kTemporalGuard - kHistoryExact > smallest)
            return fail(VPT_ERR_ARG, "vpt_denoise_band: the settings need " + std::to_string(E + kTemporalGuard - kHistoryExact) + " history rows from a neighbour but the smallest band has " +
                                         std::to_string(smallest) + " rows (lower atrousIterationNum or use fewer ranks)");
        if (p->enableTemporalAccumulation && frameNum > 0 && reprojectionBound(cam, prevCam, c->height) > (float)kTemporalGuard)
            return fail(VPT_ERR_ARG, "vpt_denoise_band: the camera moved too far for the 32-row history guard of the band-sharded temporal pass");
    }
    CU(cudaSetDevice(c->device));
    CU(waitPendingCopy(c));
    const bool sharded = c->nranks > 1;
    const int e0 = sharded ? std::max(0, rowBegin - E) : rowBegin, e1 = sharded ? std::min(c->height, rowEnd + E) : rowEnd;
    const int rc = denoiseChain(c, p, cam, prevCam, frameNum, iterationIndex, e0, e1, c->profiling, sharded); // (per-pass times: of the extended band)
    if (rc) return rc;
    return sharded ? historyExchange(c, rowBegin, rowEnd, E) : VPT_OK;
}

// The bands of IlluminationOutput collected on rank `root` (one grouped send / receive): the frame a caller reads back.
int vpt_comm_gather_output(vpt_ctx *c, int root)
{
    if (!c || !c->ncclComm) return fail(VPT_ERR_STATE, "vpt_comm_gather_output: communicator not initialised");
    if (root < 0 || root >= c->nranks) return fail(VPT_ERR_ARG, "vpt_comm_gather_output: bad root");
    CU(cudaSetDevice(c->device));
    ncclComm_t comm = (ncclComm_t)c->ncclComm;
    float *base = (float *)c->illumOutput;
    const size_t rowFloats = (size_t)c->width * 4;
    NC(g_nccl.GroupStart());
    if (c->rank == root)
    {
        for (int r = 0; r < c->nranks; ++r)
        {
            if (r == root) continue;
            int a, b; vpt_band_rows(c->height, c->nranks, r, &a, &b);
            NC(g_nccl.Recv(base + (size_t)a * rowFloats, (size_t)(b - a) * rowFloats, ncclFloat, r, comm, c->stream));
        }
    }
    else
    {
        int a, b; vpt_band_rows(c->height, c->nranks, c->rank, &a, &b);
        NC(g_nccl.Send(base + (size_t)a * rowFloats, (size_t)(b - a) * rowFloats, ncclFloat, root, comm, c->stream));
    }
    NC(g_nccl.GroupEnd());
    return VPT_OK;
}

} // extern "C"
