/*
 * vpt.h — C ABI of libvpt.so: the B200-native (sm_100a) voxel path-tracing + denoising hot path.
 *
 * The reference (wangkepfe/Real-time-path-tracing-voxel-blocks) has no FFI; its hot path sits behind C++
 * singletons called from one host thread. Each entry point below names the reference interface it replaces
 * (paths relative to /root/reference). Conventions: opaque handle, int status (0 = ok, non-zero =
 * error, text via vpt_last_error()), no exceptions across the boundary, caller-owned host memory, library-owned
 * device memory, one context per GPU, calls on a context serialised by the caller. There is NO CPU fallback:
 * every compute entry point fails with VPT_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef VPT_H
#define VPT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPT_OK 0
#define VPT_ERR_ARG 1
#define VPT_ERR_CUDA 2
#define VPT_ERR_STATE 3
#define VPT_ERR_IO 4
#define VPT_ERR_NCCL 5

typedef struct vpt_ctx vpt_ctx;

/* renderer/shaders/Camera.h:6-28 — same field order, 212 bytes. Matrices are 3x3, column storage
 * (m00,m10,m20, m01,m11,m21, m02,m12,m22) as renderer/shaders/LinearMath.h:1040-1050. */
typedef struct VptCamera
{
    float resolution[2];
    float inversedResolution[2];
    float tanHalfFov[2];
    float pos[3];
    float dir[3];
    float posDelta[3];
    float yaw;
    float pitch;
    float uvToWorld[9];
    float worldToUv[9];
    float uvToView[9];
    float viewToUv[9];
} VptCamera;

/* renderer/shaders/SystemParameter.h:11-38 (MaterialParameter) without the texture handles. */
typedef struct VptMaterial
{
    float albedo[3];
    float roughness;
    float translucency;
    float uvScale;
    int32_t metallic;
    int32_t materialId;
    int32_t useWorldGridUV;
    int32_t isEmissive;
    int32_t isThinfilm;
    int32_t pad;
} VptMaterial;

/* renderer/shaders/Light.h:13-24 (LightInfo): a packed triangle light — centroid, fp16 edge lengths, fp16 radiance, oct-encoded
 * unit edge directions (TriangleLight::Store / Create, Light.h:84-136). */
typedef struct VptLightInfo
{
    float center[3];
    uint32_t scalars;
    uint32_t radiance[2];
    uint32_t direction1;
    uint32_t direction2;
} VptLightInfo;

/* renderer/shaders/AliasTable.h:11-16 */
typedef struct VptAliasBin
{
    float q;
    float p;
    int32_t alias;
} VptAliasBin;

/* renderer/shaders/RestirCommon.h:6-13 (DIReservoir, 20 bytes) */
typedef struct VptReservoir
{
    uint32_t lightData;
    uint32_t uvData;
    float weightSum;
    float targetPdf;
    float M;
} VptReservoir;

/* renderer/core/GlobalSettings.h:82-141 (DenoisingParams) */
typedef struct VptDenoisingParams
{
    int32_t enableHitDistanceReconstruction;
    int32_t enablePrePass;
    int32_t enableTemporalAccumulation;
    int32_t enableHistoryFix;
    int32_t enableHistoryClamping;
    int32_t enableSpatialFiltering;
    int32_t enableFireflyFilter;
    float maxAccumulatedFrameNum;
    float maxFastAccumulatedFrameNum;
    float phiLuminance;
    float lobeAngleFraction;
    float roughnessFraction;
    float depthThreshold;
    int32_t atrousIterationNum;
    float disocclusionThreshold;
    float disocclusionThresholdAlternate;
    float denoisingRange;
} VptDenoisingParams;

/* Subset of renderer/core/BufferManager.h:8-46 (Buffer2DName). F4 = float4 per pixel, F1 = float per pixel. */
typedef enum VptBufferName
{
    VPT_BUF_Illumination = 0,          /* F4 noisy radiance rgb + primary hit distance (RayGen.cu:181) */
    VPT_BUF_IlluminationOutput = 1,    /* F4 denoised, albedo re-applied (BufferCopy.h:36-116) */
    VPT_BUF_IlluminationPing = 2,      /* F4 */
    VPT_BUF_IlluminationPong = 3,      /* F4 */
    VPT_BUF_NormalRoughness = 4,       /* F4 */
    VPT_BUF_Depth = 5,                 /* F1 */
    VPT_BUF_Material = 6,              /* F1 */
    VPT_BUF_Albedo = 7,                /* F4 */
    VPT_BUF_HistoryLength = 8,         /* F1 */
    VPT_BUF_PrevDepth = 9,             /* F1 */
    VPT_BUF_PrevMaterial = 10,         /* F1 */
    VPT_BUF_PrevIllumination = 11,     /* F4 */
    VPT_BUF_PrevFastIllumination = 12, /* F4 */
    VPT_BUF_PrevHistoryLength = 13,    /* F1 */
    VPT_BUF_PrevNormalRoughness = 14,  /* F4 */
    VPT_BUF_GeoNormalThinfilm = 15,    /* F4 */
    VPT_BUF_MaterialParameter = 16,    /* F4 */
    VPT_BUF_PrevMaterialParameter = 17,/* F4 */
    VPT_BUF_PrevGeoNormalThinfilm = 18,/* F4 */
    VPT_BUF_PrevAlbedo = 19,           /* F4 */
    VPT_BUF_PrimaryHits = 22           /* int4 per pixel: voxel x,y,z and face id (-1 = miss); new in this build */
} VptBufferName;

/* Per-stage device timings of the last vpt_render / vpt_denoise call, CUDA events on the context stream (ms). */
typedef struct VptTimings
{
    float trace_ms;
    float resolve_ms;
    float firefly_ms;
    float temporal_ms;
    float history_fix_ms;
    float history_clamp_ms;
    float atrous_smem_ms;
    float atrous_ms;        /* sum over the Atrous passes */
    float composite_ms;     /* BufferCopySky + BufferCopyNonSky */
    float denoise_total_ms; /* whole chain, first launch to last */
    int32_t atrous_passes;
    int32_t kernel_launches; /* kernels launched by the last render + denoise calls */
    float trace_dda_ms;     /* sum over the DDA (traversal) kernels of the last render */
    float trace_shade_ms;   /* sum over the raygen / shading-stage / accumulate kernels of the last render */
    int32_t trace_dda_launches;
    int32_t trace_shade_launches;
} VptTimings;

/* ---- lifetime. Replaces OfflineBackend::init + BufferManager::init + OptixRenderer::init
 * (renderer/core/OfflineBackend.cpp:26-44, BufferManager.cpp:107-241, OptixRenderer.cpp:928-1366). */
int vpt_create(int device, int width, int height, vpt_ctx **out);
void vpt_destroy(vpt_ctx *ctx);
const char *vpt_last_error(void);
int vpt_sync(vpt_ctx *ctx);
/* cudaStream_t of the context (OfflineBackend::getCudaStream, OfflineBackend.h:41) as an opaque pointer. */
void *vpt_stream(vpt_ctx *ctx);

/* ---- inputs */
/* util/RandGenHost.cpp:5-20 (BlueNoiseRandGeneratorHost::init): sobol[65536], scrambling[131072], ranking[131072]. */
int vpt_set_tables(vpt_ctx *ctx, const uint8_t *sobol, const uint8_t *scrambling, const uint8_t *ranking);
/* VoxelEngine::voxelChunks (voxelengine/VoxelEngine.h:25-76): ids are chunk-major, chunk index
 * cx + CX*(cz + CZ*cy) (VoxelEngine.cu:221-224), 32768 bytes per chunk in GetLinearId order (VoxelMath.h:120-127). */
int vpt_set_grid(vpt_ctx *ctx, int chunksX, int chunksY, int chunksZ, const uint8_t *ids);
int vpt_get_grid(vpt_ctx *ctx, uint8_t *ids_out, size_t bytes);
/* VoxelEngine::setVoxelAtGlobal (VoxelEngine.cu:265-276). 
 * The first edit after a render keeps a snapshot of the traversal masks: that frame's temporal-ReSTIR bias rays walk the world as
 * the previous render saw it (sysParam.prevTopObject, closesthit.cu:736-755). */
int vpt_set_voxel(vpt_ctx *ctx, int x, int y, int z, int blockId);
/* initVoxelsMultiChunk + GenerateVoxelChunk (voxelengine/VoxelSceneGen.cu:341-388, 61-165): noise is
 * chunks x 32 x 32 floats, noise[chunk][z][x]. Runs on the device. */
int vpt_generate_terrain(vpt_ctx *ctx, int chunksX, int chunksY, int chunksZ, const float *noise);
/* The block picker (VoxelEngine::performRayTraversal, voxelengine/VoxelEngine.cu:1040-1166; result type
 * VoxelEngine.h:85-93): walks the resident grid from `origin` along `direction` (the reference passes camera.pos / camera.dir)
 * with the reference's comparison order and tMax += tDelta accumulation, at most 1000 voxels, stopping when it leaves the
 * grid (an origin outside the grid finds nothing). deletePos / deleteBlockId = first solid voxel; createPos = the last empty
 * voxel before it. VoxelEngine::update (:906-985) then calls deleteBlock(deletePos) or addBlock(createPos, id): vpt_set_voxel. */
typedef struct VptPickResult
{
    int32_t hasSpaceToCreate;
    int32_t hitSurface;
    int32_t createPos[3];
    int32_t deletePos[3];
    int32_t deleteBlockId;
} VptPickResult;
int vpt_pick_voxel(vpt_ctx *ctx, const float *origin /*[3]*/, const float *direction /*[3]*/, VptPickResult *out);

/* MaterialManager (renderer/assets/MaterialManager.cpp:85-120): material table + block id -> material index. */
int vpt_set_materials(vpt_ctx *ctx, const VptMaterial *materials, int count, const uint16_t *blockToMaterial /*[256]*/);
/* SkyModel buffers (renderer/sky/Sky.cu:355-396): RGBA32F sky (equal-area sphere map) and sun-disk maps with
 * their alias tables (shaders/AliasTable.cu:66-153) and the sun direction. */
int vpt_set_sky(vpt_ctx *ctx, const float *skyRGBA, int skyW, int skyH, const float *sunRGBA, int sunW, int sunH,
                const VptAliasBin *skyAlias, const VptAliasBin *sunAlias, const float *sunDir /*[3]*/);
/* Textured materials (SURVEY 8a S5 / 8f "next" row 3; closesthit.cu:166-254, sampler state TextureManager.cu:222-240: wrap
 * addressing, trilinear, normalised coordinates, no sRGB, mip levels down to 4x4). The reference feeds NVTT-encoded BC7/BC5/BC4
 * blocks to the texture unit — neither is reproducible — so this build filters UNCOMPRESSED RGBA8 mip chains in software
 * (fp32): nTextures square power-of-two textures, texture t has levels[t] levels of (widths[t] >> l)^2 RGBA8 texels (r = low
 * byte), all levels of all textures concatenated in `texels`. slots4[m] = albedo, normal, roughness, metallic texture index
 * of material m (-1 = none; single-channel maps are read from .r); texSize2[m] = MaterialParameter::texSize. nTextures = 0
 * removes them. Must follow vpt_set_materials (same material count). */
int vpt_set_textures(vpt_ctx *ctx, int nTextures, const int32_t *widths, const int32_t *levels, const uint32_t *texels, int nMaterials,
                     const int32_t *slots4, const float *texSize2);
/* Asset tables: materials.yaml -> MaterialParameter rows in file order (materialId = index, MaterialManager.cpp:85-100;
 * defaults MaterialDefinition.h:18-28; metallic float -> bool; emissive materials carry emissive_radiance in albedo, :151-167) and
 * blocks.yaml -> block id -> material index (MaterialManager.cpp:104-120; unmapped -> 0). paths (may be NULL) receives the texture
 * files of each material as the reference resolves them ("data/" + the YAML value, AssetRegistry.cpp:86-97), "" = none.
 * blocksYamlPath / blockToMaterial256 may be NULL. Host only. */
typedef struct VptMaterialTexturePaths
{
    char albedo[256];
    char normal[256];
    char roughness[256];
    char metallic[256];
} VptMaterialTexturePaths;
int vpt_load_materials(const char *materialsYamlPath, const char *blocksYamlPath, VptMaterial *materials, VptMaterialTexturePaths *paths,
                       int maxMaterials, int *count, uint16_t *blockToMaterial256);
/* PNG decode to RGBA8 words (r = low byte), what stbi_load gives TextureManager::init (TextureManager.cu:178-200): 8/16-bit grey,
 * grey+alpha, RGB, palette, RGBA; non-interlaced. Grey is replicated to rgb, missing alpha = 255. out == NULL only queries
 * width/height/channels (channels = the file's channel count, as stb reports it). Host only. */
int vpt_load_png_rgba8(const char *path, uint32_t *out, size_t maxTexels, int *width, int *height, int *channels);
/* The reference's image-diff tool (renderer/util/ImageDiff.cpp:95-372; thresholds :119-121; docs/image-diffing-system.md), which
 * mainOffline --test-canonical applies to its last frame (mainOffline.cpp:422-497): pixels whose per-channel difference exceeds
 * 0.01 of full scale, RMSE over all samples, global SSIM of the 3x3-gaussian filtered luma (K1 0.01, K2 0.03, L 255);
 * IDENTICAL = no different pixel, VERY CLOSE = SSIM > 0.99 and RMSE < 1, CLOSE = SSIM > 0.95 and RMSE < 5. The images are RGBA8
 * words (vpt_load_png_rgba8 layout); `channels` = how many of r,g,b,a take part (min of the two files' channel counts).
 * Sums are taken in double. Host only. Returns VPT_ERR_ARG for images of different size. */
typedef struct VptImageDiffResult
{
    int32_t differentPixels;
    int32_t totalPixels;
    float pixelDifferenceRatio;
    float rmse;
    float ssim;
    int32_t isIdentical;
    int32_t isVeryClose;
    int32_t isClose;
} VptImageDiffResult;
int vpt_image_diff(const uint32_t *imageA, const uint32_t *imageB, int width, int height, int channels, VptImageDiffResult *result);
/* ImageDiff::compare(pathA, pathB) on two PNG files; diffPngOrNull != NULL also writes the difference picture of
 * ImageDiff::generateDiffImage (:138-190: |a - b| per channel, amplified 3x, capped at 255, RGB). VPT_ERR_IO if a file cannot be read. */
int vpt_image_diff_files(const char *pngA, const char *pngB, VptImageDiffResult *result, const char *diffPngOrNull);
/* Mip chain of one square power-of-two RGBA8 image the way TextureManager::init builds it (renderer/assets/TextureManager.cu:
 * 82-115, 216-217, 395-411): level l+1 = per-channel 2x2 box average of level l, truncated to 8 bits; levels stop at 4x4
 * (numLods = log2(width) - 1; images smaller than 4x4 keep 1 level). Host only. out receives all levels concatenated (the layout
 * vpt_set_textures takes) and must hold vpt_mip_chain_texels(width) texels (0 for an invalid width). Returns the number of levels, 0 on a bad argument. */
int vpt_mip_chain_texels(int width);
int vpt_build_mip_chain(const uint32_t *level0, int width, uint32_t *out);

/* SkyParams (renderer/core/GlobalSettings.h:188-204) */
typedef struct VptSkyParams
{
    float timeOfDay;     /* 0.25 */
    float sunAxisAngle;  /* 45 */
    float sunAxisRotate; /* 0 */
    float skyBrightness; /* 1 */
} VptSkyParams;
/* SkyModel::update (renderer/sky/Sky.cu:355-396): evaluates the Hosek-Wilkie sky (1024 x 512 equal-area sphere map) and the
 * solar disc (32 x 32 cone map) ON THE DEVICE from the parameters, builds both alias tables (CPU build path like the
 * reference, shaders/AliasTable.cu:66-153) and installs them like vpt_set_sky. tables = the 2460 floats of
 * data/sky_tables.bin (the coefficient datasets of renderer/sky/SkyData.h). */
int vpt_generate_sky(vpt_ctx *ctx, const VptSkyParams *params, const float *tables);
/* Current sky state: RGBA32F maps, sun direction (any pointer may be NULL); sizes via vpt_sky_size. */
int vpt_read_sky(vpt_ctx *ctx, float *skyRGBA, float *sunRGBA, float *sunDir3);
int vpt_sky_size(vpt_ctx *ctx, int *skyW, int *skyH, int *sunW, int *sunH);
/* Host part of SkyModel::update: sun direction (Sky.cu:362-367) and updateSkyState (Sky.cu:52-79). */
void vpt_sky_state(const VptSkyParams *params, const float *tables, float *configs90, float *radiances10, float *sunDir3);
/* New parameters of this build (SURVEY "five facts" #2); the reference is spp=1, limits 3/1, ReSTIR on
 * (renderer/shaders/RayGen.cu:146-147). */
int vpt_set_trace_params(vpt_ctx *ctx, int spp, int totalBounceLimit, int diffuseBounceLimit, int enableRestir);

/* Wavefront sizing: the renderer processes at most `maxPaths` (pixel slots x samples) per wave; more samples run wave by
 * wave (default 16 Mi paths ~ 4.9 GB of path state). A tuning / test knob: results do not depend on it. */
int vpt_set_wave_budget(vpt_ctx *ctx, size_t maxPaths);

/* ---- the hot path */
/* OptixRenderer::render (renderer/core/OptixRenderer.cpp:411-485) == __raygen__pathtracer over W x H
 * (renderer/shaders/RayGen.cu:102-181). iterationIndex is the pre-increment value the reference stores in
 * sysParam (OptixRenderer.cpp:419-420). Asynchronous on the context stream; vpt_sync/vpt_read_buffer wait. */
int vpt_render(vpt_ctx *ctx, const VptCamera *camera, const VptCamera *prevCamera, int iterationIndex);
/* Sample-sharded form (multi-GPU): renders samples k = sampleBegin, sampleBegin+sampleStep, ... < spp and leaves
 * the un-normalised radiance SUM in Illumination; vpt_resolve divides by spp (after the cross-GPU sum). */
int vpt_render_shard(vpt_ctx *ctx, const VptCamera *camera, const VptCamera *prevCamera, int iterationIndex,
                     int sampleBegin, int sampleStep);
/* Contiguous form: samples sampleBegin .. sampleBegin + sampleCount - 1 (clipped to spp; sampleCount 0 = an empty shard that
 * contributes zero). Lets a caller size the shards by COST: the rank that renders sample 0 also runs the temporal ReSTIR pass
 * (about 1.5 plain samples of extra work) and, in the offline flow, the denoiser — python/vpt_shard.py:balanced_ranges. */
int vpt_render_range(vpt_ctx *ctx, const VptCamera *camera, const VptCamera *prevCamera, int iterationIndex,
                     int sampleBegin, int sampleCount);
/* Same shard, but its FIRST sample (sampleBegin) owns this context's G-buffer, reservoir plane and temporal ReSTIR pass — every
 * rank then runs the reference's whole per-frame algorithm (RayGen.cu:102-181: one ReSTIR sample + spp-1 plain samples) on its own
 * sample subset with rank-local ReSTIR state ("keep it rank-local"), so the ranks do equal work. The sum over ranks is
 * an average of independent frames; it is NOT bit-comparable to one GPU rendering N x spp (vpt_render_shard is). */
int vpt_render_shard_local(vpt_ctx *ctx, const VptCamera *camera, const VptCamera *prevCamera, int iterationIndex,
                           int sampleBegin, int sampleStep);
int vpt_resolve(vpt_ctx *ctx);
/* Denoiser::run (renderer/denoising/Denoiser.cu:24-408). frameNum == OfflineBackend::getFrameNum();
 * iterationIndex is the POST-increment GlobalSettings::iterationIndex the reference reads there (:37,296). */
int vpt_denoise(vpt_ctx *ctx, const VptDenoisingParams *params, const VptCamera *camera, const VptCamera *prevCamera,
                int frameNum, int iterationIndex);
/* Denoiser on caller-supplied inputs (config 3): flips the G-buffer ping-pong like a render would; the caller
 * then uploads Illumination/Depth/NormalRoughness/Material/Albedo with vpt_write_buffer and calls vpt_denoise. */
int vpt_begin_external_frame(vpt_ctx *ctx);
/* One call = H2D of the five G-buffer planes + vpt_denoise + D2H of IlluminationOutput (the e2e path of cfg 3). */
int vpt_denoise_external(vpt_ctx *ctx, const VptDenoisingParams *params, const VptCamera *camera, const VptCamera *prevCamera,
                         int frameNum, int iterationIndex, const float *illumination, const float *depth,
                         const float *normalRoughness, const float *material, const float *albedo, float *outputRGBA);

/* ---- outputs. BufferManager::GetBuffer2D (renderer/core/BufferManager.h:84) + OfflineBackend::storeFrameInBatch. */
int vpt_read_buffer(vpt_ctx *ctx, VptBufferName name, void *host, size_t bytes);
int vpt_write_buffer(vpt_ctx *ctx, VptBufferName name, const void *host, size_t bytes);
/* The same from DEVICE memory of this GPU (a caller that owns a CUDA context in the process, e.g. a frame producer): an asynchronous
 * device-to-device copy ordered on the context's stream. No reference counterpart (its producers write the surfaces in place). */
int vpt_write_buffer_device(vpt_ctx *ctx, VptBufferName name, const void *device, size_t bytes);
/* Pipelined read-back (OfflineBackend::storeFrameInBatch without stalling the frame loop): the copy is queued on a copy
 * stream behind everything submitted so far and returns at once; `host` should be pinned. The next vpt_denoise /
 * vpt_render that would overwrite the plane waits for the copy on the DEVICE, so the transfer overlaps the next frame's
 * trace. vpt_sync (or vpt_read_wait) completes it. */
int vpt_read_buffer_async(vpt_ctx *ctx, VptBufferName name, void *host, size_t bytes);
int vpt_read_wait(vpt_ctx *ctx);
/* BufferManager::reservoirBuffer (BufferManager.cpp:206-207): plane `parity` of the 2 x W x H reservoir array. */
int vpt_read_reservoirs(vpt_ctx *ctx, int parity, VptReservoir *host, size_t bytes);
int vpt_write_reservoirs(vpt_ctx *ctx, int parity, const VptReservoir *host, size_t bytes);
/* Device pointer of a buffer (for callers that own a CUDA context in the same process, e.g. NCCL plumbing). */
void *vpt_device_ptr(vpt_ctx *ctx, VptBufferName name);
/* Device-side counters of the last render: traversal calls (rays) and voxel steps. */
int vpt_get_counters(vpt_ctx *ctx, uint64_t *rays, uint64_t *steps);
/* Traversal calls summed over every render since the context was created (or since the last call with reset != 0): lets a
 * frame loop be timed without a per-frame read-back. No reference counterpart (the reference counts nothing); the convention is
 * SURVEY 8d's: one per optixTraverse site reached (RayGen.cu:49, closesthit.cu:458, 616, 745, 801). */
int vpt_get_total_rays(vpt_ctx *ctx, uint64_t *rays, int reset);
/* Local emissive lights. Replaces, on this path, launchGenerateLightInfos + buildAliasTable (voxelengine/VoxelEngine.cu:53-192), which
 * the reference runs over the triangles of its emissive instanced meshes whenever the scene changes; here the lights are the exposed
 * faces of emissive voxels (two triangles per face; voxel order x + W*(z + D*y), face 0..5, triangle 0, 1) and the list is brought up to
 * date by the first vpt_render after vpt_set_grid / vpt_generate_terrain / vpt_set_voxel / vpt_set_materials. Returns the number of
 * lights (after refreshing the list), < 0 on error; lights / alias (count entries) and faceKeys (count / 2 entries:
 * (linear voxel << 3) | face) may be NULL or hold at least `capacity` lights. */
int vpt_get_lights(vpt_ctx *ctx, VptLightInfo *lights, VptAliasBin *alias, uint32_t *faceKeys, int capacity);
/* Debug counter: shared-memory tile loads of the denoiser whose completion barrier timed out (0 in a healthy build). */
int vpt_debug_tma_timeouts(void);
/* Debug: raw read-back of one 16-byte-per-path plane of the last wave's state (which: 0 candidate C, 1 RIS state, 2 / 3 stored
 * reservoir A / B, 4 light sample A, 5 second light sample A); entries <= the paths of a wave. tools/lights_debug.py. */
int vpt_debug_read_wave(vpt_ctx *ctx, int which, void *host, size_t entries);
int vpt_get_timings(vpt_ctx *ctx, VptTimings *out);
/* Toggle CUDA-event stage timing and the DDA step counter (default on; adds event records between kernels and one
 * add per DDA step). Throughput runs switch it off. */
int vpt_set_profiling(vpt_ctx *ctx, int enabled);

/* ---- multi-GPU: one context per rank, NCCL communicator owned by the library. */
/* 128-byte ncclUniqueId, created on rank 0 and broadcast by the caller's own plumbing. */
int vpt_comm_unique_id(uint8_t *id128);
int vpt_comm_init(vpt_ctx *ctx, int rank, int nranks, const uint8_t *id128);
/* ncclAllReduce(sum) of the W*H*4 float accumulation buffer (spp sharding), then every rank may vpt_resolve. */
int vpt_comm_allreduce_illumination(vpt_ctx *ctx);
/* Broadcast the sample-0 G-buffer, depth and current reservoir plane from rank 0 (so any rank can denoise). */
int vpt_comm_broadcast_gbuffer(vpt_ctx *ctx, int iterationIndex);
/* Row-band sharded Denoiser::run (the reference runs one GPU). Rank r of the communicator owns the rows vpt_band_rows gives
 * it (boundaries at multiples of 4: the firefly filter's 8x4 tiles) and leaves its band of IlluminationOutput exact — bit-identical
 * to the single-GPU chain. Every rank runs the whole chain, with no exchange, on its band extended by the depth of the chain's
 * dependency cone; one grouped ncclSend/ncclRecv per frame then brings in the history rows (PrevIllumination,
 * PrevFastIllumination, PrevHistoryLength) the next frame's temporal pass needs from the two neighbours. Inputs (Illumination and
 * the current G-buffer) must be valid on the band +- vpt_band_input_halo(params) rows. VPT_ERR_ARG when a band is smaller than
 * the rows a neighbour must supply, or when the camera pair can reproject further than the 32-row history guard. */
int vpt_denoise_band(vpt_ctx *ctx, const VptDenoisingParams *params, const VptCamera *camera, const VptCamera *prevCamera,
                     int frameNum, int iterationIndex, int rowBegin, int rowEnd);
/* Rows [rowBegin,rowEnd) of rank `rank` of `nranks` for an image of `height` rows (host only). */
void vpt_band_rows(int height, int nranks, int rank, int *rowBegin, int *rowEnd);
/* Rows either side of its band on which a rank's inputs must be valid for vpt_denoise_band with these settings (host only). */
int vpt_band_input_halo(const VptDenoisingParams *params);
/* The bands of IlluminationOutput collected on rank `root` (one grouped exchange): the frame a caller reads back. */
int vpt_comm_gather_output(vpt_ctx *ctx, int root);

/* ---- host-side helpers mirroring the reference's host code (no device work) */
/* Camera::init / Camera::update (renderer/shaders/Camera.h:29-99). */
void vpt_camera_init(VptCamera *cam, int width, int height);
void vpt_camera_update(VptCamera *cam);
/* mainOffline.cpp:227-247: position, direction (normalised like SceneConfigParser does) and horizontal fov in degrees. */
void vpt_camera_from_scene(VptCamera *cam, int width, int height, const float *position, const float *direction, float fovDegrees);
/* initVoxelsMultiChunk's CPU noise maps (VoxelSceneGen.cu:358-376): PerlinNoiseGenerator(4 octaves, seed), chunks x 32 x 32. */
void vpt_perlin_noise_chunks(int chunksX, int chunksY, int chunksZ, unsigned seed, float *out);
/* AliasTable::update, CPU build path (renderer/shaders/AliasTable.cu:66-153). */
void vpt_build_alias_table(const float *weights, unsigned n, VptAliasBin *bins);
/* GlobalSettings::LoadFromYAML (renderer/core/GlobalSettings.cpp:69-138), "denoising" section -> params. Returns
 * VPT_ERR_IO when the file cannot be opened (values stay default), like the reference returns false. */
int vpt_load_denoising_settings(const char *yamlPath, VptDenoisingParams *params);
void vpt_default_denoising_params(VptDenoisingParams *params);
/* SceneConfigParser::LoadFromFile (renderer/core/SceneConfig.cpp:6-114): camera position/direction(normalised)/up/fov
 * and chunk_config. out9 = pos[3], dir[3], up[3]. */
int vpt_load_scene_config(const char *yamlPath, float *out9, float *fov, unsigned *chunks3);

/* ---- output stage (SURVEY 8f "next" row 2): the deterministic part of PostProcessor::run.
 * ToneMappingParams (renderer/core/GlobalSettings.h:145-172); curve: 0 Narkowicz ACES, 1 Uncharted 2, 2 Reinhard. */
typedef struct VptToneMappingParams
{
    float manualExposure; /* 10 (C++ default); the shipped yaml sets 0.8 */
    int32_t curve;
    float highlightDesaturation;
    float whitePoint;
    float contrast;
    float saturation;
    float lift;
    float gain;
} VptToneMappingParams;
void vpt_default_tonemapping_params(VptToneMappingParams *params);
/* GlobalSettings::LoadFromYAML, "postprocess" section (GlobalSettings.cpp:275-296) and "sky" section. */
int vpt_load_tonemapping_settings(const char *yamlPath, VptToneMappingParams *params);
int vpt_load_sky_settings(const char *yamlPath, VptSkyParams *params);
/* FilmicToneMapping (renderer/postprocessing/FilmicToneMapping.h:58-117) with the MANUAL exposure (auto-exposure, bloom,
 * lens flare and vignette are wall-clock / history dependent and stay out of scope) on IlluminationOutput, then the PNG
 * conversion of OfflineBackend::writeFrameBufferToPNG (renderer/core/OfflineBackend.cpp:191-221: y flip, clamp, *255).
 * rgb8: W*H*3 bytes, top row first (may be NULL); rgbaLDR: W*H float4 sRGB-encoded, bottom row first (may be NULL). */
int vpt_tonemap(vpt_ctx *ctx, const VptToneMappingParams *params, uint8_t *rgb8, float *rgbaLDR);

/* ---- world chunk files (SURVEY 8f "next" row 4, I/O part): the reference's content-addressed raw chunk format
 * (renderer/core/WorldSceneManager.cpp:240-308: 32768 bytes per chunk in GetLinearId order, file name = FNV-1a-64 of the
 * bytes as 16 hex digits + ".bin") and the scene file that lists them (SceneConfigParser::SaveToFile / LoadFromFile,
 * renderer/core/SceneConfig.cpp:95-148: sections camera, chunk_config, chunks "index: hash"). Host only. */
void vpt_chunk_hash(const uint8_t *chunk32768, char *hex17);
/* WorldSceneManager::SaveScene (:310-359): writes every chunk to chunkDir/<hash>.bin and the scene file. cam9 = position,
 * direction, up. Returns VPT_ERR_IO if anything could not be written. */
int vpt_save_world(const char *sceneYamlPath, const char *chunkDir, int chunksX, int chunksY, int chunksZ, const uint8_t *ids,
                   const float *cam9, float fov);
/* WorldSceneManager::LoadScene (:361-458), chunk part: ids holds the runtime grid (chunksX*Y*Z*32768 bytes) and is overwritten
 * chunk by chunk for every "index: hash" record whose file exists and has the right size; bad records are skipped like the
 * reference does. loaded/failed (may be NULL) receive the record counts. Returns VPT_ERR_IO when the scene file cannot be opened. */
int vpt_load_world(const char *sceneYamlPath, const char *chunkDir, int chunksX, int chunksY, int chunksZ, uint8_t *ids, int *loaded, int *failed);

/* Test hook (no device work): the launch-invariant division the kernels use for index decoding. */
void vpt_debug_fastdiv(uint32_t n, uint32_t d, uint32_t *q, uint32_t *r);

#ifdef __cplusplus
}
#endif
#endif /* VPT_H */
