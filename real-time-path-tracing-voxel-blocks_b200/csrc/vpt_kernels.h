// Host-visible argument blocks and launchers of the sm_100a kernels (internal to libvpt.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vpt.h"
#include "vpt_fastdiv.h"
#include "vpt_lights.h"

namespace vpt {

// Device-side voxel grid (repacked):
//   idsChunk  chunk-major bytes exactly as the reference's VoxelChunk::data (the API layout)
//   idsLinear x + W*(z + D*y): 1 byte per voxel, fetched once per hit by the shading stage
//   occ       traversal mask, 1 bit per voxel of the PADDED volume (W+2) x (H+2) x (D+2) whose one-voxel shell is
//             all ones: a ray that leaves the grid "hits" the shell, so the DDA step loop carries no bounds test.
//             Rows are padded to Wp = roundup32(W+2) bits; voxel (x,y,z) is bit linP = ((y+1)*Dp + (z+1))*Wp + (x+1).
//             A second copy — the UPWARD mask — follows it: identical below y = upH (highest solid voxel + 1) and solid
//             from there up; rays with dir.y > 0 walk it and retire as soon as they rise above everything solid
//             (same results, fewer steps). Four all-zero spare words follow: lanes without a ray are parked on bit
//             parkLin there. 128x32x128 (16 chunks): 2 x 160 x 34 x 130 bits = 172.7 KiB -> staged whole in shared
//             memory by the DDA kernel (one 1024-thread CTA per SM).
struct GridView
{
    int W, H, D;        // voxels
    int Wp, Hp, Dp;     // padded mask dimensions
    int maskWords;      // Wp/32 * Hp * Dp: one padded mask
    int occWords;       // 2 * maskWords + 4: full mask, "upward" mask, spare words
    int parkLin;        // 2 * maskWords * 32
    int upH;            // the upward mask is solid from y = upH (highest solid voxel + 1): nothing to hit above it
    const uint32_t *occ;
    const uint8_t *idsLinear;
    FastDiv divW, divD, divWp, divDp;
};
inline int paddedW(int W) { return ((W + 2) + 31) & ~31; }
inline size_t paddedMaskWords(int W, int H, int D) { return (size_t)(paddedW(W) / 32) * (H + 2) * (D + 2); }
inline size_t paddedOccWords(int W, int H, int D) { return 2 * paddedMaskWords(W, H, D) + 4; }

struct GBufferPtrs
{
    float *depth, *material;
    float4 *normalRoughness, *geoNormalThinfilm, *materialParameter, *albedo;
};

// ---- wavefront path-tracing state (csrc/vpt_wave.cu). One "wave" = nSlots pixel slots x samplesInWave samples;
// path p = sampleInWave * nSlots + slot, slot = tile * 32 + lane of an 8x4 pixel tile.
struct WaveBuffers
{
    // per path
    float4 *dirT;       // ray direction of the current segment (xyz)
    float4 *org;        // ray origin of the current segment (depth > 0; depth 0 starts at the camera)
    float *hitT;        // closest-hit distance
    uint32_t *hitPacked; // (linear voxel index << 3) | face, 0xFFFFFFFF = miss
    float4 *surfA;      // spawn position xyz, hit distance
    float4 *surfB;      // wo xyz, bits: face | material index << 3
    float4 *surfC;      // textured scenes only: shading normal xyz, roughness (regularised)
    float4 *surfD;      // textured scenes only: albedo rgb, metallic
    uint32_t *pflag;    // path flags (vpt_wave.cu: F_*)
    float4 *candA;      // sun idx, sun weightSum, sun targetPdf, sky idx
    float4 *candB;      // sky weightSum, sky targetPdf, local-light uv (unquantised)
    uint4 *candC;       // local lights: selected light index (or 0xFFFFFFFF), weightSum bits, targetPdf bits
    float4 *dir1;       // direction of the BSDF-candidate ray
    uint4 *ris;         // lightData, uvData, weightSum, targetPdf (M == 1)
    float4 *lightA;     // selected light sample: direction xyz, solidAnglePdf
    float4 *lightB;     // radiance rgb, light type
    uint8_t *vis1, *vis2, *vis4; // occlusion results of ray #1 (BSDF candidate), #2 (RIS visibility), #5 (final visibility)
    float4 *rad;        // accumulated radiance rgb, primary hit distance
    float4 *thr;        // throughput rgb
    float4 *nextD;      // continuation direction
    float4 *bop;        // continuation bsdfOverPdf
    // per pixel slot (temporal ReSTIR, sample 0)
    uint4 *rstA;        // lightData, uvData, weightSum, targetPdf
    uint4 *rstB;        // M, meta (cached mask | selected idx | ray mask | valid), -, -
    float4 *light2A, *light2B;
    float4 *psA;        // ps[0..2], -
    float4 *psB;        // M[0..2], -
    uint8_t *vis3;      // 3 per slot: bias-correction rays
    // queues
    uint4 *queue;       // prepared rays, 48 bytes each
    int *listA, *listB; // active path lists for depth > 0 (ping-pong)
    unsigned *cnt;      // device counters: see vpt_wave.cu (CNT_*)
};

struct TraceArgs
{
    VptCamera cam, prevCam;
    int width, height;
    int iterationIndex, spp, totalBounceLimit, diffuseBounceLimit, enableRestir;
    int sampleBegin, sampleStep;
    int sampleLimit;             // > 0: at most this many samples of the shard (vpt_render_range)
    const uint32_t *occPrev; int upHPrev; // masks of the world as the previous render saw it (nullptr: unchanged); bias rays only
    int ownerSample; // the sample that owns the G-buffer, the reservoir and the temporal ReSTIR pass: 0, or sampleBegin with a rank-local owner
    GridView grid;
    int occInSmem;
    const uint8_t *sobol, *scrambling, *ranking;
    const VptMaterial *materials;
    const uint16_t *blockToMaterial;
    const float4 *sky, *sun;
    const VptAliasBin *skyAlias, *sunAlias;
    int skyW, skyH, sunW, sunH;
    float sunDir[3];
    float sunCosThetaMax;
    // textured materials (optional): RGBA8 mip chains, per-texture (first texel, width, levels), per-material slots + |texSize|
    const uint32_t *texels;
    const int4 *texDescs;
    const int4 *matTexSlots;
    const float *matTexMip0Size;
    int nTextures;
    LightView lv;                // local emissive lights (numLights == 0: none, and no stage does anything for them)
    GBufferPtrs cur, prev;
    float4 *illumination;
    VptReservoir *resCur;
    const VptReservoir *resPrev;
    int4 *primaryHits;
    unsigned long long *counters; // [0] rays, [1] steps
    // wavefront
    WaveBuffers wb;
    int tilesX, nSlots;          // 8x4 tiles per row, slots = tiles * 32
    FastDiv divSlots, divTilesX, divSkyW, divSunW;
    int samplesInWave, waveFirst; // this wave renders local samples waveFirst .. waveFirst+samplesInWave-1 of the shard
    int nPaths;                  // nSlots * samplesInWave
    // a wave is cut into PARTS of whole tiles, each with its own queue / counters, run on its own stream so the
    // issue-bound DDA kernel of one part overlaps the latency-bound shading stage of the other
    int slotBase, partSlots, partPaths; // this launch covers slots [slotBase, slotBase + partSlots) of every sample in the wave
    FastDiv divPartSlots;
    int depthRounds;             // 1 when no path can continue past its first hit (all-diffuse materials, diffuse limit 1)
    int countSteps;              // DDA step statistics on/off
    float resolveSpp;            // > 0: the last wave's accumulate also divides by spp (vpt_render); 0: leave the sum (vpt_render_shard)
};

struct WaveWorkspace
{
    void *arena = nullptr;
    size_t arenaBytes = 0;
    int nSlots = 0, maxSamplesInWave = 0;
    WaveBuffers wb = {};
};
// bytes of workspace for nSlots pixel slots and samplesInWave samples per wave; carve() lays the arena out
size_t waveWorkspaceBytes(int nSlots, int samplesInWave);
void waveCarve(WaveWorkspace &ws, int nSlots, int samplesInWave);

struct DenoiseBuffers
{
    GBufferPtrs cur, prev;
    float4 *illumination, *illumOutput, *ping, *pong, *prevIllum, *prevFastIllum;
    float *historyLength, *prevHistoryLength;
    VptReservoir *reservoirs; // plane of the parity being filtered
};

struct FireflyPatch
{
    int pixel;
    float4 color;
    VptReservoir reservoir;
};

// Optional per-launch timing: an event is recorded after every kernel of launchTrace (kind 0 = DDA, 1 = shading stage).
struct TraceProfile
{
    static constexpr int kMax = 96;
    cudaEvent_t ev[kMax + 1] = {}; // ev[0] = start, ev[i+1] = after launch i
    int kind[kMax];
    int n = 0;
    bool enabled = false;
};
// Renders every sample of the shard (a.sampleBegin/a.sampleStep) wave by wave; returns the number of kernels launched.
struct TraceStreams
{
    cudaStream_t part[2] = {nullptr, nullptr}; // side streams, one per part
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
};
cudaError_t launchTrace(TraceArgs &a, int maxSamplesInWave, cudaStream_t s, const TraceStreams *ts, int smCount, size_t smemOptIn, int *launches,
                        TraceProfile *prof);
cudaError_t launchResolve(float4 *illum, int npix, float spp, cudaStream_t s);

// DDA over a queue of prepared rays (csrc/vpt_dda.cu)
struct DdaArgs
{
    const uint4 *queue;
    const unsigned *count; // entries in the queue (device)
    unsigned *cursor;      // chunk cursor (device, zero at launch)
    float *hitT;           // closest mode outputs, indexed by the ray's result slot
    uint32_t *hitPacked;
    uint8_t *vis;          // visibility mode output
    GridView grid;
    unsigned long long *counters;
};
cudaError_t launchDda(const DdaArgs &a, bool closest, bool occInSmem, bool countSteps, bool farEnd, bool nearEnd, cudaStream_t s, int smCount); // farEnd: rays carry a finite tmax; nearEnd: a tmin > 0

// upH (device int): highest solid y + 1, maintained by the repack / set-voxel kernels; the upward mask is rebuilt from it
// sky generator (vpt_sky.cu)
cudaError_t launchSkyUpper(const float *configs, const float *radiances, const float *sunDir, float brightness, float4 *sky, float *pdf, int W, int H,
                           cudaStream_t s);
cudaError_t launchSkyLower(float4 *sky, float *pdf, int W, int H, float sumSkyPdf, cudaStream_t s);
cudaError_t launchTonemap(const float4 *hdr, int W, int H, const VptToneMappingParams &p, uint8_t *rgb8, float4 *ldr, cudaStream_t s);
cudaError_t launchSkySun(const float *sunDir, float brightness, const float *solar, const float *limb, float4 *sun, float *pdf, int W, int H, cudaStream_t s);

cudaError_t launchRepackGrid(const uint8_t *idsChunk, uint8_t *idsLinear, uint32_t *occ, int *upHDev, int *upHHost, int cx, int cy, int cz, cudaStream_t s);
cudaError_t launchSetVoxel(uint8_t *idsChunk, uint8_t *idsLinear, uint32_t *occ, int *upHHost, int cx, int cy, int cz, int x, int y, int z, int id, cudaStream_t s);
cudaError_t launchPick(const uint8_t *idsLinear, int cx, int cy, int cz, const float *origin, const float *dir, VptPickResult *outDev, cudaStream_t s);
cudaError_t launchGenerateTerrain(const float *noise, uint8_t *idsChunk, int cx, int cy, int cz, cudaStream_t s);

// Pixel-space view-vector basis of the current camera: M*(u,v,1) = M0 + x*Mx + y*My (vpt_denoise.cu)
struct DnView { float pos[3], M0[3], Mx[3], My[3]; };
DnView makeDnView(const VptCamera &c);

// Denoiser passes; rows [rowBegin,rowEnd) are processed (whole image: 0,H).
struct DenoiseLaunch
{
    int width, height, rowBegin, rowEnd;
    VptCamera cam, prevCam;
    VptDenoisingParams p;
    DenoiseBuffers b;
    DnView view;
    float4 *G;        // packed denoiser G-buffer: normal.xyz, view-scaled depth
    uint32_t *MQ;     // material id << 16 | Load2DUshort1 value
    unsigned *counters; // [0] firefly candidates, [1] HistoryFix list length
    int4 *fireflyList;  // firefly candidates found by the prep pass
    int *fixList;       // pixels with historyLength <= 4 (HistoryFix work list, filled by the temporal pass)
    cudaStream_t stream;
};
// packed G-buffer over rows [prepRow0,prepRow1) + sky copy; firefly detection/apply over the launch rows
cudaError_t launchPrep(const DenoiseLaunch &d, int prepRow0, int prepRow1, bool firefly, FireflyPatch *patches, int maxPatches);
cudaError_t launchTemporal(const DenoiseLaunch &d);
cudaError_t launchHistoryFix(const DenoiseLaunch &d, int smCount);
cudaError_t launchHistoryClamping(const DenoiseLaunch &d);
cudaError_t launchAtrousSmem(const DenoiseLaunch &d, const float4 *in, float4 *out);
cudaError_t launchAtrous(const DenoiseLaunch &d, const float4 *in, float4 *out, unsigned frameIndex, unsigned step, bool composite);
cudaError_t launchCompositeNonSky(const DenoiseLaunch &d, const float4 *finalBuf);
// Shared-memory tile versions (vpt_dn_tiles.cu: TMA + mbarrier halo tiles; column-walking history clamp). *handled = false when the
// shape is outside what the tiles cover (a-trous steps other than 2/4/8, widths that are not a multiple of 4): the caller then
// uses the gather kernel above.
cudaError_t launchAtrousTiled(const DenoiseLaunch &d, const float4 *in, float4 *out, unsigned frameIndex, unsigned step, bool composite, bool *handled);
cudaError_t launchAtrousSmemTiled(const DenoiseLaunch &d, const float4 *in, float4 *out, bool *handled);
cudaError_t launchHistoryClampingCols(const DenoiseLaunch &d);
bool tileClampEnabled();
unsigned debugTmaTimeouts(); // tile loads that timed out since the process started (0 unless a tensor map / byte count is wrong)
cudaError_t launchFrame0Init(const DenoiseLaunch &d);
cudaError_t launchHitDist(const DenoiseLaunch &d);                 // IlluminationBuffer.w -> IlluminationPing (default off)
cudaError_t launchPrePass(const DenoiseLaunch &d, int frameIndex); // IlluminationPing -> IlluminationBuffer (default off)

} // namespace vpt
