// Host-visible argument blocks and launchers of the sm_100a kernels (internal to libvpt.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vpt.h"

namespace vpt {

// Device-side voxel grid (repacked):
//   idsChunk  chunk-major bytes exactly as the reference's VoxelChunk::data (the API layout)
//   idsLinear x + W*(z + D*y): the traversal layout (1 byte, fetched once per hit)
//   occ       1 bit per voxel, 32 x-consecutive voxels per word, word = (y*D + z)*(W/32) + x/32.
//             128x32x128 (16 chunks) = 64 KiB -> staged whole in shared memory by the trace kernel.
struct GridView
{
    int W, H, D;        // voxels
    int cx, cy, cz;     // chunks
    int wordsX;         // W/32
    int occWords;       // wordsX*H*D
    const uint32_t *occ;
    const uint8_t *idsLinear;
};

struct GBufferPtrs
{
    float *depth, *material;
    float4 *normalRoughness, *geoNormalThinfilm, *materialParameter, *albedo;
};

struct TraceArgs
{
    VptCamera cam, prevCam;
    int width, height;
    int iterationIndex, spp, totalBounceLimit, diffuseBounceLimit, enableRestir;
    int sampleBegin, sampleStep;
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
    GBufferPtrs cur, prev;
    float4 *illumination;
    VptReservoir *resCur;
    const VptReservoir *resPrev;
    int4 *primaryHits;
    unsigned long long *counters; // [0] rays, [1] steps, [2] tile scheduler
};

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

cudaError_t launchTrace(const TraceArgs &a, cudaStream_t s, int smCount, size_t smemOptIn);
cudaError_t launchResolve(float4 *illum, int npix, float spp, cudaStream_t s);

cudaError_t launchRepackGrid(const uint8_t *idsChunk, uint8_t *idsLinear, uint32_t *occ, int cx, int cy, int cz, cudaStream_t s);
cudaError_t launchSetVoxel(uint8_t *idsChunk, uint8_t *idsLinear, uint32_t *occ, int cx, int cy, int cz, int x, int y, int z, int id, cudaStream_t s);
cudaError_t launchGenerateTerrain(const float *noise, uint8_t *idsChunk, int cx, int cy, int cz, cudaStream_t s);

// Denoiser passes; rows [rowBegin,rowEnd) are processed (whole image: 0,H).
struct DenoiseLaunch
{
    int width, height, rowBegin, rowEnd;
    VptCamera cam, prevCam;
    VptDenoisingParams p;
    DenoiseBuffers b;
    cudaStream_t stream;
};
cudaError_t launchFirefly(const DenoiseLaunch &d, FireflyPatch *patches, int *patchCount, int maxPatches);
cudaError_t launchCopySky(const DenoiseLaunch &d);
cudaError_t launchTemporal(const DenoiseLaunch &d);
cudaError_t launchHistoryFix(const DenoiseLaunch &d);
cudaError_t launchHistoryClamping(const DenoiseLaunch &d);
cudaError_t launchAtrousSmem(const DenoiseLaunch &d, const float4 *in, float4 *out);
cudaError_t launchAtrous(const DenoiseLaunch &d, const float4 *in, float4 *out, unsigned frameIndex, unsigned step);
cudaError_t launchCompositeNonSky(const DenoiseLaunch &d, const float4 *finalBuf);
cudaError_t launchFrame0Init(const DenoiseLaunch &d);

} // namespace vpt
