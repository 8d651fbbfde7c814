// Voxel grid producers and the device-side repack.
//
//  * generateTerrainKernel == GenerateVoxelChunk (/root/reference/voxelengine/VoxelSceneGen.cu:61-165):
//    height-field terrain from a per-chunk 32x32 noise map, written in the reference's chunk-major
//    GetLinearId order (voxelengine/VoxelMath.h:120-127). Compiled -fmad=false: the float height tests
//    must agree with the CPU restatement for every voxel (bit-exact ids).
//  * repackIdsKernel / maxSolidYKernel / repackMaskKernel build the traversal layouts from the chunk-major bytes: a linear id
//    volume x + W*(z + D*y) and the padded 1-bit occupancy mask with a solid one-voxel shell (GridView,
//    vpt_kernels.h; one warp ballot per 32 x-consecutive bits).
#include "vpt_kernels.h"

namespace vpt {

__device__ __forceinline__ float fminr_(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float fmaxr_(float a, float b) { return a > b ? a : b; }

// One thread per voxel, x fastest: a warp covers 32 x-consecutive voxels of one (y,z) row of one chunk.
__global__ void generateTerrainKernel(const float *__restrict__ noise, uint8_t *__restrict__ idsChunk, int cx, int cy, int cz)
{
    const size_t total = (size_t)cx * cy * cz * 32768;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int chunk = (int)(i >> 15);
    const int local = (int)(i & 32767);
    const int x = local & 31, z = (local >> 5) & 31, y = local >> 10;
    const int chunkX = chunk % cx, chunkZ = (chunk / cx) % cz, chunkY = chunk / (cx * cz);
    const float noiseVal = __ldg(noise + (size_t)chunk * 1024 + z * 32 + x);
    const unsigned width = (unsigned)(32 * cy); // height scale: 32 in the reference (chunksY == 1)
    const int gy = chunkY * 32 + y;
    const unsigned gx = chunkX * 32 + x, gz = chunkZ * 32 + z;
    uint8_t id = 0;
    float terrainHeight = fmaxr_(0.1f, (noiseVal * 1.4f - 0.7f + 0.25f) * width);
    terrainHeight = fminr_(terrainHeight, width * 0.9f);
    if (gy < terrainHeight)
    {
        float verticalDepth = terrainHeight - gy;
        if (terrainHeight < width * (0.25f + 0.05f))
            id = (verticalDepth < 3.5f) ? 1 : 7; // Sand / Rocks
        else if (terrainHeight < width * (0.25f + 0.6f) && terrainHeight > width * (0.25f + 0.3f))
            id = (verticalDepth < 5.5f) ? 3 : 7; // Cliff / Rocks
        else
            id = (verticalDepth < 1.5f) ? 2 : (verticalDepth < 5.5f ? 3 : 7); // Soil / Cliff / Rocks
    }
    // the ten shader-ball mesh cells (VoxelSceneGen.cu:126-161) hold no cube: empty in the DDA world
    if (gy == 7 && gz == 43 && gx >= 30 && gx <= 39) id = 0;
    idsChunk[i] = id;
}

// idsLinear: one thread per voxel in LINEAR order (x fastest over the whole world width).
__global__ void repackIdsKernel(const uint8_t *__restrict__ idsChunk, uint8_t *__restrict__ idsLinear, int cx, int cy, int cz)
{
    const int W = cx * 32, D = cz * 32;
    const size_t total = (size_t)cx * cy * cz * 32768;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W);
    const size_t r = i / W;
    const int z = (int)(r % D), y = (int)(r / D);
    const int chunk = (x >> 5) + cx * ((z >> 5) + cz * (y >> 5));
    idsLinear[i] = __ldg(idsChunk + (size_t)chunk * 32768 + (x & 31) + 32 * ((z & 31) + 32 * (y & 31)));
}

// Highest solid voxel: upH = max(y + 1) over solid voxels (0 for an empty world).
__global__ void maxSolidYKernel(const uint8_t *__restrict__ idsLinear, int *upH, int W, int D, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int y1 = 0;
    if (i < total && __ldg(idsLinear + i) != 0) y1 = (int)(i / ((size_t)W * D)) + 1;
    y1 = __reduce_max_sync(0xffffffffu, y1);
    if ((threadIdx.x & 31) == 0 && y1 > 0) atomicMax(upH, y1);
}

// Padded traversal masks (GridView): one thread per bit of the padded volume, x fastest, so the 32 lanes of a warp
// are the 32 bits of one word -> __ballot_sync builds it. The one-voxel shell (and the row padding beyond it) is solid.
// Mask 0 is the grid; mask 1 (the upward mask, for rays with dir.y > 0) is additionally solid from y = upH up.
// Words are stored BIT-REVERSED (voxel k of a word at bit 31-k): the DDA tests `(int)(word << (lin & 31)) < 0`, a shift
// whose count is the low five bits of lin as they are, and a sign test (2 instructions instead of shift + and + compare).
__global__ void repackMaskKernel(const uint8_t *__restrict__ idsLinear, uint32_t *__restrict__ occ, int W, int H, int D, int Wp, int upH)
{
    const int Dp = D + 2, Hp = H + 2;
    const size_t total = (size_t)Wp * Hp * Dp; // a multiple of 32; the block size is a multiple of 32
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool solid = false, shell = false;
    int y = 0;
    if (i < total)
    {
        const int xp = (int)(i % Wp);
        const size_t r = i / Wp;
        const int zp = (int)(r % Dp), yp = (int)(r / Dp);
        const int x = xp - 1, z = zp - 1;
        y = yp - 1;
        if (x < 0 || y < 0 || z < 0 || x >= W || y >= H || z >= D) shell = true;
        else solid = __ldg(idsLinear + ((size_t)y * D + z) * W + x) != 0;
    }
    const unsigned word0 = __brev(__ballot_sync(0xffffffffu, solid || shell));
    const unsigned word1 = __brev(__ballot_sync(0xffffffffu, solid || shell || y >= upH));
    if ((threadIdx.x & 31) == 0)
    {
        if (i < total) { occ[i >> 5] = word0; occ[(total >> 5) + (i >> 5)] = word1; }
        else if (i < total + 128) occ[2 * (total >> 5) + ((i - total) >> 5)] = 0u; // the four spare (parking) words
    }
}

__global__ void setVoxelKernel(uint8_t *idsChunk, uint8_t *idsLinear, uint32_t *occ, int cx, int cy, int cz, int Wp, int upH, int x, int y, int z, int id)
{
    const int W = cx * 32, H = cy * 32, D = cz * 32, Dp = D + 2;
    const int chunk = (x >> 5) + cx * ((z >> 5) + cz * (y >> 5));
    idsChunk[(size_t)chunk * 32768 + (x & 31) + 32 * ((z & 31) + 32 * (y & 31))] = (uint8_t)id;
    idsLinear[((size_t)y * D + z) * W + x] = (uint8_t)id;
    const size_t linP = ((size_t)(y + 1) * Dp + (z + 1)) * Wp + (x + 1);
    const size_t maskWords = (size_t)(Wp / 32) * (H + 2) * Dp;
    const uint32_t bit = 0x80000000u >> (linP & 31); // bit-reversed words
    uint32_t w = occ[linP >> 5];
    occ[linP >> 5] = id ? (w | bit) : (w & ~bit);
    if (y < upH) // above upH the upward mask stays solid
    {
        w = occ[maskWords + (linP >> 5)];
        occ[maskWords + (linP >> 5)] = id ? (w | bit) : (w & ~bit);
    }
}

cudaError_t launchGenerateTerrain(const float *noise, uint8_t *idsChunk, int cx, int cy, int cz, cudaStream_t s)
{
    const size_t total = (size_t)cx * cy * cz * 32768;
    generateTerrainKernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(noise, idsChunk, cx, cy, cz);
    return cudaGetLastError();
}
static cudaError_t launchMask(const uint8_t *idsLinear, uint32_t *occ, int cx, int cy, int cz, int upH, cudaStream_t s)
{
    const int Wp = paddedW(cx * 32);
    const size_t bits = (size_t)Wp * (cy * 32 + 2) * (cz * 32 + 2) + 128;
    repackMaskKernel<<<(unsigned)((bits + 255) / 256), 256, 0, s>>>(idsLinear, occ, cx * 32, cy * 32, cz * 32, Wp, upH);
    return cudaGetLastError();
}
// Builds idsLinear, finds upH (synchronises the stream to read it back: grid uploads are not on the frame path) and
// builds both masks.
cudaError_t launchRepackGrid(const uint8_t *idsChunk, uint8_t *idsLinear, uint32_t *occ, int *upHDev, int *upHHost, int cx, int cy, int cz, cudaStream_t s)
{
    const size_t total = (size_t)cx * cy * cz * 32768;
    repackIdsKernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(idsChunk, idsLinear, cx, cy, cz);
    cudaError_t e = cudaMemsetAsync(upHDev, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    maxSolidYKernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(idsLinear, upHDev, cx * 32, cz * 32, total);
    e = cudaMemcpyAsync(upHHost, upHDev, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    return launchMask(idsLinear, occ, cx, cy, cz, *upHHost, s);
}
// upH is the host's copy; a solid block above it raises it and rebuilds the masks (rare: an edit above the skyline).
cudaError_t launchSetVoxel(uint8_t *idsChunk, uint8_t *idsLinear, uint32_t *occ, int *upHHost, int cx, int cy, int cz, int x, int y, int z, int id, cudaStream_t s)
{
    setVoxelKernel<<<1, 1, 0, s>>>(idsChunk, idsLinear, occ, cx, cy, cz, paddedW(cx * 32), *upHHost, x, y, z, id);
    if (id != 0 && y + 1 > *upHHost)
    {
        *upHHost = y + 1;
        return launchMask(idsLinear, occ, cx, cy, cz, *upHHost, s);
    }
    return cudaGetLastError();
}

// The block picker (VoxelEngine::performRayTraversal, voxelengine/VoxelEngine.cu:1040-1166): one Amanatides-Woo walk from the
// camera along camera.dir over the resident ids, IEEE arithmetic throughout (this TU is built -fmad=false; divisions and the
// square root are the round-to-nearest intrinsics), at most 1000 voxels, stops when it leaves the grid. One thread.
__global__ void pickKernel(const uint8_t *__restrict__ idsLinear, int W, int H, int D, float ox, float oy, float oz, float dx, float dy, float dz,
                           VptPickResult *out)
{
    VptPickResult r;
    r.hasSpaceToCreate = 0; r.hitSurface = 0; r.deleteBlockId = -1;
    for (int k = 0; k < 3; ++k) { r.createPos[k] = -1; r.deletePos[k] = -1; }
    const float len = __fsqrt_rn(dx * dx + dy * dy + dz * dz);
    if (!(len <= 1e-8f))
    {
        const float o[3] = {ox, oy, oz}, d[3] = {__fdiv_rn(dx, len), __fdiv_rn(dy, len), __fdiv_rn(dz, len)};
        const int dim[3] = {W, H, D};
        int v[3], step[3];
        float tDelta[3], tMax[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            v[k] = (int)floorf(o[k]);
            step[k] = d[k] > 0.0f ? 1 : -1;
            const bool flat = fabsf(d[k]) < 1e-8f;
            tDelta[k] = flat ? 3.402823466e+38f : __fdiv_rn(1.0f, fabsf(d[k]));
            const float boundary = step[k] > 0 ? (float)(v[k] + 1) : (float)v[k];
            tMax[k] = flat ? 3.402823466e+38f : __fdiv_rn(boundary - o[k], d[k]);
        }
        for (int it = 0; it < 1000; ++it)
        {
            if (v[0] < 0 || v[0] >= W || v[1] < 0 || v[1] >= H || v[2] < 0 || v[2] >= D) break;
            const int id = idsLinear[(size_t)v[0] + (size_t)W * ((size_t)v[2] + (size_t)D * (size_t)v[1])];
            if (id == 0) { r.hasSpaceToCreate = 1; r.createPos[0] = v[0]; r.createPos[1] = v[1]; r.createPos[2] = v[2]; }
            else
            {
                r.hitSurface = 1; r.deleteBlockId = id;
                r.deletePos[0] = v[0]; r.deletePos[1] = v[1]; r.deletePos[2] = v[2];
                break;
            }
            // x only when strictly smallest, y over z only when strictly smaller, z otherwise (the reference's comparison order)
            const int axis = tMax[0] < tMax[1] ? (tMax[0] < tMax[2] ? 0 : 2) : (tMax[1] < tMax[2] ? 1 : 2);
            v[axis] += step[axis];
            tMax[axis] = tMax[axis] + tDelta[axis];
        }
        (void)dim;
    }
    *out = r;
}
cudaError_t launchPick(const uint8_t *idsLinear, int cx, int cy, int cz, const float *origin, const float *dir, VptPickResult *outDev, cudaStream_t s)
{
    pickKernel<<<1, 1, 0, s>>>(idsLinear, cx * 32, cy * 32, cz * 32, origin[0], origin[1], origin[2], dir[0], dir[1], dir[2], outDev);
    return cudaGetLastError();
}

} // namespace vpt
