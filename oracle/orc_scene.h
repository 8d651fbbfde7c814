// ORACLE — test infrastructure only (see orc_math.h header). CPU restatement of the
// reference's camera, RNG, voxel grid, terrain producer and Perlin noise.
//
//   Camera            /root/reference/renderer/shaders/Camera.h:6-149, setup mainOffline.cpp:227-247
//   RNG               /root/reference/renderer/shaders/RandGen.h:21-45, SystemParameter.h:143-174
//   Voxel grid        /root/reference/voxelengine/VoxelMath.h:120-127 (GetLinearId),
//                     VoxelEngine.cu:221-263 (chunk index, getVoxelAtGlobal)
//   Terrain producer  /root/reference/voxelengine/VoxelSceneGen.cu:61-165 (GenerateVoxelChunk), :341-388
//   Perlin noise      /root/reference/voxelengine/ext/PerlinNoise.hpp:228-243 (Shuffle), :449-497 (noise3D),
//                     :315-330 (Octave2D), :281-292 (RemapClamp_01); Noise.cpp:4-18 (4 octaves, seed 124)
//                     -- third-party siv::PerlinNoise v3, vendored in the reference tree; pinned bit-exactly
//                     against oracle/_ref/libref_noise.so (compiled from the reference's own Noise.cpp).
#pragma once
#include "orc_math.h"
#include <vector>
#include <random>
#include <array>

namespace orc {

// ---------------------------------------------------------------- Camera (212-byte POD, Camera.h:6-28)
struct Camera
{
    f2 resolution, inversedResolution, tanHalfFov;
    f3 pos, dir, posDelta;
    float yaw, pitch;
    mat3 uvToWorld, worldToUv, uvToView, viewToUv;
};
static_assert(sizeof(Camera) == 212, "Camera POD must be 212 bytes");

inline void cameraInit(Camera &c, int width, int height)
{
    std::memset(&c, 0, sizeof c);
    c.pos = {16.0f, 25.0f, 16.0f};
    c.dir = normalize(F3(1.0f, -1.0f, 1.0f));
    c.resolution = {(float)width, (float)height};
    c.inversedResolution = {1.0f / c.resolution.x, 1.0f / c.resolution.y};
    float fovX = 90.0f * kPiOver180;
    float fovY = fovX * (c.resolution.y / c.resolution.x);
    c.tanHalfFov = {tanf(fovX * 0.5f), tanf(fovY * 0.5f)};
}
inline void cameraUpdateMatrices(Camera &c)
{
    c.dir = yawPitchToDir(c.yaw, c.pitch);
    f3 worldUp = {0.0f, 1.0f, 0.0f};
    f3 left = normalize(cross(worldUp, c.dir));
    f3 up = normalize(cross(c.dir, left));
    mat3 uvToNdc = mat3Cols({2, 0, 0}, {0, 2, 0}, {-1, -1, 1});
    mat3 ndcToView = mat3Zero();
    ndcToView.m00 = c.tanHalfFov.x; ndcToView.m11 = c.tanHalfFov.y; ndcToView.m22 = 1.0f;
    mat3 viewToWorld = mat3Cols(-left, up, c.dir);
    c.uvToWorld = mul(mul(viewToWorld, ndcToView), uvToNdc);
    c.uvToView = mul(viewToWorld, ndcToView);
    mat3 ndcToUv = mat3Cols({0.5f, 0, 0}, {0, 0.5f, 0}, {0.5f, 0.5f, 1.0f});
    mat3 worldToView = transpose(viewToWorld);
    mat3 viewToNdc = mat3Zero();
    viewToNdc.m00 = 1.0f / c.tanHalfFov.x; viewToNdc.m11 = 1.0f / c.tanHalfFov.y; viewToNdc.m22 = 1.0f;
    c.worldToUv = mul(mul(ndcToUv, viewToNdc), worldToView);
    c.viewToUv = mul(viewToNdc, worldToView);
}
inline void cameraUpdate(Camera &c)
{
    if (std::isnan(c.posDelta.x) || std::isnan(c.posDelta.y) || std::isnan(c.posDelta.z) ||
        fabsf(c.posDelta.x) > 1000.0f || fabsf(c.posDelta.y) > 1000.0f || fabsf(c.posDelta.z) > 1000.0f)
        c.posDelta = {0, 0, 0};
    c.pos = c.pos + c.posDelta;
    c.posDelta = {0, 0, 0};
    cameraUpdateMatrices(c);
}
// Camera::getRayConeWidth (Camera.h:133-149): angular width of the pixel
inline float getRayConeWidth(const Camera &c, int ix, int iy)
{
    const f2 pixelCenter = {((float)ix + 0.5f) - c.resolution.x / 2, ((float)iy + 0.5f) - c.resolution.y / 2};
    const f2 pixelOffset = {copysignf(0.5f, pixelCenter.x), copysignf(0.5f, pixelCenter.y)};
    const f2 uvNear = {(pixelCenter.x - pixelOffset.x) * c.inversedResolution.x * 2, (pixelCenter.y - pixelOffset.y) * c.inversedResolution.y * 2};
    const f2 uvFar = {(pixelCenter.x + pixelOffset.x) * c.inversedResolution.x * 2, (pixelCenter.y + pixelOffset.y) * c.inversedResolution.y * 2};
    const f2 pn = {uvNear.x * c.tanHalfFov.x, uvNear.y * c.tanHalfFov.y}, pf = {uvFar.x * c.tanHalfFov.x, uvFar.y * c.tanHalfFov.y};
    const float angleNear = atanf(sqrtf(pn.x * pn.x + pn.y * pn.y)), angleFar = atanf(sqrtf(pf.x * pf.x + pf.y * pf.y));
    return angleFar - angleNear;
}
inline f3 uvToWorldDirection(const Camera &c, f2 uv) { return normalize(mul(c.uvToWorld, F3(uv.x, uv.y, 1.0f))); }
inline f2 worldDirectionToUV(const Camera &c, f3 d)
{
    f3 h = mul(c.worldToUv, d);
    return {h.x / h.z, h.y / h.z};
}
inline float pixelWorldSizeScaleToDepth(const Camera &c) { return c.tanHalfFov.x / (c.resolution.x / 2); }

// ---------------------------------------------------------------- RNG (RandGen.h:21-45)
// rankingTile is read with an un-wrapped dimension (RandGen.h:30) so the last pixels of the tile
// index up to 247 bytes past its end; the oracle defines those bytes as 0 (padded table).
struct Tables
{
    std::vector<uint8_t> sobol;      // 65536
    std::vector<uint8_t> scrambling; // 131072
    std::vector<uint8_t> ranking;    // 131072 + 256 zero pad
};
inline float blueNoiseRand(const Tables &t, int px, int py, int sampleIndex, int dim)
{
    px &= 127; py &= 127; sampleIndex &= 255;
    int ranked = sampleIndex ^ t.ranking[dim + (px + py * 128) * 8];
    int value = t.sobol[dim + ranked * 256];
    value ^= t.scrambling[(dim % 8) + (px + py * 128) * 8];
    return value / 256.0f;
}

// ---------------------------------------------------------------- Voxel grid
// ids are chunk-major: chunk index cx + CX*(cz + CZ*cy) (VoxelEngine.cu:221-224), 32768 bytes per chunk
// in GetLinearId order x + 32*(z + 32*y) (VoxelMath.h:120-127).
struct Grid
{
    int cx = 0, cy = 0, cz = 0; // chunk counts
    std::vector<uint8_t> ids;
    int W() const { return cx * 32; }
    int H() const { return cy * 32; }
    int D() const { return cz * 32; }
    inline size_t index(int x, int y, int z) const
    {
        int chunk = (x >> 5) + cx * ((z >> 5) + cz * (y >> 5));
        return (size_t)chunk * 32768 + (x & 31) + 32 * ((z & 31) + 32 * (y & 31));
    }
    inline uint8_t at(int x, int y, int z) const
    {
        if (x < 0 || y < 0 || z < 0 || x >= W() || y >= H() || z >= D()) return 0;
        return ids[index(x, y, z)];
    }
};

// ---------------------------------------------------------------- Perlin noise (siv::BasicPerlinNoise<float>)
struct Perlin
{
    std::array<uint8_t, 256> perm;
    explicit Perlin(uint32_t seed)
    {
        for (int i = 0; i < 256; ++i) perm[i] = (uint8_t)i;
        std::mt19937 rng(seed);
        // perlin_detail::Shuffle (PerlinNoise.hpp:228-243): for it=1..255 swap(it, rng() % (it+1))
        for (int i = 1; i < 256; ++i)
        {
            uint64_t n = (uint64_t)i;
            uint64_t j = (uint64_t)rng() % (n + 1);
            std::swap(perm[i], perm[(size_t)j]);
        }
    }
    static float fade(float t) { return t * t * t * (t * (t * 6 - 15) + 10); }
    static float lerp(float a, float b, float t) { return a + (b - a) * t; }
    static float grad(uint8_t hash, float x, float y, float z)
    {
        const uint8_t h = hash & 15;
        const float u = h < 8 ? x : y;
        const float v = h < 4 ? y : (h == 12 || h == 14 ? x : z);
        return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
    }
    float noise3D(float x, float y, float z) const
    {
        const float _x = std::floor(x), _y = std::floor(y), _z = std::floor(z);
        const int ix = (int)_x & 255, iy = (int)_y & 255, iz = (int)_z & 255;
        const float fx = x - _x, fy = y - _y, fz = z - _z;
        const float u = fade(fx), v = fade(fy), w = fade(fz);
        const uint8_t A = (perm[ix & 255] + iy) & 255;
        const uint8_t B = (perm[(ix + 1) & 255] + iy) & 255;
        const uint8_t AA = (perm[A] + iz) & 255;
        const uint8_t AB = (perm[(A + 1) & 255] + iz) & 255;
        const uint8_t BA = (perm[B] + iz) & 255;
        const uint8_t BB = (perm[(B + 1) & 255] + iz) & 255;
        const float p0 = grad(perm[AA], fx, fy, fz);
        const float p1 = grad(perm[BA], fx - 1, fy, fz);
        const float p2 = grad(perm[AB], fx, fy - 1, fz);
        const float p3 = grad(perm[BB], fx - 1, fy - 1, fz);
        const float p4 = grad(perm[(AA + 1) & 255], fx, fy, fz - 1);
        const float p5 = grad(perm[(BA + 1) & 255], fx - 1, fy, fz - 1);
        const float p6 = grad(perm[(AB + 1) & 255], fx, fy - 1, fz - 1);
        const float p7 = grad(perm[(BB + 1) & 255], fx - 1, fy - 1, fz - 1);
        const float q0 = lerp(p0, p1, u), q1 = lerp(p2, p3, u), q2 = lerp(p4, p5, u), q3 = lerp(p6, p7, u);
        const float r0 = lerp(q0, q1, v), r1 = lerp(q2, q3, v);
        return lerp(r0, r1, w);
    }
    float octave2D_01(float x, float y, int octaves, float persistence = 0.5f) const
    {
        float result = 0, amplitude = 1;
        for (int i = 0; i < octaves; ++i)
        {
            result += noise3D(x, y, (float)0.34567) * amplitude; // SIVPERLIN_DEFAULT_Z (PerlinNoise.hpp:76)
            x *= 2; y *= 2; amplitude *= persistence;
        }
        if (result <= -1.0f) return 0.0f;
        if (1.0f <= result) return 1.0f;
        return result * 0.5f + 0.5f;
    }
};

// Noise map of one chunk, noise[z*32+x] (VoxelSceneGen.cu:361-376)
inline void chunkNoiseMap(const Perlin &p, int chunkX, int chunkZ, int globalWidth, float *noise)
{
    float freq = 1.0f / globalWidth;
    for (int x = 0; x < 32; ++x)
        for (int z = 0; z < 32; ++z)
        {
            float gx = (float)(chunkX * 32 + x);
            float gz = (float)(chunkZ * 32 + z);
            noise[z * 32 + x] = p.octave2D_01(gx * freq, gz * freq, 4);
        }
}

// GenerateVoxelChunk (VoxelSceneGen.cu:61-165). Block ids: Sand 1, Soil 2, Cliff 3, Rocks 7.
// Instanced (triangle-mesh) block ids >= 13 are outside this build's scope: the ten
// shader-ball cells at global (30..39, 7, 43) are written as the mesh id by the reference and as
// EMPTY (0) here — the cube of that cell is absent in the reference's voxel mesh as well.
inline uint8_t terrainVoxel(float noiseVal, int y, unsigned width, unsigned gx, unsigned gy, unsigned gz)
{
    uint8_t id = 0;
    float terrainHeight = fmaxr(0.1f, (noiseVal * 1.4f - 0.7f + 0.25f) * width);
    terrainHeight = fminr(terrainHeight, width * 0.9f);
    if (y < terrainHeight)
    {
        float verticalDepth = terrainHeight - y;
        if (terrainHeight < width * (0.25f + 0.05f))
            id = (verticalDepth < 3.5f) ? 1 : 7;
        else if (terrainHeight < width * (0.25f + 0.6f) && terrainHeight > width * (0.25f + 0.3f))
            id = (verticalDepth < 5.5f) ? 3 : 7;
        else
            id = (verticalDepth < 1.5f) ? 2 : (verticalDepth < 5.5f ? 3 : 7);
    }
    if (gy == 7 && gz == 43 && gx >= 30 && gx <= 39) id = 0;
    return id;
}
// Generalisation for chunksY > 1 (cfg5, not exercised by the reference whose chunksY is always 1 and whose
// globalOffsetY is always 0, VoxelSceneGen.cu:381): heights scale with the global height 32*cy and y is the
// global y. For cy == 1 this is exactly the reference kernel.
inline void generateTerrain(Grid &g, int cx, int cy, int cz, const float *noisePerChunk /* nChunks*1024 */)
{
    g.cx = cx; g.cy = cy; g.cz = cz;
    g.ids.assign((size_t)cx * cy * cz * 32768, 0);
    for (int c = 0; c < cx * cy * cz; ++c)
    {
        int chunkX = c % cx, chunkZ = (c / cx) % cz, chunkY = c / (cx * cz);
        const float *noise = noisePerChunk + (size_t)c * 1024;
        for (int y = 0; y < 32; ++y)
            for (int z = 0; z < 32; ++z)
                for (int x = 0; x < 32; ++x)
                    g.ids[(size_t)c * 32768 + x + 32 * (z + 32 * y)] =
                        terrainVoxel(noise[z * 32 + x], chunkY * 32 + y, (unsigned)(32 * cy),
                                     chunkX * 32 + x, chunkY * 32 + y, chunkZ * 32 + z);
    }
}

// ---------------------------------------------------------------- Materials / sky / reservoirs (PODs shared with tests)
struct Material // mirrors MaterialParameter (SystemParameter.h:11-38) without texture handles
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
};
static_assert(sizeof(Material) == 48, "Material POD");

struct AliasBin { float q, p; int32_t alias; }; // AliasTable.h:11-16
struct Reservoir { uint32_t lightData, uvData; float weightSum, targetPdf, M; }; // RestirCommon.h:6-13
static_assert(sizeof(Reservoir) == 20, "DIReservoir POD");

struct Sky
{
    int skyW = 0, skyH = 0, sunW = 0, sunH = 0;
    std::vector<f4> sky, sun;
    std::vector<AliasBin> skyAlias, sunAlias;
    f3 sunDir = {0, 1, 0};
};

// Alias table build (AliasTable.cu:66-153, the live CPU path)
inline void buildAliasTable(const float *weights, unsigned n, AliasBin *bins)
{
    float sum = 0.0f;
    // thrust::reduce order is unspecified; the oracle sums in double and rounds once.
    double acc = 0.0;
    for (unsigned i = 0; i < n; ++i) acc += weights[i];
    sum = (float)acc;
    std::vector<float> prob(n), scaled(n);
    std::vector<int> alias(n, -1);
    for (unsigned i = 0; i < n; ++i) { float p = weights[i] / sum; prob[i] = p; scaled[i] = p * n; }
    std::vector<int> smallQ, largeQ; // FIFO queues
    smallQ.reserve(n); largeQ.reserve(n);
    size_t sh = 0, lh = 0;
    for (unsigned i = 0; i < n; ++i) (scaled[i] < 1.0f ? smallQ : largeQ).push_back((int)i);
    while (sh < smallQ.size() && lh < largeQ.size())
    {
        int s = smallQ[sh++], l = largeQ[lh++];
        alias[s] = l;
        scaled[l] -= (1.0f - scaled[s]);
        (scaled[l] < 1.0f ? smallQ : largeQ).push_back(l);
    }
    while (sh < smallQ.size()) scaled[smallQ[sh++]] = 1.0f;
    while (lh < largeQ.size()) scaled[largeQ[lh++]] = 1.0f;
    for (unsigned i = 0; i < n; ++i) { bins[i].p = prob[i]; bins[i].q = scaled[i]; bins[i].alias = alias[i]; }
}

} // namespace orc
