// Local emissive lights of the voxel world — host side (g++, IEEE arithmetic, no contraction).
//
// Replaces, on this path, the reference's light-list plumbing: generateLightInfosKernel / launchGenerateLightInfos
// (/root/reference/voxelengine/VoxelEngine.cu:53-139: one LightInfo per triangle of every emissive instanced mesh),
// extractRadianceKernel + buildAliasTable (:141-192: alias table over luminance x area), the instance -> light offset table that
// __closesthit__bsdf_light searches (renderer/shaders/closesthit.cu:854-901) and the previous -> current id table of
// LoadDIReservoir (renderer/shaders/Restir.h:60-75). Instanced meshes are outside this build; SURVEY 8f #4 names
// the replacement: the lights are the EXPOSED FACES OF EMISSIVE VOXELS, two triangles per face, voxel order x + W*(z + D*y), face
// order 0..5 (VoxelSceneGen.cu:192-199), triangle order 0, 1 — packed into the reference's own 32-byte LightInfo
// (renderer/shaders/Light.h:13-24, TriangleLight::Store :125-136), so the device side decodes and samples them exactly as the
// reference does (vpt_wave.cu). A scene without emissive voxels has no list and pays nothing.
#include "../csrc/vpt_lights.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace vpt {

// ---- IEEE binary16 <-> binary32, round to nearest even (what __float2half_rn / __half2float do on the device)
static uint32_t halfBits(float f)
{
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x > 0x7f800000u) return sign | 0x7e00u;
    if (x >= 0x477ff000u) return sign | 0x7c00u; // 65520 and above round to infinity
    if (x <= 0x33000000u) return sign;           // 2^-25 and below round to zero (the tie goes to even)
    const int e = (int)(x >> 23) - 127;
    const uint32_t m = (x & 0x7fffffu) | 0x800000u;
    const int shift = e < -14 ? 13 + (-14 - e) : 13;
    uint32_t h = m >> shift;
    const uint32_t rest = m & ((1u << shift) - 1u), tie = 1u << (shift - 1);
    if (rest > tie || (rest == tie && (h & 1u))) ++h;
    return e < -14 ? (sign | h) : (sign | (((uint32_t)(e + 15) << 10) + (h - 0x400u)));
}
static float halfValue(uint32_t h)
{
    const uint32_t sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu, x;
    if (e == 0)
    {
        if (m == 0) x = sign;
        else
        {
            int k = 0;
            while (!(m & 0x400u)) { m <<= 1; ++k; }
            x = sign | ((uint32_t)(113 - k) << 23) | ((m & 0x3ffu) << 13);
        }
    }
    else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

struct V3 { float x, y, z; };
static V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
static float len(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
static float sat(float v) { return std::fmin(std::fmax(v, 0.0f), 1.0f); }
// ndirToOctUnorm32 (LinearMath.h:2104-2123) of a unit vector
static uint32_t octEncode(V3 n)
{
    const float inv = 1.f / (std::fabs(n.x) + std::fabs(n.y) + std::fabs(n.z));
    float px = n.x * inv, py = n.y * inv;
    if (n.z < 0.f)
    {
        const float wx = (1.0f - std::fabs(py)) * (px >= 0.0f ? 1.0f : -1.0f), wy = (1.0f - std::fabs(px)) * (py >= 0.0f ? 1.0f : -1.0f);
        px = wx; py = wy;
    }
    px = sat(px * 0.5f + 0.5f); py = sat(py * 0.5f + 0.5f);
    return (uint32_t)(px * 0xfffe) | ((uint32_t)(py * 0xfffe) << 16);
}
static V3 octDecode(uint32_t u)
{
    float px = sat(float(u & 0xffffu) / 0xfffe) * 2.0f - 1.0f, py = sat(float(u >> 16) / 0xfffe) * 2.0f - 1.0f;
    V3 n = {px, py, 1.0f - std::fabs(px) - std::fabs(py)};
    const float t = std::fmax(0.0f, -n.z);
    n.x += n.x >= 0.0f ? -t : t;
    n.y += n.y >= 0.0f ? -t : t;
    const float l = len(n);
    if (l < 1e-8f || std::isnan(l)) return {0.0f, 0.0f, 1.0f};
    return {n.x / l, n.y / l, n.z / l};
}
// TriangleLight::Store (Light.h:125-136); edges of a voxel face are unit axis vectors
static VptLightInfo packTriangle(V3 base, V3 e1, V3 e2, const float *radiance)
{
    VptLightInfo li;
    std::memset(&li, 0, sizeof li);
    li.radiance[0] = halfBits(radiance[0]) | (halfBits(radiance[1]) << 16);
    li.radiance[1] = halfBits(radiance[2]) | (halfBits(0.0f) << 16);
    const V3 s = add(e1, e2);
    li.center[0] = base.x + s.x / 3.0f; li.center[1] = base.y + s.y / 3.0f; li.center[2] = base.z + s.z / 3.0f;
    const float l1 = len(e1), l2 = len(e2);
    li.direction1 = octEncode({e1.x / l1, e1.y / l1, e1.z / l1});
    li.direction2 = octEncode({e2.x / l2, e2.y / l2, e2.z / l2});
    li.scalars = halfBits(l1) | (halfBits(l2) << 16);
    return li;
}
// luminance(radiance) * surfaceArea of the DECODED light (extractRadianceKernel, VoxelEngine.cu:141-149)
static float lightWeight(const VptLightInfo &li)
{
    const float f0 = halfValue(li.scalars & 0xffffu), f1 = halfValue(li.scalars >> 16);
    const V3 d1 = octDecode(li.direction1), d2 = octDecode(li.direction2);
    const V3 e1 = {d1.x * f0, d1.y * f0, d1.z * f0}, e2 = {d2.x * f1, d2.y * f1, d2.z * f1};
    const V3 n = {e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x};
    const float l = len(n), area = l > 0.0f ? 0.5f * l : 0.0f;
    const float r = halfValue(li.radiance[0] & 0xffffu), g = halfValue(li.radiance[0] >> 16), b = halfValue(li.radiance[1] & 0xffffu);
    return (r * 0.2126f + g * 0.7152f + b * 0.0722f) * area;
}

void faceFrame(int face, int x, int y, int z, float *A, float *u, float *v)
{
    // corner A and in-face axes with u x v = the outward normal (face ids: 0 +y, 1 -y, 2 -x, 3 +x, 4 +z, 5 -z)
    static const float frames[6][9] = {{0, 1, 0, 0, 0, 1, 1, 0, 0}, {0, 0, 0, 1, 0, 0, 0, 0, 1}, {0, 0, 0, 0, 0, 1, 0, 1, 0},
                                       {1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 0, 1, 1, 0, 0, 0, 1, 0}, {0, 0, 0, 0, 1, 0, 1, 0, 0}};
    const float *f = frames[face];
    A[0] = (float)x + f[0]; A[1] = (float)y + f[1]; A[2] = (float)z + f[2];
    u[0] = f[3]; u[1] = f[4]; u[2] = f[5];
    v[0] = f[6]; v[1] = f[7]; v[2] = f[8];
}

void buildLightList(const uint8_t *idsChunk, int cx, int cy, int cz, const VptMaterial *materials, int nMaterials, const uint16_t *blockToMaterial,
                    LightList &out)
{
    out.lights.clear(); out.faceKeys.clear(); out.alias.clear();
    bool any = false;
    for (int m = 0; m < nMaterials; ++m) any = any || materials[m].isEmissive != 0;
    if (!any) return;
    const int W = cx * 32, H = cy * 32, D = cz * 32;
    auto at = [&](int x, int y, int z) -> uint8_t {
        if (x < 0 || y < 0 || z < 0 || x >= W || y >= H || z >= D) return 0;
        const size_t chunk = (size_t)(x >> 5) + (size_t)cx * ((size_t)(z >> 5) + (size_t)cz * (size_t)(y >> 5));
        return idsChunk[chunk * 32768 + (size_t)((x & 31) + 32 * ((z & 31) + 32 * (y & 31)))];
    };
    static const int nb[6][3] = {{0, 1, 0}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 0, 1}, {0, 0, -1}};
    for (int y = 0; y < H; ++y)
        for (int z = 0; z < D; ++z)
            for (int x = 0; x < W; ++x)
            {
                const uint8_t id = at(x, y, z);
                if (id == 0) continue;
                const VptMaterial &mat = materials[blockToMaterial[id]];
                if (!mat.isEmissive) continue;
                const uint32_t lin = (uint32_t)(x + W * (z + D * y));
                for (int f = 0; f < 6; ++f)
                {
                    if (at(x + nb[f][0], y + nb[f][1], z + nb[f][2]) != 0) continue; // a covered face emits nothing visible
                    float A[3], u[3], v[3];
                    faceFrame(f, x, y, z, A, u, v);
                    const V3 a = {A[0], A[1], A[2]}, eu = {u[0], u[1], u[2]}, ev = {v[0], v[1], v[2]};
                    out.faceKeys.push_back((lin << 3) | (uint32_t)f);
                    out.lights.push_back(packTriangle(a, eu, ev, mat.albedo));
                    out.lights.push_back(packTriangle(add(add(a, eu), ev), neg(eu), neg(ev), mat.albedo));
                }
            }
    if (out.lights.empty()) return;
    std::vector<float> w(out.lights.size());
    for (size_t i = 0; i < w.size(); ++i) w[i] = lightWeight(out.lights[i]);
    out.alias.resize(w.size());
    vpt_build_alias_table(w.data(), (unsigned)w.size(), out.alias.data());
}

void buildLightRemap(const std::vector<uint32_t> &prevKeys, const std::vector<uint32_t> &curKeys, std::vector<int> &prevToCur)
{
    prevToCur.assign(prevKeys.size() * 2, -1);
    for (size_t i = 0; i < prevKeys.size(); ++i)
    {
        const auto it = std::lower_bound(curKeys.begin(), curKeys.end(), prevKeys[i]);
        if (it == curKeys.end() || *it != prevKeys[i]) continue;
        const int c = 2 * (int)(it - curKeys.begin());
        prevToCur[2 * i] = c; prevToCur[2 * i + 1] = c + 1;
    }
}

} // namespace vpt
