// ORACLE — test infrastructure only (see orc_math.h header). Local emissive lights.
//
// The reference's local lights are the triangles of emissive INSTANCED meshes (lantern blocks): a LightInfo per triangle
// (/root/reference/voxelengine/VoxelEngine.cu:53-116 generateLightInfosKernel), an alias table over luminance x area
// (:152-192 extractRadianceKernel / buildAliasTable), TriangleLight pack / unpack / sampling
// (/root/reference/renderer/shaders/Light.h:44-137), the helpers of LinearMath.h:2048-2066 (SampleTriangle /
// InverseTriangleSample), :2069-2123 (octahedral unorm32), :2125-2128 (PdfAtoW), :2165-2190 (fp16 packing), the
// instance -> light mapping of __closesthit__bsdf_light (closesthit.cu:854-901) and the previous -> current light id
// remap of LoadDIReservoir (Restir.h:48-79).
//
// Instanced meshes are outside this build; SURVEY 8f #4 names the replacement: the light list is made of
// the EXPOSED FACES OF EMISSIVE VOXELS, two triangles per face, in voxel order x + W*(z + D*y), face order 0..5
// (VoxelSceneGen.cu:192-199), triangle order 0, 1. Everything downstream of the list is the reference's own arithmetic:
// LightInfo packing (fp16 edge lengths and radiance, oct-encoded edge directions, centroid), TriangleLight::Create,
// calcSample, the alias table, the 8 RIS candidates, the BSDF-ray hit -> (light, barycentrics) classification (a sorted
// face-key table searched like instanceLightMapping), finite-tmax visibility rays, and the id remap after an edit.
#pragma once
#include "orc_scene.h"
#include <algorithm>

namespace orc {

struct LightInfo // Light.h:13-24, 32 bytes
{
    float center[3];
    uint32_t scalars;    // 2 x fp16: |edge1|, |edge2|
    uint32_t radiance[2]; // fp16 x 4
    uint32_t direction1, direction2; // oct-encoded unit edge directions
};
static_assert(sizeof(LightInfo) == 32, "LightInfo POD");

// ---- fp16 (IEEE binary16, round to nearest even): __float2half_rn / __half2float
inline uint32_t f32ToF16Bits(float f)
{
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return sign | (x > 0x7f800000u ? 0x7e00u : 0x7c00u); // NaN / inf
    if (x >= 0x477ff000u) return sign | 0x7c00u;                                  // rounds to inf (>= 65520)
    if (x < 0x33000001u) return sign;                                             // < 2^-25 (or == 2^-25: ties to even 0)
    int e = (int)(x >> 23) - 127;
    uint32_t m = (x & 0x7fffffu) | 0x800000u;
    int shift = e < -14 ? 13 + (-14 - e) : 13; // subnormal halves lose more bits
    uint32_t h = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    if (e < -14) return sign | h;                      // subnormal (a carry lands on the smallest normal: still right)
    return sign | (((uint32_t)(e + 15) << 10) + (h - 0x400u)); // h has the implicit bit at 0x400; a mantissa carry bumps the exponent
}
inline float f16BitsToF32(uint32_t h)
{
    const uint32_t sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    uint32_t x;
    if (e == 0)
    {
        if (m == 0) x = sign;
        else
        {
            int k = 0; uint32_t mm = m;
            while (!(mm & 0x400u)) { mm <<= 1; ++k; }
            x = sign | ((uint32_t)(127 - 15 - k + 1) << 23) | ((mm & 0x3ffu) << 13);
        }
    }
    else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f; std::memcpy(&f, &x, 4);
    return f;
}

// ---- octahedral unorm32 (LinearMath.h:2069-2123)
inline f3 octToNdirSigned(f2 p)
{
    f3 n = {p.x, p.y, 1.0f - fabsf(p.x) - fabsf(p.y)};
    const float t = fmaxf(0.0f, -n.z);
    n.x += n.x >= 0.0f ? -t : t;
    n.y += n.y >= 0.0f ? -t : t;
    return normalize(n);
}
inline f3 octToNdirUnorm32(uint32_t u)
{
    f2 p = {saturate(float(u & 0xffffu) / 0xfffe), saturate(float(u >> 16) / 0xfffe)};
    p = {p.x * 2.0f - 1.0f, p.y * 2.0f - 1.0f};
    return octToNdirSigned(p);
}
inline uint32_t ndirToOctUnorm32(f3 n)
{
    const float inv = 1.f / (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
    f2 p = {n.x * inv, n.y * inv};
    if (n.z < 0.f) p = {(1.0f - fabsf(p.y)) * (p.x >= 0.0f ? 1.0f : -1.0f), (1.0f - fabsf(p.x)) * (p.y >= 0.0f ? 1.0f : -1.0f)};
    p = {saturate(p.x * 0.5f + 0.5f), saturate(p.y * 0.5f + 0.5f)};
    return (uint32_t)(p.x * 0xfffe) | ((uint32_t)(p.y * 0xfffe) << 16);
}

struct TriangleLight // Light.h:44-137
{
    f3 base, edge1, edge2, radiance, normal;
    float surfaceArea;
};
inline LightInfo storeTriangleLight(f3 base, f3 edge1, f3 edge2, f3 radiance)
{
    LightInfo li{};
    li.radiance[0] = f32ToF16Bits(radiance.x) | (f32ToF16Bits(radiance.y) << 16);
    li.radiance[1] = f32ToF16Bits(radiance.z) | (f32ToF16Bits(0.0f) << 16);
    const f3 c = base + (edge1 + edge2) / 3.0f;
    li.center[0] = c.x; li.center[1] = c.y; li.center[2] = c.z;
    li.direction1 = ndirToOctUnorm32(normalize(edge1));
    li.direction2 = ndirToOctUnorm32(normalize(edge2));
    li.scalars = f32ToF16Bits(length(edge1)) | (f32ToF16Bits(length(edge2)) << 16);
    return li;
}
inline TriangleLight createTriangleLight(const LightInfo &li)
{
    TriangleLight t;
    const float f0 = f16BitsToF32(li.scalars & 0xffffu), f1 = f16BitsToF32(li.scalars >> 16);
    t.edge1 = octToNdirUnorm32(li.direction1) * f0;
    t.edge2 = octToNdirUnorm32(li.direction2) * f1;
    t.base = F3(li.center[0], li.center[1], li.center[2]) - (t.edge1 + t.edge2) / 3.0f;
    t.radiance = {f16BitsToF32(li.radiance[0] & 0xffffu), f16BitsToF32(li.radiance[0] >> 16), f16BitsToF32(li.radiance[1] & 0xffffu)};
    const f3 n = cross(t.edge1, t.edge2);
    const float len = length(n);
    if (len > 0.0f) { t.surfaceArea = 0.5f * len; t.normal = n / len; }
    else { t.surfaceArea = 0.0f; t.normal = F3(0.0f); }
    return t;
}
inline f3 sampleTriangle(f2 u) { const float s = sqrtf(u.x); return {1.0f - s, s * (1.0f - u.y), s * u.y}; }
inline f2 inverseTriangleSample(f2 hitUV)
{
    const f3 b = {1.0f - hitUV.x - hitUV.y, hitUV.x, hitUV.y};
    const float s = 1 - b.x;
    return {s * s, b.z / s};
}

// ---- the light list of a grid: exposed faces of emissive voxels
// corner A and the in-face axes (u, v) with u x v = outward normal, per face id
inline void faceFrame(int face, int x, int y, int z, f3 &A, f3 &u, f3 &v)
{
    const float fx = (float)x, fy = (float)y, fz = (float)z;
    switch (face)
    {
    case 0: A = {fx, fy + 1.0f, fz}; u = {0, 0, 1}; v = {1, 0, 0}; break; // +y
    case 1: A = {fx, fy, fz}; u = {1, 0, 0}; v = {0, 0, 1}; break;        // -y
    case 2: A = {fx, fy, fz}; u = {0, 0, 1}; v = {0, 1, 0}; break;        // -x
    case 3: A = {fx + 1.0f, fy, fz}; u = {0, 1, 0}; v = {0, 0, 1}; break; // +x
    case 4: A = {fx, fy, fz + 1.0f}; u = {1, 0, 0}; v = {0, 1, 0}; break; // +z
    default: A = {fx, fy, fz}; u = {0, 1, 0}; v = {1, 0, 0}; break;       // -z
    }
}
struct LightList
{
    std::vector<LightInfo> lights;       // 2 per face
    std::vector<uint32_t> faceKeys;      // ascending (linear voxel << 3) | face; light index = 2 * position + triangle
    std::vector<AliasBin> alias;
    float accumulatedLuminance = 0.0f;
};
inline void buildLightList(const Grid &g, const std::vector<Material> &materials, const uint16_t *blockToMaterial, LightList &out)
{
    out.lights.clear(); out.faceKeys.clear(); out.alias.clear(); out.accumulatedLuminance = 0.0f;
    const int W = g.W(), H = g.H(), D = g.D();
    static const int nb[6][3] = {{0, 1, 0}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 0, 1}, {0, 0, -1}};
    bool anyEmissive = false;
    for (size_t m = 0; m < materials.size(); ++m) anyEmissive |= materials[m].isEmissive != 0;
    if (!anyEmissive) return;
    for (int y = 0; y < H; ++y)
        for (int z = 0; z < D; ++z)
            for (int x = 0; x < W; ++x)
            {
                const uint8_t id = g.ids[g.index(x, y, z)];
                if (id == 0) continue;
                const Material &mat = materials[blockToMaterial[id]];
                if (!mat.isEmissive) continue;
                const uint32_t lin = (uint32_t)(x + W * (z + D * y));
                for (int f = 0; f < 6; ++f)
                {
                    if (g.at(x + nb[f][0], y + nb[f][1], z + nb[f][2]) != 0) continue; // covered face
                    f3 A, u, v;
                    faceFrame(f, x, y, z, A, u, v);
                    const f3 rad = {mat.albedo[0], mat.albedo[1], mat.albedo[2]};
                    out.faceKeys.push_back((lin << 3) | (uint32_t)f);
                    out.lights.push_back(storeTriangleLight(A, u, v, rad));
                    out.lights.push_back(storeTriangleLight(A + u + v, -u, -v, rad));
                }
            }
    if (out.lights.empty()) return;
    std::vector<float> w(out.lights.size());
    for (size_t i = 0; i < w.size(); ++i)
    {
        const TriangleLight t = createTriangleLight(out.lights[i]);
        w[i] = luminance(t.radiance) * t.surfaceArea; // extractRadianceKernel
    }
    out.alias.resize(w.size());
    buildAliasTable(w.data(), (unsigned)w.size(), out.alias.data());
    double acc = 0.0;
    for (float x : w) acc += x;
    out.accumulatedLuminance = (float)acc;
}
// light index of (voxel, face, triangle); -1 when the face is not a light
inline int findLight(const LightList &l, uint32_t lin, int face, int tri)
{
    const uint32_t key = (lin << 3) | (uint32_t)face;
    auto it = std::lower_bound(l.faceKeys.begin(), l.faceKeys.end(), key);
    if (it == l.faceKeys.end() || *it != key) return -1;
    return 2 * (int)(it - l.faceKeys.begin()) + tri;
}
// prevLightIdToCurrentId (Restir.h:60-75) after the list changed
inline void buildLightRemap(const std::vector<uint32_t> &prevKeys, const std::vector<uint32_t> &curKeys, std::vector<int> &prevToCur)
{
    prevToCur.assign(prevKeys.size() * 2, -1);
    for (size_t i = 0; i < prevKeys.size(); ++i)
    {
        auto it = std::lower_bound(curKeys.begin(), curKeys.end(), prevKeys[i]);
        if (it == curKeys.end() || *it != prevKeys[i]) continue;
        const int c = 2 * (int)(it - curKeys.begin());
        prevToCur[2 * i] = c; prevToCur[2 * i + 1] = c + 1;
    }
}

} // namespace orc
