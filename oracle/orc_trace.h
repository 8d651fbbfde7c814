// ORACLE — test infrastructure only (see orc_math.h header). CPU restatement of the per-pixel
// path-tracing hot path: ray–voxel DDA, Disney BSDF, RIS + temporal ReSTIR direct light, miss/sky.
//
//   DDA        /root/reference/voxelengine/VoxelEngine.cu:1040-1166 (performRayTraversal), generalised to an
//              arbitrary origin/direction with an entry clip, [tmin,tmax) and face id + t.
//   raygen     /root/reference/renderer/shaders/RayGen.cu:8-181
//   closesthit /root/reference/renderer/shaders/closesthit.cu:96-852 (triangle fetch / SelfHit :20-88 are
//              replaced by exact voxel-face arithmetic)
//   miss       /root/reference/renderer/shaders/miss.cu:9-94
//   BSDF       /root/reference/renderer/shaders/Bsdf.h:12-22, 202-245, 371-617
//   ReSTIR     /root/reference/renderer/shaders/Restir.h (all), AliasTable.h:34-54, Sampler.h:652-698
//
// Deliberate, documented departures from the reference (which is not reproducible at these points):
//  * OptiX triangle traversal -> DDA over the grid (B200 has no RT cores). Face ids follow
//    VoxelSceneGen.cu:192-199: 0 Up(+y) 1 Down(-y) 2 Left(-x) 3 Right(+x) 4 Back(+z) 5 Front(-z).
//  * Spawn point: hit point with the face coordinate snapped exactly onto the integer plane, pushed
//    kSpawnEps = 2^-10 along the geometric normal (replaces SelfIntersectionAvoidance, closesthit.cu:40-73).
//  * Thin-film and the motion vector of animated meshes are outside this build's scenes (static cube voxels;
//    SURVEY §8a S4/S5, §8f): motionWS == 0. Local emissive lights are the exposed faces of emissive voxels
//    (orc_lights.h) instead of the triangles of emissive instanced meshes; from the light list on, the reference's
//    arithmetic (closesthit.cu:330-375, 470-600, 616-626, 736-755, 801-820; Restir.h:48-79, 383-415).
//  * C++ leaves the evaluation order of rand2/rand4's constructor arguments unspecified; nvcc evaluates
//    left to right, which is what is restated here (x = first dimension drawn).
//  * New parameters (SURVEY "five facts" #2): spp (sample index = iterationIndex*spp + k; sample 0 owns the
//    G-buffer and the temporal ReSTIR pass, samples k>0 shade with the per-sample RIS reservoir) and runtime
//    totalBounceLimit / diffuseBounceLimit. spp=1, limits 3/1 is exactly the reference.
#pragma once
#include "orc_scene.h"
#include "orc_lights.h"

namespace orc {

constexpr float kSpawnEps = 0.0009765625f; // 2^-10
constexpr uint32_t kLightValidBit = 0x80000000u, kLightIndexMask = 0x7FFFFFFFu;
constexpr uint32_t kInvalidLight = 0x7FFFFFFFu, kSkyLight = 0x7FFFFFFEu, kSunLight = 0x7FFFFFFDu;
enum { LightInvalid = 0, LightSky = 1, LightSun = 2, LightLocalTriangle = 3 };

// ---------------------------------------------------------------- DDA
struct Hit
{
    int hit;     // 1 = solid voxel found
    int x, y, z; // voxel
    int face;    // 0..5, or 6 when the ray starts inside a solid voxel (no entered face)
    int id;      // block id
    float t;     // ray parameter of the entered face (0 for face 6)
    int steps;   // voxel steps walked (for statistics)
};

inline Hit ddaTrace(const Grid &g, f3 o, f3 d, float tmin, float tmax)
{
    Hit h{};
    h.t = kRayMax;
    const int W = g.W(), H = g.H(), D = g.D();
    int x = (int)std::floor(o.x), y = (int)std::floor(o.y), z = (int)std::floor(o.z);
    int hitAxis = -1;
    float tCur = 0.0f;

    // Entry clip (not in the reference picker, which simply breaks when out of bounds, :1089-1094)
    if (x < 0 || x >= W || y < 0 || y >= H || z < 0 || z >= D)
    {
        const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
        const float dim[3] = {(float)W, (float)H, (float)D};
        float tEnter = -FLT_MAX, tExit = FLT_MAX;
        int axis = -1;
        for (int a = 0; a < 3; ++a)
        {
            if (fabsf(dd[a]) < 1e-8f)
            {
                if (oo[a] < 0.0f || oo[a] >= dim[a]) return h;
                continue;
            }
            float ta = (0.0f - oo[a]) / dd[a], tb = (dim[a] - oo[a]) / dd[a];
            float tn = fminr(ta, tb), tf = fmaxr(ta, tb);
            if (tn > tEnter) { tEnter = tn; axis = a; }
            if (tf < tExit) tExit = tf;
        }
        if (axis < 0 || tEnter > tExit || tExit < 0.0f || tEnter < 0.0f) return h;
        f3 p = o + d * tEnter;
        x = clampi((int)std::floor(p.x), 0, W - 1);
        y = clampi((int)std::floor(p.y), 0, H - 1);
        z = clampi((int)std::floor(p.z), 0, D - 1);
        if (axis == 0) x = dd[0] > 0.0f ? 0 : W - 1;
        if (axis == 1) y = dd[1] > 0.0f ? 0 : H - 1;
        if (axis == 2) z = dd[2] > 0.0f ? 0 : D - 1;
        hitAxis = axis;
        tCur = tEnter;
    }

    const int stepX = (d.x > 0.0f) ? 1 : -1, stepY = (d.y > 0.0f) ? 1 : -1, stepZ = (d.z > 0.0f) ? 1 : -1;
    const float tDeltaX = (fabsf(d.x) < 1e-8f) ? FLT_MAX : (1.0f / fabsf(d.x));
    const float tDeltaY = (fabsf(d.y) < 1e-8f) ? FLT_MAX : (1.0f / fabsf(d.y));
    const float tDeltaZ = (fabsf(d.z) < 1e-8f) ? FLT_MAX : (1.0f / fabsf(d.z));
    const float nbX = (stepX > 0) ? (float)(x + 1) : (float)x;
    const float nbY = (stepY > 0) ? (float)(y + 1) : (float)y;
    const float nbZ = (stepZ > 0) ? (float)(z + 1) : (float)z;
    float tMaxX = (fabsf(d.x) < 1e-8f) ? FLT_MAX : (nbX - o.x) / d.x;
    float tMaxY = (fabsf(d.y) < 1e-8f) ? FLT_MAX : (nbY - o.y) / d.y;
    float tMaxZ = (fabsf(d.z) < 1e-8f) ? FLT_MAX : (nbZ - o.z) / d.z;

    const int maxIter = W + H + D + 4; // always enough to leave the grid (the picker's 1000 never binds for <=128^3)
    for (int it = 0; it < maxIter; ++it)
    {
        if (x < 0 || x >= W || y < 0 || y >= H || z < 0 || z >= D) break;
        if (tCur >= tmax) break;
        uint8_t id = g.ids[g.index(x, y, z)];
        if (id != 0 && tCur >= tmin)
        {
            h.hit = 1; h.x = x; h.y = y; h.z = z; h.id = id; h.t = tCur;
            if (hitAxis == 0) h.face = stepX > 0 ? 2 : 3;
            else if (hitAxis == 1) h.face = stepY > 0 ? 1 : 0;
            else if (hitAxis == 2) h.face = stepZ > 0 ? 5 : 4;
            else h.face = 6;
            return h;
        }
        ++h.steps;
        if (tMaxX < tMaxY)
        {
            if (tMaxX < tMaxZ) { x += stepX; tCur = tMaxX; tMaxX += tDeltaX; hitAxis = 0; }
            else               { z += stepZ; tCur = tMaxZ; tMaxZ += tDeltaZ; hitAxis = 2; }
        }
        else
        {
            if (tMaxY < tMaxZ) { y += stepY; tCur = tMaxY; tMaxY += tDeltaY; hitAxis = 1; }
            else               { z += stepZ; tCur = tMaxZ; tMaxZ += tDeltaZ; hitAxis = 2; }
        }
    }
    return h;
}

inline f3 faceNormal(int face, f3 rayDir)
{
    switch (face)
    {
    case 0: return {0, 1, 0};
    case 1: return {0, -1, 0};
    case 2: return {-1, 0, 0};
    case 3: return {1, 0, 0};
    case 4: return {0, 0, 1};
    case 5: return {0, 0, -1};
    default: // ray started inside a solid voxel: face the ray along its dominant axis
    {
        float ax = fabsf(rayDir.x), ay = fabsf(rayDir.y), az = fabsf(rayDir.z);
        if (ax >= ay && ax >= az) return {rayDir.x > 0 ? -1.0f : 1.0f, 0, 0};
        if (ay >= az) return {0, rayDir.y > 0 ? -1.0f : 1.0f, 0};
        return {0, 0, rayDir.z > 0 ? -1.0f : 1.0f};
    }
    }
}
// Hit point with the entered face's coordinate snapped to its integer plane.
inline f3 hitPoint(const Hit &h, f3 o, f3 d)
{
    f3 p = o + d * h.t;
    switch (h.face)
    {
    case 0: p.y = (float)(h.y + 1); break;
    case 1: p.y = (float)h.y; break;
    case 2: p.x = (float)h.x; break;
    case 3: p.x = (float)(h.x + 1); break;
    case 4: p.z = (float)(h.z + 1); break;
    case 5: p.z = (float)h.z; break;
    default: break;
    }
    return p;
}

// ---------------------------------------------------------------- BSDF (Bsdf.h)
constexpr float kRoughnessThreshold = 0.00001f, kTranslucencyThreshold = 0.001f;
constexpr float kDisneyMinPdf = 1e-5f, kDisneyMaxThroughput = 32.0f, kDisneyMinLobeProb = 0.05f;

inline f3 clampDisneyThroughput(f3 v)
{
    float a = fabsf(luminance(v));
    if (a > kDisneyMaxThroughput && a > 0.0f) return v * (kDisneyMaxThroughput / a);
    return v;
}
inline float fresnelDielectric(float et, float cosIn)
{
    const float cosi = fabsf(cosIn);
    float sint = 1.0f - cosi * cosi;
    sint = (0.0f < sint) ? sqrtf(sint) / et : 0.0f;
    if (1.0f < sint) return 1.0f;
    float cost = 1.0f - sint * sint;
    cost = (0.0f < cost) ? sqrtf(cost) : 0.0f;
    const float et_cosi = et * cosi, et_cost = et * cost;
    const float rPerp = (cosi - et_cost) / (cosi + et_cost);
    const float rPar = (et_cosi - cost) / (et_cosi + cost);
    const float result = (rPar * rPar + rPerp * rPerp) * 0.5f;
    return (result <= 1.0f) ? result : 1.0f;
}
inline float disneyDiffuseFresnel(float cosWo, float cosWi, float roughness)
{
    float energyBias = lerpf(0.0f, 0.5f, roughness);
    float energyFactor = lerpf(1.0f, 1.0f / 1.51f, roughness);
    float fd90 = energyBias + 2.0f * roughness * cosWi * cosWi;
    float f0 = 1.0f;
    float lightScatter = f0 + (fd90 - f0) * pow5(1.0f - cosWo);
    float viewScatter = f0 + (fd90 - f0) * pow5(1.0f - cosWi);
    return lightScatter * viewScatter * energyFactor;
}
inline float gtr2Aniso(float cosH, float sinH, float sinPhi, float cosPhi, float ax, float ay)
{
    float ax2 = ax * ax, ay2 = ay * ay;
    float s = (cosPhi * cosPhi) / ax2 + (sinPhi * sinPhi) / ay2;
    float t = sinH * sinH * s + cosH * cosH;
    return 1.0f / (kPi * ax * ay * t * t);
}
inline float smithGGX(float cosTheta, float alpha)
{
    float a2 = alpha * alpha, c2 = cosTheta * cosTheta;
    return 2.0f / (1.0f + sqrtf(1.0f + a2 * (1.0f - c2) / c2));
}
inline f3 disneyC0(f3 albedo, float metalness)
{
    float lum = 0.299f * albedo.x + 0.587f * albedo.y + 0.114f * albedo.z;
    f3 tint = lum > 0.0f ? albedo / lum : F3(1.0f);
    f3 specularColor = lerp3(F3(1.0f), tint, 0.0f);
    return lerp3(0.08f * 0.5f * specularColor, albedo, metalness);
}
inline float disneySpecularProb(float avgF, float metalness, bool &valid, float &diffuseProb)
{
    float specularWeight = avgF;
    float diffuseWeight = (1.0f - metalness) * (1.0f - avgF);
    float totalWeight = specularWeight + diffuseWeight;
    valid = !(totalWeight < kSafeCosEps);
    if (!valid) { diffuseProb = 0.0f; return 0.0f; }
    float specularProb = specularWeight / totalWeight;
    if (diffuseWeight > kSafeCosEps && specularWeight > kSafeCosEps)
        specularProb = clampf(specularProb, kDisneyMinLobeProb, 1.0f - kDisneyMinLobeProb);
    specularProb = clampf(specularProb, 0.0f, 1.0f);
    diffuseProb = fmaxf(0.0f, 1.0f - specularProb);
    return specularProb;
}

inline void disneySample(f4 u, f3 n, f3 ng, f3 wo, f3 albedo, bool metallic, float translucency, float roughness,
                         f3 &wi, f3 &bsdfOverPdf, float &pdf, bool &transmissive)
{
    if (roughness < kRoughnessThreshold)
    {
        transmissive = false;
        if (translucency < kTranslucencyThreshold)
        {
            wi = reflect3(-wo, n);
            if (dot(wi, n) <= 0.0f || dot(wi, ng) <= 0.0f) { bsdfOverPdf = F3(0.0f); pdf = 0.0f; }
            else { bsdfOverPdf = albedo; pdf = 1.0f; }
            pdf = fmaxf(pdf, kDisneyMinPdf);
            bsdfOverPdf = clampDisneyThroughput(bsdfOverPdf);
        }
        else if (translucency > 1.0f - kTranslucencyThreshold)
        {
            const float ior = 1.4f;
            const bool front = dot(wo, ng) > 0.0f;
            const float eta = front ? ior / 1.0f : 1.0f / ior;
            f3 wr = reflect3(-wo, n), wt;
            float R = 1.0f;
            if (refract(wt, -wo, n, eta)) R = fresnelDielectric(eta, dot(wo, n));
            if (u.x <= R) { wi = wr; pdf = R; }
            else { wi = wt; pdf = 1.0f - R; transmissive = true; }
            bsdfOverPdf = albedo / pdf;
            pdf = fmaxf(pdf, kDisneyMinPdf);
            bsdfOverPdf = clampDisneyThroughput(bsdfOverPdf);
        }
        else { bsdfOverPdf = F3(0.0f); pdf = 0.0f; }
        return;
    }
    transmissive = false;
    const float metalness = metallic ? 1.0f : 0.0f;
    float alpha = fmaxf(roughness * roughness, kRoughnessThreshold);
    float cosWo = fmaxf(kSafeCosEps, dot(n, wo));
    f3 C0 = disneyC0(albedo, metalness);
    f3 F = C0 + (F3(1.0f) - C0) * pow5(1.0f - cosWo);
    float avgF = (F.x + F.y + F.z) / 3.0f;
    bool valid; float diffuseProb;
    float specularProb = disneySpecularProb(avgF, metalness, valid, diffuseProb);
    if (!valid) { bsdfOverPdf = F3(0.0f); pdf = 0.0f; return; }

    if (u.w < specularProb)
    {
        float cosTheta = sqrtf((1.0f - u.x) / (1.0f + (alpha * alpha - 1.0f) * u.x));
        cosTheta = clampf(cosTheta, kSafeCosEps, 1.0f);
        float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
        float phi = kTwoPi * u.y;
        f3 wh = {sinTheta * cosf(phi), sinTheta * sinf(phi), cosTheta};
        alignVector(n, wh);
        wi = normalize(reflect3(-wo, wh));
        if (dot(wi, n) <= 0.0f || dot(wi, ng) <= 0.0f) { bsdfOverPdf = F3(0.0f); pdf = 0.0f; return; }
        float cosWi = dot(wi, n);
        float cosWh = fmaxf(kSafeCosEps, fabsf(dot(wh, n)));
        float cosWoWh = fmaxf(kSafeCosEps, fabsf(dot(wo, wh)));
        float sinWh = sqrtf(fmaxf(0.0f, 1.0f - cosWh * cosWh));
        float Dm = gtr2Aniso(cosWh, sinWh, 0.0f, 1.0f, alpha, alpha);
        f3 Fs = C0 + (F3(1.0f) - C0) * pow5(1.0f - cosWoWh);
        float G = smithGGX(cosWo, alpha) * smithGGX(cosWi, alpha);
        f3 brdf = Fs * Dm * G / (4.0f * cosWo * cosWi);
        float microPdf = Dm * cosWh / (4.0f * cosWoWh);
        microPdf = fmaxf(microPdf, kDisneyMinPdf);
        float wSpec = fmaxf(specularProb, kDisneyMinPdf);
        pdf = microPdf * wSpec;
        pdf = fmaxf(pdf, kDisneyMinPdf);
        bsdfOverPdf = clampDisneyThroughput(brdf * cosWi / pdf);
    }
    else
    {
        float cosTheta = sqrtf(u.x);
        float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
        float phi = kTwoPi * u.y;
        wi = {sinTheta * cosf(phi), sinTheta * sinf(phi), cosTheta};
        alignVector(n, wi);
        if (dot(wi, ng) <= 0.0f) { bsdfOverPdf = F3(0.0f); pdf = 0.0f; return; }
        float cosWi = fmaxf(kSafeCosEps, dot(wi, n));
        float fl = disneyDiffuseFresnel(cosWo, cosWi, roughness);
        f3 diffuseBrdf = albedo * (1.0f - metalness) * fl / kPi;
        float diffusePdf = cosWi / kPi;
        diffusePdf = fmaxf(diffusePdf, kDisneyMinPdf);
        float wDiff = fmaxf(diffuseProb, kDisneyMinPdf);
        pdf = diffusePdf * wDiff;
        pdf = fmaxf(pdf, kDisneyMinPdf);
        bsdfOverPdf = clampDisneyThroughput(diffuseBrdf * cosWi / pdf);
    }
}

inline void disneyEvaluate(f3 n, f3 ng, f3 wi, f3 wo, f3 albedo, bool metallic, float /*translucency*/, float roughness,
                           f3 &bsdf, float &pdf)
{
    bsdf = F3(0.0f);
    if (roughness < kRoughnessThreshold) { pdf = 0.0f; return; }
    if (dot(wo, n) <= 0.0f || dot(wi, n) <= 0.0f || dot(wo, ng) <= 0.0f || dot(wi, ng) <= 0.0f) { pdf = 0.0f; return; }
    const float metalness = metallic ? 1.0f : 0.0f;
    float alpha = fmaxf(roughness * roughness, kRoughnessThreshold);
    float cosWo = dot(wo, n), cosWi = dot(wi, n);
    f3 wh = normalize(wi + wo);
    float cosWh = fmaxf(kSafeCosEps, fabsf(dot(wh, n)));
    float cosWoWh = fmaxf(kSafeCosEps, fabsf(dot(wo, wh)));
    f3 C0 = disneyC0(albedo, metalness);
    f3 F = C0 + (F3(1.0f) - C0) * pow5(1.0f - cosWoWh);
    f3 diffuse = F3(0.0f);
    if (!metallic)
    {
        float fl = disneyDiffuseFresnel(cosWo, cosWi, roughness);
        diffuse = albedo * (1.0f - metalness) * fl / kPi;
    }
    float sinWh = sqrtf(fmaxf(0.0f, 1.0f - cosWh * cosWh));
    float Dm = gtr2Aniso(cosWh, sinWh, 0.0f, 1.0f, alpha, alpha);
    float G = smithGGX(cosWo, alpha) * smithGGX(cosWi, alpha);
    f3 specular = F * Dm * G / (4.0f * cosWo * cosWi);
    bsdf = clampDisneyThroughput(diffuse + specular);
    float avgF = (F.x + F.y + F.z) / 3.0f;
    bool valid; float diffuseProb;
    float specularProb = disneySpecularProb(avgF, metalness, valid, diffuseProb);
    if (!valid) { pdf = 0.0f; return; }
    float diffusePdf = fmaxf(cosWi / kPi, kDisneyMinPdf);
    float specularPdf = fmaxf(Dm * cosWh / (4.0f * cosWoWh), kDisneyMinPdf);
    float wSpec = fmaxf(specularProb, kDisneyMinPdf), wDiff = fmaxf(diffuseProb, kDisneyMinPdf);
    pdf = diffusePdf * wDiff + specularPdf * wSpec;
    pdf = fmaxf(pdf, kDisneyMinPdf);
}

// ---------------------------------------------------------------- frame state
struct GBufferSet
{
    std::vector<float> depth, material;
    std::vector<f4> normalRoughness, geoNormalThinfilm, materialParameter, albedo;
    void resize(size_t n)
    {
        depth.assign(n, 0.0f); material.assign(n, 0.0f);
        normalRoughness.assign(n, F4(0.0f)); geoNormalThinfilm.assign(n, F4(0.0f));
        materialParameter.assign(n, F4(0.0f)); albedo.assign(n, F4(0.0f));
    }
};
struct TraceParams { int spp = 1, totalBounceLimit = 3, diffuseBounceLimit = 1, enableRestir = 1; };

// Textured materials (closesthit.cu:166-254; sampler state TextureManager.cu:222-240: wrap addressing, linear filter, linear
// mip filter, normalised coordinates, no sRGB, levels down to 4x4: maxLod = log2(width) - 2). The reference feeds BC7/BC5/BC4
// blocks encoded by NVTT to the texture unit; neither the encoder nor the unit's 8-bit filter weights are reproducible, so the
// restatement (and the CUDA path) filter UNCOMPRESSED RGBA8 mip chains supplied by the caller in software, in fp32.
struct Texture
{
    int width = 0, levels = 0;
    std::vector<uint32_t> texels;      // all levels, level l is (width >> l)^2 texels, RGBA8 little endian (r = low byte)
    std::vector<size_t> levelOffset;
};
struct MaterialTextures { int albedo = -1, normal = -1, roughness = -1, metallic = -1; float texSizeX = 1024.0f, texSizeY = 1024.0f; };
inline f4 texelRGBA(const Texture &t, int level, int x, int y)
{
    const int n = t.width >> level;
    x = ((x % n) + n) % n; y = ((y % n) + n) % n;    // cudaAddressModeWrap
    const uint32_t v = t.texels[t.levelOffset[level] + (size_t)y * n + x];
    return {(float)(v & 0xff) / 255.0f, (float)((v >> 8) & 0xff) / 255.0f, (float)((v >> 16) & 0xff) / 255.0f, (float)(v >> 24) / 255.0f};
}
inline f4 texBilinear(const Texture &t, int level, float u, float v)
{
    const int n = t.width >> level;
    const float x = u * n - 0.5f, y = v * n - 0.5f;
    const float fx = std::floor(x), fy = std::floor(y);
    const float a = x - fx, b = y - fy;
    const int i = (int)fx, j = (int)fy;
    const f4 t00 = texelRGBA(t, level, i, j), t10 = texelRGBA(t, level, i + 1, j), t01 = texelRGBA(t, level, i, j + 1), t11 = texelRGBA(t, level, i + 1, j + 1);
    return ((1.0f - a) * (1.0f - b)) * t00 + (a * (1.0f - b)) * t10 + ((1.0f - a) * b) * t01 + (a * b) * t11;
}
// tex2DLod with trilinear filtering
inline f4 tex2DLod(const Texture &t, float u, float v, float lod)
{
    const float maxLod = (float)(t.levels - 1);
    lod = lod < 0.0f ? 0.0f : (lod > maxLod ? maxLod : lod);   // also maps -inf / NaN-free inputs into range
    if (!(lod >= 0.0f)) lod = 0.0f;
    const int l0 = (int)std::floor(lod);
    const int l1 = l0 + 1 < t.levels ? l0 + 1 : l0;
    const float beta = lod - (float)l0;
    const f4 c0 = texBilinear(t, l0, u, v), c1 = texBilinear(t, l1, u, v);
    return c0 + beta * (c1 - c0);
}

struct Scene
{
    int width = 0, height = 0;
    Tables tables;
    Grid grid;
    Grid prevGrid; bool havePrevGrid = false; // snapshot taken by the first edit after a render (the reference's prevTopObject)
    std::vector<Material> materials;
    std::vector<Texture> textures;
    std::vector<MaterialTextures> matTex; // empty or one entry per material
    uint16_t blockToMaterial[256] = {0};
    Sky sky;
    TraceParams tp;
    // frame buffers
    GBufferSet gb[2];
    int cur = 1; // toggled at the start of every render; first frame renders into set 0
    std::vector<f4> illumination;         // (radiance rgb, primary hit distance)
    std::vector<Reservoir> reservoirs;    // 2 * W*H, parity = iterationIndex & 1
    std::vector<int32_t> primaryHits;     // x,y,z,face per pixel (face -1 = miss), sample 0
    uint64_t rayCount = 0, stepCount = 0; // statistics of the last render
    // local emissive lights (orc_lights.h): rebuilt by the first render after the grid or the materials changed
    LightList lights;
    std::vector<int> prevLightToCur; int prevNumLights = 0;
    std::vector<uint32_t> renderedKeys; // face keys of the list the last render used
    bool lightsStale = true;       // grid / materials changed since the list was built
    bool lightsStateDirty = false; // this render's previous-frame reservoirs hold ids of the previous list (Restir.h:52)
};
// the list as the next render will see it
inline void refreshLights(Scene &sc)
{
    if (!sc.lightsStale) return;
    sc.lights = LightList();
    buildLightList(sc.grid, sc.materials, sc.blockToMaterial, sc.lights);
    sc.lightsStale = false;
}
// at the start of a render (single-threaded): if the list differs from the one the previous render used, the stored reservoirs
// hold stale light ids -> previous -> current id table (Restir.h:60-75), lightsStateDirty for this frame
inline void prepareLightRemap(Scene &sc)
{
    refreshLights(sc);
    sc.lightsStateDirty = sc.renderedKeys != sc.lights.faceKeys;
    if (!sc.lightsStateDirty) return;
    buildLightRemap(sc.renderedKeys, sc.lights.faceKeys, sc.prevLightToCur);
    sc.prevNumLights = (int)sc.renderedKeys.size() * 2;
    sc.renderedKeys = sc.lights.faceKeys;
}

struct Surface
{
    f3 pos; float depth; bool isThinfilm; int materialId;
    f3 normal, geoNormal, albedo, wo; float roughness; bool metallic; float translucency;
};
struct LightSample { f3 position{0, 0, 0}, normal{0, 0, 0}, radiance{0, 0, 0}; float solidAnglePdf = 0; int lightType = LightInvalid; };

inline Reservoir emptyReservoir() { return {0, 0, 0.0f, 0.0f, 0.0f}; }
inline bool isValidReservoir(const Reservoir &r) { return r.lightData != 0; }

struct PixelCtx
{
    const Scene *sc; const Camera *cam, *prevCam;
    int px, py, iterationIndex, sampleIndex, prevSampleIndex;
    int randIdx;
    uint64_t rays, steps;
    float rnd() { return blueNoiseRand(sc->tables, px, py, sampleIndex, randIdx++); }
    float rndPrev(int &idx) const { return blueNoiseRand(sc->tables, px, py, prevSampleIndex, idx++); }
    f2 rnd2() { float a = rnd(); float b = rnd(); return {a, b}; }
    f4 rnd4() { float a = rnd(); float b = rnd(); float c = rnd(); float d = rnd(); return {a, b, c, d}; }
    float rnd16() { f2 u = rnd2(); return u.x + u.y / 256.0f; }
    Hit trace(f3 o, f3 d, float tmin, float tmax)
    {
        Hit h = ddaTrace(sc->grid, o, d, tmin, tmax);
        ++rays; steps += (uint64_t)h.steps;
        return h;
    }
    // against the world as the previous render saw it (usePrevBvh / sysParam.prevTopObject, closesthit.cu:736-755)
    Hit tracePrev(f3 o, f3 d, float tmin, float tmax)
    {
        Hit h = ddaTrace(sc->havePrevGrid ? sc->prevGrid : sc->grid, o, d, tmin, tmax);
        ++rays; steps += (uint64_t)h.steps;
        return h;
    }
};

inline f4 loadClamp(const std::vector<f4> &buf, int w, int h, int x, int y)
{
    x = clampi(x, 0, w - 1); y = clampi(y, 0, h - 1);
    return buf[(size_t)y * w + x];
}
// AliasTable::sample (AliasTable.h:34-50)
inline unsigned aliasSample(const std::vector<AliasBin> &bins, float u, float &pmf)
{
    int len = (int)bins.size();
    int offset = std::min(int(u * len), int(len - 1));
    float up = fminr(u * len - offset, 0.999999f);
    if (up < bins[offset].q) { pmf = bins[offset].p; return (unsigned)offset; }
    int alias = bins[offset].alias;
    pmf = bins[alias].p;
    return (unsigned)alias;
}
inline float sunCosThetaMax() { return cosf(0.51f * kPi / 180.0f / 2.0f); }

inline LightSample createSunLightSample(const Sky &s, int idx)
{
    int ix = idx % s.sunW, iy = idx / s.sunW;
    f2 uv = {(ix + 0.5f) / float(s.sunW), (iy + 0.5f) / float(s.sunH)};
    const float cmax = sunCosThetaMax();
    LightSample ls;
    ls.solidAnglePdf = (s.sunW * s.sunH) / (kTwoPi * (1.0f - cmax));
    ls.position = equalAreaMapCone(s.sunDir, uv.x, uv.y, cmax);
    ls.radiance = xyz(loadClamp(s.sun, s.sunW, s.sunH, ix, iy));
    ls.lightType = LightSun;
    return ls;
}
inline LightSample createSkyLightSample(const Sky &s, int idx)
{
    int ix = idx % s.skyW, iy = idx / s.skyW;
    f2 uv = {(ix + 0.5f) / float(s.skyW), (iy + 0.5f) / float(s.skyH)};
    LightSample ls;
    ls.solidAnglePdf = (s.skyW * s.skyH) / (4.0f * kPi);
    ls.position = equalAreaSphereMap(uv.x, uv.y);
    ls.radiance = xyz(loadClamp(s.sky, s.skyW, s.skyH, ix, iy));
    ls.lightType = LightSky;
    return ls;
}
inline float surfaceBrdfPdf(const Surface &s, f3 wi)
{
    f3 f; float pdf;
    disneyEvaluate(s.normal, s.geoNormal, wi, s.wo, s.albedo, s.metallic, s.translucency, s.roughness, f, pdf);
    return pdf;
}
inline float targetPdfForSurface(const LightSample &ls, const Surface &s)
{
    if (ls.solidAnglePdf <= 0 || ls.lightType == LightInvalid) return 0.0f;
    f3 wi = (ls.lightType == LightLocalTriangle) ? normalize(ls.position - s.pos) : ls.position;
    f3 f; float pdf;
    disneyEvaluate(s.normal, s.geoNormal, wi, s.wo, s.albedo, s.metallic, s.translucency, s.roughness, f, pdf);
    f3 refl = ls.radiance * f * fabsf(dot(wi, s.normal)) / ls.solidAnglePdf;
    return luminance(refl);
}
// LightBrdfMisWeight (Restir.h:286-328) with brdfCutoff == 0 (closesthit.cu:344)
inline float lightBrdfMisWeight(const Surface &s, const LightSample &ls, float lightSelectionPdf, float lightMisWeight,
                                bool /*isEnv*/, float brdfMisWeight)
{
    float lpdf = ls.solidAnglePdf;
    if (brdfMisWeight == 0.0f || lpdf <= 0.0f || std::isinf(lpdf) || std::isnan(lpdf)) return lightMisWeight * lightSelectionPdf;
    f3 lightDir;
    if (ls.lightType == LightSky || ls.lightType == LightSun) lightDir = ls.position;
    else { f3 toLight = ls.position - s.pos; float dist = length(toLight); lightDir = toLight / dist; }
    float brdfPdf = surfaceBrdfPdf(s, lightDir); // maxDistance = FLT_MAX: never shortened
    float sourcePdfWrtSolidAngle = lightSelectionPdf * lpdf;
    float blended = lightMisWeight * sourcePdfWrtSolidAngle + brdfMisWeight * brdfPdf;
    return blended / lpdf;
}
inline bool streamSample(Reservoir &r, uint32_t lightIndex, f2 uv, float random, float targetPdf, float invSourcePdf)
{
    float risWeight = targetPdf * invSourcePdf;
    r.M += 1;
    r.weightSum += risWeight;
    bool sel = (random * r.weightSum < risWeight);
    if (sel)
    {
        r.lightData = lightIndex | kLightValidBit;
        r.uvData = (uint32_t)(saturate(uv.x) * 0xffff) | ((uint32_t)(saturate(uv.y) * 0xffff) << 16);
        r.targetPdf = targetPdf;
    }
    return sel;
}
inline bool combineReservoirs(Reservoir &r, const Reservoir &nr, float random, float targetPdf)
{
    float risWeight = targetPdf * (nr.weightSum * nr.M);
    r.M += nr.M;
    r.weightSum += risWeight;
    bool sel = (random * r.weightSum < risWeight);
    if (sel) { r.lightData = nr.lightData; r.uvData = nr.uvData; r.targetPdf = targetPdf; }
    return sel;
}
inline void finalizeResampling(Reservoir &r, float num, float den)
{
    float d = r.targetPdf * den;
    r.weightSum = (d == 0.0f) ? 0.0f : (r.weightSum * num) / d;
}
inline f2 reservoirUV(const Reservoir &r) { return {float(r.uvData & 0xffff) / float(0xffff), float(r.uvData >> 16) / float(0xffff)}; }

// TriangleLight::calcSample / calcSolidAnglePdf (Light.h:54-82)
inline LightSample triangleLightSample(const TriangleLight &t, f2 random, f3 viewerPosition)
{
    LightSample r;
    const f3 bary = sampleTriangle(random);
    r.position = t.base + t.edge1 * bary.y + t.edge2 * bary.z;
    r.normal = t.normal;
    f3 L = r.position - viewerPosition;
    const float Ldist = length(L);
    L = L / Ldist;
    const float areaPdf = 1.0f / t.surfaceArea;
    const float cosTheta = saturate(dot(L, -r.normal));
    r.solidAnglePdf = areaPdf * (Ldist * Ldist) / cosTheta; // PdfAtoW
    r.radiance = t.radiance;
    r.lightType = LightLocalTriangle;
    return r;
}
// direction and far end of a visibility ray towards a light sample (closesthit.cu:616-617, 743-744, 801-802)
inline f3 lightRayDir(const LightSample &ls, f3 from) { return ls.lightType == LightLocalTriangle ? normalize(ls.position - from) : ls.position; }
inline float lightRayTmax(const LightSample &ls, f3 from, float extraRayOffset)
{
    return ls.lightType == LightLocalTriangle ? length(ls.position - from) - 0.01f - extraRayOffset : kRayMax;
}
// LoadDIReservoir's id remap (Restir.h:48-79)
inline Reservoir remapReservoir(const Scene &sc, Reservoir r)
{
    if (!sc.lightsStateDirty) return r;
    const uint32_t prevIdx = r.lightData & kLightIndexMask;
    if (prevIdx >= kSunLight) return r;
    if (sc.prevNumLights > 0 && prevIdx < (uint32_t)sc.prevNumLights)
    {
        const int curIdx = sc.prevLightToCur[prevIdx];
        if (curIdx < 0 || curIdx >= (int)sc.lights.lights.size()) return emptyReservoir();
        r.lightData = (r.lightData & ~kLightIndexMask) | (uint32_t)curIdx;
    }
    return r;
}
inline bool lightSampleFromReservoir(const Scene &sc, LightSample &ls, const Reservoir &r, const Surface &surface, bool hasLocal)
{
    const Sky &sky = sc.sky;
    uint32_t li = r.lightData & kLightIndexMask;
    f2 uv = reservoirUV(r);
    if (li == kSkyLight)
    {
        int x = clampi(int(uv.x * sky.skyW), 0, sky.skyW - 1), y = clampi(int(uv.y * sky.skyH), 0, sky.skyH - 1);
        ls = createSkyLightSample(sky, y * sky.skyW + x);
    }
    else if (li == kSunLight)
    {
        int x = clampi(int(uv.x * sky.sunW), 0, sky.sunW - 1), y = clampi(int(uv.y * sky.sunH), 0, sky.sunH - 1);
        ls = createSunLightSample(sky, y * sky.sunW + x);
    }
    else if (hasLocal && li < (uint32_t)sc.lights.lights.size())
    {
        ls = triangleLightSample(createTriangleLight(sc.lights.lights[li]), uv, surface.pos);
        return true;
    }
    return li < kInvalidLight;
}
inline i2 clampSamplePositionIntoView(i2 p, int width, int height)
{
    if (p.x < 0) p.x = -p.x;
    if (p.y < 0) p.y = -p.y;
    if (p.x >= width) p.x = 2 * width - p.x - 1;
    if (p.y >= height) p.y = 2 * height - p.y - 1;
    return p;
}
// GetPrevSurface (Restir.h:348-381). The previous-frame jitter is drawn for the CURRENT launch pixel
// (randPrev uses optixGetLaunchIndex), not for `p` — restated as is.
inline bool getPrevSurface(const PixelCtx &c, Surface &s, i2 p)
{
    const Scene &sc = *c.sc;
    const Camera &pc = *c.prevCam;
    if (p.x < 0 || p.y < 0 || p.x >= pc.resolution.x || p.y >= pc.resolution.y) return false;
    const GBufferSet &g = sc.gb[sc.cur ^ 1];
    size_t i = (size_t)p.y * sc.width + p.x;
    s.depth = g.depth[i];
    if (s.depth == kRayMax) return false;
    f4 nr = g.normalRoughness[i], gt = g.geoNormalThinfilm[i], mp = g.materialParameter[i];
    float material = g.material[i];
    int prevSeed = 0;
    float j0 = c.rndPrev(prevSeed), j1 = c.rndPrev(prevSeed);
    f2 prevUV = {(float(p.x) + j0) * pc.inversedResolution.x, (float(p.y) + j1) * pc.inversedResolution.y};
    f3 viewDir = uvToWorldDirection(pc, prevUV);
    s.pos = pc.pos + viewDir * s.depth;
    s.isThinfilm = (gt.w == 1.0f);
    s.materialId = (int)material;
    s.wo = -viewDir;
    s.normal = xyz(nr);
    s.geoNormal = xyz(gt);
    s.albedo = xyz(g.albedo[i]);
    s.roughness = nr.w;
    s.metallic = (mp.x == 1.0f);
    s.translucency = mp.y;
    return true;
}

// ---------------------------------------------------------------- per-ray state (OptixShaderCommon.h:14-38)
struct RayData
{
    f3 pos; float distance; f3 wo, wi; unsigned depth; f3 radiance, bsdfOverPdf; float pdf;
    bool hitFirstDiffuseSurface, shouldTerminate, isCurrentBounceDiffuse, isLastBounceDiffuse, hitFrontFace, transmissionEvent;
    float rayConeWidth, rayConeSpread;
};

// __miss__radiance (miss.cu:9-82)
inline void missRadiance(PixelCtx &c, RayData &rd, bool writeGBuffer)
{
    Scene &sc = const_cast<Scene &>(*c.sc);
    const Sky &sky = sc.sky;
    size_t pix = (size_t)c.py * sc.width + c.px;
    if (rd.depth == 0 && writeGBuffer)
    {
        sc.reservoirs[pix + (size_t)(c.iterationIndex & 1) * sc.width * sc.height] = emptyReservoir();
        GBufferSet &g = sc.gb[sc.cur];
        g.albedo[pix] = F4(1.0f);
        g.material[pix] = (float)0xFFFF;
        g.normalRoughness[pix] = {0.0f, -1.0f, 0.0f, 0.0f};
        g.geoNormalThinfilm[pix] = {0.0f, -1.0f, 0.0f, 0.0f};
        g.materialParameter[pix] = {0, 0, 0, 0};
    }
    f3 emission = F3(0.0f);
    const f3 rayDir = rd.wi;
    f2 uv = equalAreaSphereMapInv(rayDir);
    {
        // SampleBicubicSmoothStep<..., BoundaryFuncRepeatXClampY> (Sampler.h:652-698, :252-278)
        f2 UV = {uv.x * sky.skyW, uv.y * sky.skyH};
        f2 tc = {std::floor(UV.x - 0.5f) + 0.5f, std::floor(UV.y - 0.5f) + 0.5f};
        f2 f = UV - tc;
        f2 f2_ = f * f, f3_ = f2_ * f;
        f2 w1 = {-2.0f * f3_.x + 3.0f * f2_.x, -2.0f * f3_.y + 3.0f * f2_.y};
        f2 w0 = {1.0f - w1.x, 1.0f - w1.y};
        int tx0 = (int)std::floor(UV.x - 0.5f), ty0 = (int)std::floor(UV.y - 0.5f);
        int xs[4] = {tx0, tx0 + 1, tx0, tx0 + 1}, ys[4] = {ty0, ty0, ty0 + 1, ty0 + 1};
        float ws[4] = {w0.x * w0.y, w1.x * w0.y, w0.x * w1.y, w1.x * w1.y};
        f3 out = F3(0.0f);
        float sumW = 0;
        for (int i = 0; i < 4; ++i)
        {
            int x = xs[i], y = ys[i];
            if (x >= sky.skyW) x %= sky.skyW;
            if (x < 0) x = sky.skyW - (-x) % sky.skyW;
            if (y >= sky.skyH) y = sky.skyH - 1;
            if (y < 0) y = 0;
            sumW += ws[i];
            out += xyz(loadClamp(sky.sky, sky.skyW, sky.skyH, x, y)) * ws[i];
        }
        out /= sumW;
        emission += out;
    }
    if (equalAreaMapConeInv(uv, sky.sunDir, rayDir, sunCosThetaMax()))
    {
        int sx = (int)(uv.x * sky.sunW), sy = (int)(uv.y * sky.sunH);
        if (sx >= sky.sunW) sx %= sky.sunW;
        if (sx < 0) sx = sky.sunW - (-sx) % sky.sunW;
        emission += xyz(loadClamp(sky.sun, sky.sunW, sky.sunH, sx, sy));
    }
    rd.radiance = emission;
    rd.distance = kRayMax;
    rd.shouldTerminate = true;
}

// __closesthit__radiance (closesthit.cu:96-852) for a voxel-face hit.
inline void closestHit(PixelCtx &c, RayData &rd, const Hit &h, f3 rayOrig, bool ownsGBuffer)
{
    Scene &sc = const_cast<Scene &>(*c.sc);
    const Sky &sky = sc.sky;
    const size_t pix = (size_t)c.py * sc.width + c.px;
    const size_t npix = (size_t)sc.width * sc.height;
    GBufferSet &g = sc.gb[sc.cur];
    const bool gbufferPass = ownsGBuffer && rd.depth == 0;

    rd.distance = h.t;
    const f3 geoNormal = faceNormal(h.face, rd.wi);
    const f3 surfPos = hitPoint(h, rayOrig, rd.wi);
    const f3 frontPos = surfPos + geoNormal * kSpawnEps;
    // motionWS == 0 (static voxels): the motion vector buffer is identically zero and is elided.
    rd.pos = frontPos;
    const bool hitFrontFace = dot(rd.wo, geoNormal) > 0.0f;
    const Material &mat = sc.materials[sc.blockToMaterial[h.id]];

    if (mat.isEmissive)
    {
        if (!rd.hitFirstDiffuseSurface)
        {
            rd.radiance = {mat.albedo[0], mat.albedo[1], mat.albedo[2]};
            if (ownsGBuffer)
            { // the reference writes these at any depth (closesthit.cu:113-117)
                g.albedo[pix] = F4(1.0f);
                g.material[pix] = (float)0xFFFF;
                g.normalRoughness[pix] = {0.0f, -1.0f, 0.0f, 0.0f};
                g.geoNormalThinfilm[pix] = {0.0f, -1.0f, 0.0f, 0.0f};
                g.materialParameter[pix] = {0, 0, 0, 0};
            }
        }
        rd.shouldTerminate = true;
        return;
    }
    rd.hitFrontFace = hitFrontFace;

    Surface s;
    s.geoNormal = geoNormal;
    s.wo = rd.wo;
    // texture coordinates: world-grid UV by the dominant normal axis (closesthit.cu:166-187), ray-cone LOD (:194-200)
    const size_t matIndex = sc.blockToMaterial[h.id];
    const MaterialTextures *mt = matIndex < sc.matTex.size() ? &sc.matTex[matIndex] : nullptr;
    f2 texCoords = {0.0f, 0.0f};
    if (mat.useWorldGridUV)
    {
        if (fabsf(geoNormal.x) > 0.9f) texCoords = {fmodf(rd.pos.z, mat.uvScale), fmodf(rd.pos.y, mat.uvScale)};
        else if (fabsf(geoNormal.y) > 0.9f) texCoords = {fmodf(rd.pos.x, mat.uvScale), fmodf(rd.pos.z, mat.uvScale)};
        else if (fabsf(geoNormal.z) > 0.9f) texCoords = {fmodf(rd.pos.x, mat.uvScale), fmodf(rd.pos.y, mat.uvScale)};
    }
    rd.rayConeWidth += rd.rayConeSpread * rd.distance;
    texCoords = {texCoords.x / mat.uvScale, texCoords.y / mat.uvScale};
    float lod = 0.0f;
    if (mt)
    {
        const float texMip0Size = sqrtf(mt->texSizeX * mt->texSizeX + mt->texSizeY * mt->texSizeY);
        lod = log2f(rd.rayConeWidth / fmaxr(dot(geoNormal, rd.wo), 0.2f) / mat.uvScale * 2.0f * texMip0Size) - 3.0f;
    }
    s.albedo = F3(mat.albedo[0], mat.albedo[1], mat.albedo[2]);
    if (mt && mt->albedo >= 0) s.albedo = s.albedo * xyz(tex2DLod(sc.textures[mt->albedo], texCoords.x, texCoords.y, lod));
    s.albedo = max3f(s.albedo, F3(0.001f));
    s.roughness = mat.roughness;
    if (mt && mt->roughness >= 0) s.roughness = tex2DLod(sc.textures[mt->roughness], texCoords.x, texCoords.y, lod).x;
    if (rd.hitFirstDiffuseSurface) s.roughness = fminr(s.roughness * 2.0f + 0.1f, 1.0f);
    const bool isDiffuse = s.roughness > kRoughnessThreshold;
    s.metallic = mat.metallic != 0;
    if (mt && mt->metallic >= 0) s.metallic = tex2DLod(sc.textures[mt->metallic], texCoords.x, texCoords.y, lod).x > 0.5f;
    s.translucency = mat.translucency;
    if (mt && mt->normal >= 0)
    {
        const f3 texNormal = xyz(tex2DLod(sc.textures[mt->normal], texCoords.x, texCoords.y, lod));
        f3 nm = normalize(texNormal - F3(0.5f));
        nm.x = -nm.x; nm.y = -nm.y;
        alignVector(geoNormal, nm);
        s.normal = lerp3(geoNormal, nm, 0.2f); // normalMapStrength (closesthit.cu:253-254)
    }
    else
        s.normal = lerp3(geoNormal, geoNormal, 0.2f); // no normal map: state.normal = geoNormal (closesthit.cu:251-254)
    rd.isCurrentBounceDiffuse = isDiffuse;

    if (gbufferPass)
    {
        g.material[pix] = (float)mat.materialId;
        g.normalRoughness[pix] = F4(s.normal, s.roughness);
        g.geoNormalThinfilm[pix] = F4(s.normal, 0.0f);
        g.materialParameter[pix] = {s.metallic ? 1.0f : 0.0f, s.translucency, 0.0f, 0.0f};
    }

    f3 bsdfWi, bsdfOverPdf; float bsdfPdf;
    rd.transmissionEvent = false;
    disneySample(c.rnd4(), s.normal, s.geoNormal, s.wo, s.albedo, s.metallic, s.translucency, s.roughness,
                 bsdfWi, bsdfOverPdf, bsdfPdf, rd.transmissionEvent);
    if (bsdfPdf <= 0.0f) rd.shouldTerminate = true;
    rd.pos = frontPos;
    rd.wi = bsdfWi;
    rd.bsdfOverPdf = bsdfOverPdf; // NOT demodulated: closesthit.cu:301 divides a dead local copy
    rd.pdf = bsdfPdf;

    bool skipAlbedoInShadowRay = false;
    if (rd.depth == 0)
    {
        rd.hitFirstDiffuseSurface = true;
        if (ownsGBuffer) g.albedo[pix] = F4(s.albedo, 1.0f);
        skipAlbedoInShadowRay = true;
    }
    const bool enableReSTIR = sc.tp.enableRestir && rd.depth == 0 && ownsGBuffer;
    Reservoir *storeSlot = (rd.depth == 0 && ownsGBuffer) ? &sc.reservoirs[pix + (size_t)(c.iterationIndex & 1) * npix] : nullptr;

    if (!isDiffuse)
    {
        if (storeSlot) *storeSlot = emptyReservoir();
        return;
    }

    s.materialId = mat.materialId;
    s.pos = rd.pos;
    s.depth = rd.distance;
    s.isThinfilm = false;

    LightSample lightSample;
    Reservoir ris = emptyReservoir();
    const bool skipSun = (dot(s.normal, sky.sunDir) < 0.0f || dot(s.geoNormal, sky.sunDir) < 0.0f);
    const LightList &ll = sc.lights;
    const int numLights = (int)ll.lights.size();
    const int nLocal = numLights > 0 ? 8 : 0; // closesthit.cu:330
    const int nSun = skipSun ? 0 : 1, nSky = 1, nBrdf = 1;
    const int nMis = nLocal + nSun + nSky + nBrdf;
    const float localMisW = float(nLocal) / nMis, sunMisW = float(nSun) / nMis, skyMisW = float(nSky) / nMis, brdfMisW = float(nBrdf) / nMis;

    Reservoir localRes = emptyReservoir();
    LightSample localSample;
    for (int i = 0; i < nLocal; ++i) // closesthit.cu:350-375
    {
        float sourcePdf;
        const int lightIndex = (int)aliasSample(ll.alias, c.rnd(), sourcePdf);
        if (lightIndex >= numLights) continue;
        const f2 uv = c.rnd2();
        const LightSample cand = triangleLightSample(createTriangleLight(ll.lights[lightIndex]), uv, s.pos);
        const float blended = lightBrdfMisWeight(s, cand, sourcePdf, localMisW, false, brdfMisW);
        const float targetPdf = targetPdfForSurface(cand, s);
        const float risRnd = c.rnd();
        if (blended != 0.0f)
            if (streamSample(localRes, (uint32_t)lightIndex, uv, risRnd, targetPdf, 1.0f / blended)) localSample = cand;
    }
    finalizeResampling(localRes, 1.0f, (float)nMis);
    localRes.M = 1;

    Reservoir sunRes = emptyReservoir();
    LightSample sunSample;
    for (int i = 0; i < nSun; ++i)
    {
        float sourcePdf;
        int idx = (int)aliasSample(sky.sunAlias, c.rnd(), sourcePdf);
        LightSample cand = createSunLightSample(sky, idx);
        int ix = idx % sky.sunW, iy = idx / sky.sunW;
        f2 uv = {(ix + 0.5f) / float(sky.sunW), (iy + 0.5f) / float(sky.sunH)};
        float blended = lightBrdfMisWeight(s, cand, sourcePdf, sunMisW, true, brdfMisW);
        float targetPdf = targetPdfForSurface(cand, s);
        float risRnd = c.rnd();
        if (streamSample(sunRes, kSunLight, uv, risRnd, targetPdf, 1.0f / blended)) sunSample = cand;
    }
    finalizeResampling(sunRes, 1.0f, (float)nMis);
    sunRes.M = 1;

    Reservoir skyRes = emptyReservoir();
    LightSample skySample;
    for (int i = 0; i < nSky; ++i)
    {
        float sourcePdf;
        int idx = (int)aliasSample(sky.skyAlias, c.rnd16(), sourcePdf);
        LightSample cand = createSkyLightSample(sky, idx);
        int ix = idx % sky.skyW, iy = idx / sky.skyW;
        f2 uv = {(ix + 0.5f) / float(sky.skyW), (iy + 0.5f) / float(sky.skyH)};
        float blended = lightBrdfMisWeight(s, cand, sourcePdf, skyMisW, true, brdfMisW);
        float targetPdf = targetPdfForSurface(cand, s);
        float risRnd = c.rnd();
        if (streamSample(skyRes, kSkyLight, uv, risRnd, targetPdf, 1.0f / blended)) skySample = cand;
    }
    finalizeResampling(skyRes, 1.0f, (float)nMis);
    skyRes.M = 1;

    Reservoir brdfRes = emptyReservoir();
    LightSample brdfSample;
    for (int i = 0; i < nBrdf; ++i)
    {
        float lightSourcePdf = 0.0f;
        f3 sampleDir;
        uint32_t lightIndex = kInvalidLight;
        f2 uv = {0, 0};
        LightSample cand;
        float brdfPdf; bool trans = false; f3 dummy;
        disneySample(c.rnd4(), s.normal, s.geoNormal, s.wo, s.albedo, s.metallic, s.translucency, s.roughness, sampleDir, dummy, brdfPdf, trans);
        if (brdfPdf > 0.0f)
        {
            // BSDF-light ray (closesthit.cu:458-468, __closesthit__bsdf_light :854-901): miss -> sky / sun disk, an emissive
            // voxel face -> that face's triangle light, any other geometry -> no light
            Hit sh = c.trace(frontPos, sampleDir, 0.0f, FLT_MAX);
            if (sh.hit && nLocal > 0 && sh.face < 6 && sc.materials[sc.blockToMaterial[sh.id]].isEmissive)
            {
                // which of the face's two triangles, and the hit's barycentrics (optixGetTriangleBarycentrics)
                f3 A, eu, ev;
                faceFrame(sh.face, sh.x, sh.y, sh.z, A, eu, ev);
                const f3 P = hitPoint(sh, frontPos, sampleDir);
                const float fs = dot(P - A, eu), ft = dot(P - A, ev);
                const int tri = (fs + ft <= 1.0f) ? 0 : 1;
                const f2 bary = tri == 0 ? f2{fs, ft} : f2{1.0f - fs, 1.0f - ft};
                const int li = findLight(ll, (uint32_t)(sh.x + sc.grid.W() * (sh.z + sc.grid.D() * sh.y)), sh.face, tri);
                if (li >= 0 && li < numLights)
                {
                    lightIndex = (uint32_t)li;
                    uv = inverseTriangleSample(bary);
                    cand = triangleLightSample(createTriangleLight(ll.lights[li]), uv, s.pos);
                    lightSourcePdf = ll.alias[li].p;
                }
            }
            else if (!sh.hit)
            {
                if (equalAreaMapConeInv(uv, sky.sunDir, sampleDir, sunCosThetaMax()))
                {
                    lightIndex = kSunLight;
                    int sx = (int)(uv.x * sky.sunW - 0.5f), sy = (int)(uv.y * sky.sunH - 0.5f);
                    if (sx >= sky.sunW) sx %= sky.sunW;
                    if (sx < 0) sx = sky.sunW - ((-sx) % sky.sunW);
                    sy = clampi(sy, 0, sky.sunH - 1);
                    int idx = sy * sky.sunW + sx;
                    cand = createSunLightSample(sky, idx);
                    cand.position = sampleDir;
                    lightSourcePdf = sky.sunAlias[idx].p;
                }
                else
                {
                    lightIndex = kSkyLight;
                    uv = equalAreaSphereMapInv(sampleDir);
                    // clamp2i's result is discarded in the reference (closesthit.cu:508) — no clamp happens
                    int kx = (int)(uv.x * sky.skyW - 0.5f), ky = (int)(uv.y * sky.skyH - 0.5f);
                    int idx = ky * sky.skyW + kx;
                    // out-of-range only when uv.x == 1 exactly; keep memory safe with a clamp of the linear index
                    idx = clampi(idx, 0, sky.skyW * sky.skyH - 1);
                    cand = createSkyLightSample(sky, idx);
                    cand.position = sampleDir;
                    lightSourcePdf = sky.skyAlias[idx].p;
                }
            }
        }
        if (lightSourcePdf == 0.0f) continue;
        float targetPdf = targetPdfForSurface(cand, s);
        bool isEnv = lightIndex == kSkyLight || lightIndex == kSunLight;
        float misW = (lightIndex == kSkyLight) ? skyMisW : ((lightIndex == kSunLight) ? sunMisW : localMisW);
        float blended = lightBrdfMisWeight(s, cand, lightSourcePdf, misW, isEnv, brdfMisW);
        float risRnd = c.rnd();
        if (streamSample(brdfRes, lightIndex, uv, risRnd, targetPdf, 1.0f / blended)) brdfSample = cand;
    }
    finalizeResampling(brdfRes, 1.0f, (float)nMis);
    brdfRes.M = 1;

    combineReservoirs(ris, localRes, 0.5f, localRes.targetPdf);
    float r0 = c.rnd(); bool selSun = combineReservoirs(ris, sunRes, r0, sunRes.targetPdf);
    float r1 = c.rnd(); bool selSky = combineReservoirs(ris, skyRes, r1, skyRes.targetPdf);
    float r2 = c.rnd(); bool selBrdf = combineReservoirs(ris, brdfRes, r2, brdfRes.targetPdf);
    finalizeResampling(ris, 1.0f, 1.0f);
    ris.M = 1;
    if (selBrdf) lightSample = brdfSample;
    else if (selSky) lightSample = skySample;
    else if (selSun) lightSample = sunSample;
    else lightSample = localSample;

    bool isLightVisible = false;
    if (lightSample.lightType != LightInvalid && isValidReservoir(ris))
    {
        Hit vh = c.trace(frontPos, lightRayDir(lightSample, s.pos), 0.0f, lightRayTmax(lightSample, s.pos, 0.0f));
        isLightVisible = !vh.hit;
        if (!isLightVisible) { ris.lightData = 0; ris.weightSum = 0; }
    }

    Reservoir restir = emptyReservoir();
    if (enableReSTIR)
    {
        const Camera &pc = *c.prevCam;
        combineReservoirs(restir, ris, 0.5f, ris.targetPdf);
        f3 prevWorldPos = s.pos; // + motionWS (== 0)
        f2 prevUV = worldDirectionToUV(pc, normalize(prevWorldPos - pc.pos));
        i2 prevPixel = {(int)(prevUV.x * pc.resolution.x), (int)(prevUV.y * pc.resolution.y)};
        float expectedPrevDepth = distance(prevWorldPos, pc.pos);
        constexpr int nTemporal = 3;
        constexpr float mCap = 20.0f;
        i2 offs[nTemporal];
        offs[0] = {prevPixel.x - c.px, prevPixel.y - c.py};
        {
            f2 dsk = concentricSampleDisk(c.rnd2()) * 64.0f;
            offs[1] = {prevPixel.x - c.px + (int)dsk.x, prevPixel.y - c.py + (int)dsk.y};
        }
        {
            f2 dsk = concentricSampleDisk(c.rnd2()) * 64.0f;
            offs[2] = {(int)dsk.x, (int)dsk.y};
        }
        unsigned cached = 0;
        int selectedLoopIdx = -1;
        const size_t prevBase = (size_t)((c.iterationIndex + 1) & 1) * npix;
        for (int i = 0; i < nTemporal; ++i)
        {
            i2 idx = clampSamplePositionIntoView({c.px + offs[i].x, c.py + offs[i].y}, sc.width, sc.height);
            Surface ts;
            if (!getPrevSurface(c, ts, idx)) continue;
            bool nOk = dot(s.normal, ts.geoNormal) >= 0.5f;
            bool dOk = fabsf(expectedPrevDepth - ts.depth) <= 0.1f * fmaxr(expectedPrevDepth, ts.depth);
            bool rOk = fabsf(s.roughness - ts.roughness) <= 0.5f * fmaxr(s.roughness, ts.roughness);
            if (!(nOk && dOk && rOk)) continue;
            cached |= (1u << i);
            Reservoir pr = remapReservoir(sc, sc.reservoirs[prevBase + (size_t)idx.y * sc.width + idx.x]);
            if (std::isnan(pr.weightSum) || std::isinf(pr.weightSum)) pr = emptyReservoir();
            if (pr.M > mCap) pr.M = mCap;
            float neighborWeight = 0;
            LightSample cand;
            if (isValidReservoir(pr))
            {
                if (!lightSampleFromReservoir(sc, cand, pr, s, nLocal > 0)) pr = emptyReservoir();
                neighborWeight = targetPdfForSurface(cand, s);
            }
            if (combineReservoirs(restir, pr, c.rnd(), neighborWeight)) { lightSample = cand; selectedLoopIdx = i; }
        }
        if (isValidReservoir(restir))
        {
            float pi = restir.targetPdf, piSum = restir.targetPdf * 1;
            for (int i = 0; i < nTemporal; ++i)
            {
                if ((cached & (1u << i)) == 0) continue;
                i2 idx = clampSamplePositionIntoView({c.px + offs[i].x, c.py + offs[i].y}, sc.width, sc.height);
                Surface ts;
                getPrevSurface(c, ts, idx);
                LightSample atNeighbor;
                lightSampleFromReservoir(sc, atNeighbor, restir, ts, nLocal > 0);
                float ps = targetPdfForSurface(atNeighbor, ts);
                if (ps > 0 && !(i == 0 && i == selectedLoopIdx))
                {
                    const float extraRayOffset = 0.01f + 0.01f * ts.depth;
                    // the ray goes to `lightSample` (the sample at the CURRENT surface), as in the reference (closesthit.cu:743-744)
                    Hit nh = c.tracePrev(ts.pos, lightRayDir(lightSample, ts.pos), extraRayOffset, lightRayTmax(lightSample, ts.pos, extraRayOffset));
                    if (nh.hit) ps = 0.0f;
                }
                Reservoir pr = remapReservoir(sc, sc.reservoirs[prevBase + (size_t)idx.y * sc.width + idx.x]);
                if (std::isnan(pr.weightSum) || std::isinf(pr.weightSum)) pr = emptyReservoir();
                if (pr.M > mCap) pr.M = mCap;
                if (selectedLoopIdx == i) pi = ps;
                piSum += ps * pr.M;
            }
            finalizeResampling(restir, pi, piSum);
        }
        if (lightSample.lightType != LightInvalid)
        {
            Hit vh = c.trace(frontPos, lightRayDir(lightSample, s.pos), 0.0f, lightRayTmax(lightSample, s.pos, 0.0f));
            isLightVisible = !vh.hit;
            if (!isLightVisible) { restir.lightData = 0; restir.weightSum = 0; }
        }
    }

    const Reservoir shading = enableReSTIR ? restir : ris;
    if (lightSample.lightType != LightInvalid && isValidReservoir(shading) && isLightVisible)
    {
        f3 sampleDir = lightRayDir(lightSample, s.pos);
        const f3 albedo = skipAlbedoInShadowRay ? F3(1.0f) : s.albedo;
        f3 bsdf; float pdf;
        disneyEvaluate(s.normal, s.geoNormal, sampleDir, s.wo, albedo, s.metallic, s.translucency, s.roughness, bsdf, pdf);
        float cosTheta = fmaxf(0.0f, dot(sampleDir, s.normal));
        f3 shadowRad = bsdf * cosTheta * lightSample.radiance * shading.weightSum / lightSample.solidAnglePdf;
        rd.radiance += shadowRad;
    }
    if (storeSlot) *storeSlot = enableReSTIR ? restir : emptyReservoir();
}

// __raygen__pathtracer (RayGen.cu:102-181), one sample. Returns radiance; primary distance via out-param.
inline f3 tracePath(PixelCtx &c, bool ownsGBuffer, float &primaryDist)
{
    Scene &sc = const_cast<Scene &>(*c.sc);
    RayData rd{};
    c.randIdx = 0;
    f2 jitter = c.rnd2();
    f2 sampleUv = {(float(c.px) + jitter.x) * c.cam->inversedResolution.x, (float(c.py) + jitter.y) * c.cam->inversedResolution.y};
    rd.pos = c.cam->pos;
    rd.wi = uvToWorldDirection(*c.cam, sampleUv);
    f3 radiance = F3(0.0f), throughput = F3(1.0f);
    rd.rayConeWidth = 0.0f;
    rd.rayConeSpread = getRayConeWidth(*c.cam, c.px, c.py); // RayGen.cu:134-135
    rd.depth = 0;
    primaryDist = kRayMax;
    bool terminated = false;
    int totalBounce = 0, diffuseBounce = 0;
    while (!terminated)
    {
        // TraceNextPath (RayGen.cu:8-100)
        rd.bsdfOverPdf = F3(1.0f); rd.pdf = 0.0f; rd.radiance = F3(0.0f); rd.wo = -rd.wi; rd.distance = kRayMax;
        rd.shouldTerminate = false;
        rd.isLastBounceDiffuse = rd.isCurrentBounceDiffuse; rd.isCurrentBounceDiffuse = false;
        rd.hitFrontFace = false; rd.transmissionEvent = false;
        const f3 orig = rd.pos;
        Hit h = c.trace(orig, rd.wi, 0.0f, kRayMax);
        if (rd.depth == 0 && ownsGBuffer)
        {
            int32_t *ph = &sc.primaryHits[((size_t)c.py * sc.width + c.px) * 4];
            ph[0] = h.hit ? h.x : -1; ph[1] = h.hit ? h.y : -1; ph[2] = h.hit ? h.z : -1; ph[3] = h.hit ? h.face : -1;
        }
        if (h.hit) closestHit(c, rd, h, orig, ownsGBuffer);
        else missRadiance(c, rd, ownsGBuffer);
        radiance += throughput * rd.radiance;
        bool cont = !(rd.shouldTerminate || rd.pdf <= 0.0f || isNull(rd.bsdfOverPdf));
        if (cont) throughput *= rd.bsdfOverPdf;
        terminated = !cont;
        ++totalBounce;
        if (rd.isCurrentBounceDiffuse) ++diffuseBounce;
        if (totalBounce == sc.tp.totalBounceLimit || diffuseBounce == sc.tp.diffuseBounceLimit) terminated = true;
        if (rd.depth == 0) primaryDist = rd.distance;
        ++rd.depth;
    }
    if (std::isnan(radiance.x) || std::isnan(radiance.y) || std::isnan(radiance.z)) radiance = F3(0.5f);
    return radiance;
}

// One pixel of OptixRenderer::render: spp samples, sample 0 owns G-buffer / depth / reservoir.
// `sampleBegin/sampleStep` shard the spp loop (multi-GPU: rank r renders k = r, r+n, ...). ownerSample = the sample that owns
// the G-buffer / reservoir / ReSTIR pass: 0 (the reference), or sampleBegin for a rank-local owner.
inline void renderPixel(Scene &sc, const Camera &cam, const Camera &prevCam, int iterationIndex, int px, int py,
                        int sampleBegin, int sampleStep, f4 *accumOut, uint64_t &rays, uint64_t &steps, int ownerSample = 0, int sampleLimit = 0)
{
    PixelCtx c{&sc, &cam, &prevCam, px, py, iterationIndex, 0, 0, 0, 0, 0};
    const int spp = sc.tp.spp;
    f3 sum = F3(0.0f);
    float depth0 = kRayMax;
    bool haveDepth = false;
    int done = 0;
    for (int k = sampleBegin; k < spp && (sampleLimit <= 0 || done < sampleLimit); k += sampleStep, ++done)
    {
        c.sampleIndex = iterationIndex * spp + k;
        c.prevSampleIndex = (iterationIndex - 1) * spp + ownerSample; // the previous frame's owner sample (its G-buffer jitter)
        float pd;
        f3 r = tracePath(c, k == ownerSample, pd);
        sum += r;
        if (k == ownerSample) { depth0 = pd; haveDepth = true; }
    }
    size_t pix = (size_t)py * sc.width + px;
    if (haveDepth) sc.gb[sc.cur].depth[pix] = depth0;
    *accumOut = F4(sum, haveDepth ? depth0 : 0.0f);
    rays += c.rays; steps += c.steps;
}

} // namespace orc
