// Wavefront path tracer for sm_100a: replaces the reference's OptiX pipeline
// (__raygen__pathtracer / __closesthit__radiance / __miss__radiance,
//  /root/reference/renderer/shaders/RayGen.cu:102-181, closesthit.cu:10-852, miss.cu:9-82) — B200 has no RT cores.
//
// The per-pixel bounce loop of the reference is cut at every ray cast into STAGES over compact queues in HBM:
//
//   gen  -> [DDA closest] -> S1 -> [DDA any] -> S2 -> [DDA any] -> S3 -> [DDA any] -> S4 -> [DDA any] -> S5 -> accumulate
//   raygen  primary hit    material, G-buffer,    RIS merge,       temporal ReSTIR    finalize,       shade,
//                          BSDF sample, sun/sky   visibility ray   candidates,        final           reservoir store,
//                          candidates, BSDF-      of the winner    bias-correction    visibility      continuation ray
//                          candidate ray                           rays (<= 3)        ray
//
// Stage kernels are one thread per path with NO traversal state (no spills, high occupancy); every ray they spawn is
// set up once (vpt_dda.cuh: prepareRay) and appended to a queue with one atomic per warp. The DDA engine (vpt_dda.cu)
// consumes a queue with warp-level lane re-arming, so traversal runs near full lane occupancy no matter how
// divergent the trip counts are. S3/S4 run for sample 0 at depth 0 only (the ReSTIR sample, closesthit.cu:636-820);
// paths that continue past their first hit (specular chains, diffuse limit > 1) loop DDA/S1/DDA/S2/DDA/S5 per depth
// over an active-path list.
//
// Arithmetic: the exact class (vpt::ex::) up to and including every DDA set-up; shading is the fast class
// (vpt_math.cuh). RNG dimensions are consumed in exactly the reference's order (randIdx travels in the path flags).
#include "vpt_dda.cuh"
#include <cuda_fp16.h>

#ifndef VPT_FN
#define VPT_FN __device__ __forceinline__
#endif

namespace vpt {

constexpr float kSpawnEps = 0.0009765625f; // 2^-10
constexpr uint32_t kLightValidBit = 0x80000000u, kLightIndexMask = 0x7FFFFFFFu;
constexpr uint32_t kInvalidLight = 0x7FFFFFFFu, kSkyLight = 0x7FFFFFFEu, kSunLight = 0x7FFFFFFDu;
enum { LightInvalid = 0, LightSky = 1, LightSun = 2, LightLocalTriangle = 3 };
constexpr float kRoughnessThreshold = 0.00001f, kTranslucencyThreshold = 0.001f;
constexpr float kDisneyMinPdf = 1e-5f, kDisneyMaxThroughput = 32.0f, kDisneyMinLobeProb = 0.05f;

VPT_DEV f3 faceNormal(int face, f3 rayDir)
{
    switch (face)
    {
    case 0: return {0, 1, 0};
    case 1: return {0, -1, 0};
    case 2: return {-1, 0, 0};
    case 3: return {1, 0, 0};
    case 4: return {0, 0, 1};
    case 5: return {0, 0, -1};
    default:
    {
        float ax = fabsf(rayDir.x), ay = fabsf(rayDir.y), az = fabsf(rayDir.z);
        if (ax >= ay && ax >= az) return {rayDir.x > 0 ? -1.0f : 1.0f, 0, 0};
        if (ay >= az) return {0, rayDir.y > 0 ? -1.0f : 1.0f, 0};
        return {0, 0, rayDir.z > 0 ? -1.0f : 1.0f};
    }
    }
}
VPT_DEV f3 hitPoint(int hx, int hy, int hz, int face, float t, f3 o, f3 d)
{
    f3 p = ex::pointAt(o, d, t);
    switch (face)
    {
    case 0: p.y = (float)(hy + 1); break;
    case 1: p.y = (float)hy; break;
    case 2: p.x = (float)hx; break;
    case 3: p.x = (float)(hx + 1); break;
    case 4: p.z = (float)(hz + 1); break;
    case 5: p.z = (float)hz; break;
    default: break;
    }
    return p;
}

// ------------------------------------------------------------------------------------------------ BSDF
VPT_DEV f3 clampDisneyThroughput(f3 v)
{
    float a = fabsf(luminance(v));
    if (a > kDisneyMaxThroughput && a > 0.0f) return v * (kDisneyMaxThroughput / a);
    return v;
}
VPT_DEV float fresnelDielectric(float et, float cosIn)
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
VPT_DEV float disneyDiffuseFresnel(float cosWo, float cosWi, float roughness)
{
    float energyBias = lerpf(0.0f, 0.5f, roughness);
    float energyFactor = lerpf(1.0f, 1.0f / 1.51f, roughness);
    float fd90 = energyBias + 2.0f * roughness * cosWi * cosWi;
    float f0 = 1.0f;
    float lightScatter = f0 + (fd90 - f0) * pow5(1.0f - cosWo);
    float viewScatter = f0 + (fd90 - f0) * pow5(1.0f - cosWi);
    return lightScatter * viewScatter * energyFactor;
}
VPT_DEV float gtr2Aniso(float cosH, float sinH, float sinPhi, float cosPhi, float ax, float ay)
{
    float ax2 = ax * ax, ay2 = ay * ay;
    float s = (cosPhi * cosPhi) / ax2 + (sinPhi * sinPhi) / ay2;
    float t = sinH * sinH * s + cosH * cosH;
    return 1.0f / (kPi * ax * ay * t * t);
}
VPT_DEV float smithGGX(float cosTheta, float alpha)
{
    float a2 = alpha * alpha, c2 = cosTheta * cosTheta;
    return 2.0f / (1.0f + sqrtf(1.0f + a2 * (1.0f - c2) / c2));
}
VPT_DEV f3 disneyC0(f3 albedo, float metalness)
{
    float lum = 0.299f * albedo.x + 0.587f * albedo.y + 0.114f * albedo.z;
    f3 tint = lum > 0.0f ? albedo / lum : F3(1.0f);
    f3 specularColor = lerp3(F3(1.0f), tint, 0.0f);
    return lerp3(0.08f * 0.5f * specularColor, albedo, metalness);
}
VPT_DEV float disneySpecularProb(float avgF, float metalness, bool &valid, float &diffuseProb)
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

VPT_FN void disneySample(f4 u, f3 n, f3 ng, f3 wo, f3 albedo, bool metallic, float translucency, float roughness,
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
        else { wi = F3(0.0f); bsdfOverPdf = F3(0.0f); pdf = 0.0f; }
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
    if (!valid) { wi = F3(0.0f); bsdfOverPdf = F3(0.0f); pdf = 0.0f; return; }

    if (u.w < specularProb)
    {
        float cosTheta = sqrtf((1.0f - u.x) / (1.0f + (alpha * alpha - 1.0f) * u.x));
        cosTheta = clampf(cosTheta, kSafeCosEps, 1.0f);
        float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
        float phi = kTwoPi * u.y;
        float sphi, cphi;
        sincosFast(phi, sphi, cphi);
        f3 wh = {sinTheta * cphi, sinTheta * sphi, cosTheta};
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
        float sphi, cphi;
        sincosFast(phi, sphi, cphi);
        wi = {sinTheta * cphi, sinTheta * sphi, cosTheta};
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

VPT_FN void disneyEvaluate(f3 n, f3 ng, f3 wi, f3 wo, f3 albedo, bool metallic, float roughness, f3 &bsdf, float &pdf)
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


// ------------------------------------------------------------------------------------------------ shading state
struct Surface
{
    f3 pos; float depth;
    f3 normal, geoNormal, albedo, wo; float roughness; bool metallic; float translucency;
};
struct LightSample { f3 position, radiance; float solidAnglePdf; int lightType; };
VPT_DEV LightSample noLight() { LightSample l; l.position = F3(0.0f); l.radiance = F3(0.0f); l.solidAnglePdf = 0.0f; l.lightType = LightInvalid; return l; }
VPT_DEV VptReservoir emptyReservoir() { VptReservoir r; r.lightData = 0; r.uvData = 0; r.weightSum = 0.0f; r.targetPdf = 0.0f; r.M = 0.0f; return r; }
VPT_DEV bool isValidReservoir(const VptReservoir &r) { return r.lightData != 0; }
VPT_DEV bool sameDir(f3 a, f3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// ------------------------------------------------------------------------------------------------ local emissive lights
// TriangleLight (renderer/shaders/Light.h:44-137) over the host-built list of emissive voxel faces (host/vpt_lights.cpp).
struct TriLight { f3 base, edge1, edge2, radiance, normal; float surfaceArea; };
VPT_DEV float halfToFloat(uint32_t bits) { return __half2float(__ushort_as_half((unsigned short)(bits & 0xffffu))); }
VPT_DEV f3 octToNdirUnorm32(uint32_t u) // LinearMath.h:2069-2089
{
    const float px = saturate(float(u & 0xffffu) / 0xfffe) * 2.0f - 1.0f, py = saturate(float(u >> 16) / 0xfffe) * 2.0f - 1.0f;
    f3 n = {px, py, 1.0f - fabsf(px) - fabsf(py)};
    const float t = fmaxf(0.0f, -n.z);
    n.x += n.x >= 0.0f ? -t : t;
    n.y += n.y >= 0.0f ? -t : t;
    return normalize(n);
}
VPT_DEV TriLight createTriLight(const VptLightInfo *li) // TriangleLight::Create (Light.h:84-122): two 16-byte loads
{
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(li)), b = __ldg(reinterpret_cast<const uint4 *>(li) + 1);
    TriLight t;
    t.edge1 = octToNdirUnorm32(b.z) * halfToFloat(a.w);
    t.edge2 = octToNdirUnorm32(b.w) * halfToFloat(a.w >> 16);
    t.base = F3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)) - (t.edge1 + t.edge2) / 3.0f;
    t.radiance = {halfToFloat(b.x), halfToFloat(b.x >> 16), halfToFloat(b.y)};
    const f3 n = cross(t.edge1, t.edge2);
    const float len = length(n);
    if (len > 0.0f) { t.surfaceArea = 0.5f * len; t.normal = n / len; }
    else { t.surfaceArea = 0.0f; t.normal = F3(0.0f); }
    return t;
}
VPT_DEV LightSample triLightSample(const TriLight &t, f2 random, f3 viewer) // calcSample + calcSolidAnglePdf (Light.h:54-82), SampleTriangle, PdfAtoW
{
    LightSample r;
    const float sq = sqrtf(random.x);
    const float by = sq * (1.0f - random.y), bz = sq * random.y;
    r.position = t.base + t.edge1 * by + t.edge2 * bz;
    f3 L = r.position - viewer;
    const float Ldist = length(L);
    L = L / Ldist;
    const float cosTheta = saturate(dot(L, -t.normal));
    r.solidAnglePdf = (1.0f / t.surfaceArea) * (Ldist * Ldist) / cosTheta;
    r.radiance = t.radiance;
    r.lightType = LightLocalTriangle;
    return r;
}
// direction and far end of a visibility ray towards a light sample (closesthit.cu:616-617, 743-744, 801-802)
template <bool kLights> VPT_DEV f3 lightRayDir(const LightSample &ls, f3 from)
{
    return (kLights && ls.lightType == LightLocalTriangle) ? normalize(ls.position - from) : ls.position;
}
template <bool kLights> VPT_DEV float lightRayTmax(const LightSample &ls, f3 from, float extraRayOffset)
{
    return (kLights && ls.lightType == LightLocalTriangle) ? length(ls.position - from) - 0.01f - extraRayOffset : kRayMax;
}
// light index of (voxel, face, triangle): binary search of the ascending face keys, like the instance -> light table that
// __closesthit__bsdf_light searches (closesthit.cu:870-895); -1 = not a light
VPT_DEV int findLight(const LightView &lv, uint32_t lin, int face, int tri)
{
    const uint32_t key = (lin << 3) | (uint32_t)face;
    int left = 0, right = lv.numFaces - 1;
    while (left <= right)
    {
        const int mid = (left + right) >> 1;
        const uint32_t v = __ldg(lv.faceKeys + mid);
        if (v == key) return 2 * mid + tri;
        if (v < key) left = mid + 1; else right = mid - 1;
    }
    return -1;
}
VPT_DEV void faceFrameDev(int face, int x, int y, int z, f3 &A, f3 &u, f3 &v)
{
    const float fx = (float)x, fy = (float)y, fz = (float)z;
    switch (face)
    {
    case 0: A = {fx, fy + 1.0f, fz}; u = {0, 0, 1}; v = {1, 0, 0}; break;
    case 1: A = {fx, fy, fz}; u = {1, 0, 0}; v = {0, 0, 1}; break;
    case 2: A = {fx, fy, fz}; u = {0, 0, 1}; v = {0, 1, 0}; break;
    case 3: A = {fx + 1.0f, fy, fz}; u = {0, 1, 0}; v = {0, 0, 1}; break;
    case 4: A = {fx, fy, fz + 1.0f}; u = {1, 0, 0}; v = {0, 1, 0}; break;
    default: A = {fx, fy, fz}; u = {0, 1, 0}; v = {1, 0, 0}; break;
    }
}

// Per-thread shading context: pixel, sample index and the RNG dimension counter (RandGen.h:21-45).
struct Ctx
{
    const TraceArgs &a;
    int px, py, sampleIndex, prevSampleIndex, randIdx;
    // ranking keys of dims 0..15 (16 contiguous bytes from this pixel's slot: the reference indexes the ranking tile with
    // the un-wrapped dimension) and the 8 scrambling keys, fetched once per thread with three 8-byte loads
    unsigned long long rkLo, rkHi, sc;
    VPT_DEV void loadKeys()
    {
        const int base = ((px & 127) + (py & 127) * 128) * 8;
        const uint2 a0 = __ldg(reinterpret_cast<const uint2 *>(a.ranking + base));
        const uint2 a1 = __ldg(reinterpret_cast<const uint2 *>(a.ranking + base + 8));
        const uint2 s0 = __ldg(reinterpret_cast<const uint2 *>(a.scrambling + base));
        rkLo = ((unsigned long long)a0.y << 32) | a0.x;
        rkHi = ((unsigned long long)a1.y << 32) | a1.x;
        sc = ((unsigned long long)s0.y << 32) | s0.x;
    }
    VPT_DEV float blueNoise(int sIdx, int dim) const
    {
        // BlueNoiseRandGenerator::rand (RandGen.h:21-45); ranking table is padded with 256 zero bytes
        sIdx &= 255;
        int rank;
        if (dim < 16) rank = (int)((unsigned)((dim < 8 ? rkLo : rkHi) >> ((dim & 7) * 8)) & 0xffu);
        else rank = __ldg(a.ranking + dim + ((px & 127) + (py & 127) * 128) * 8);
        const int ranked = sIdx ^ rank;
        int value = __ldg(a.sobol + dim + ranked * 256);
        value ^= (int)((unsigned)(sc >> ((dim & 7) * 8)) & 0xffu);
        return value / 256.0f;
    }
    VPT_DEV float rnd() { return blueNoise(sampleIndex, randIdx++); }
    VPT_DEV f2 rnd2() { float x = rnd(); float y = rnd(); return {x, y}; }
    VPT_DEV f4 rnd4() { float x = rnd(); float y = rnd(); float z = rnd(); float w = rnd(); return {x, y, z, w}; }
    VPT_DEV float rnd16() { f2 u = rnd2(); return u.x + u.y / 256.0f; }
    VPT_DEV f3 uvToWorldDirection(const VptCamera &c, f2 uv) const { return normalize(mul(mat3From(c.uvToWorld), F3(uv.x, uv.y, 1.0f))); }
    VPT_DEV f2 worldDirectionToUV(const VptCamera &c, f3 d) const { f3 h = mul(mat3From(c.worldToUv), d); return {h.x / h.z, h.y / h.z}; }
    VPT_DEV f3 sunDir() const { return {a.sunDir[0], a.sunDir[1], a.sunDir[2]}; }
    VPT_DEV f4 loadSky(int x, int y) const { x = clampi(x, 0, a.skyW - 1); y = clampi(y, 0, a.skyH - 1); return F4(__ldg(a.sky + (size_t)y * a.skyW + x)); }
    VPT_DEV f4 loadSun(int x, int y) const { x = clampi(x, 0, a.sunW - 1); y = clampi(y, 0, a.sunH - 1); return F4(__ldg(a.sun + (size_t)y * a.sunW + x)); }

    VPT_DEV unsigned aliasSample(const VptAliasBin *bins, int len, float u, float &pmf) const
    {
        int offset = min(int(u * len), int(len - 1));
        float up = fminr(u * len - offset, 0.999999f);
        if (up < __ldg(&bins[offset].q)) { pmf = __ldg(&bins[offset].p); return (unsigned)offset; }
        int alias = __ldg(&bins[offset].alias);
        pmf = __ldg(&bins[alias].p);
        return (unsigned)alias;
    }
    VPT_FN LightSample createSunLightSample(int idx) const
    {
        uint32_t ixu, iyu;
        a.divSunW.div((uint32_t)idx, iyu, ixu);
        const int ix = (int)ixu, iy = (int)iyu;
        f2 uv = {(ix + 0.5f) / float(a.sunW), (iy + 0.5f) / float(a.sunH)};
        LightSample ls;
        ls.solidAnglePdf = (a.sunW * a.sunH) / (kTwoPi * (1.0f - a.sunCosThetaMax));
        ls.position = equalAreaMapCone(sunDir(), uv.x, uv.y, a.sunCosThetaMax);
        ls.radiance = xyz(loadSun(ix, iy));
        ls.lightType = LightSun;
        return ls;
    }
    VPT_FN LightSample createSkyLightSample(int idx) const
    {
        uint32_t ixu, iyu;
        a.divSkyW.div((uint32_t)idx, iyu, ixu);
        const int ix = (int)ixu, iy = (int)iyu;
        f2 uv = {(ix + 0.5f) / float(a.skyW), (iy + 0.5f) / float(a.skyH)};
        LightSample ls;
        ls.solidAnglePdf = (a.skyW * a.skyH) / (4.0f * kPi);
        ls.position = equalAreaSphereMap(uv.x, uv.y);
        ls.radiance = xyz(loadSky(ix, iy));
        ls.lightType = LightSky;
        return ls;
    }
    // GetLightSampleTargetPdfForSurface (Restir.h:194-211) and LightBrdfMisWeight (Restir.h:286-328, brdfCutoff == 0)
    // evaluate the same Disney BSDF for the same direction; evalCandidate evaluates it once and returns both:
    // the RIS target pdf and (optionally) the MIS-blended source pdf.
    VPT_FN float evalCandidate(const Surface &s, const LightSample &ls, float lightSelectionPdf, float lightMisWeight,
                                                float brdfMisWeight, float *blendedSourcePdf) const
    {
        const bool invalid = ls.solidAnglePdf <= 0 || ls.lightType == LightInvalid;
        const float lpdf = ls.solidAnglePdf;
        const bool plainMis = (brdfMisWeight == 0.0f || lpdf <= 0.0f || isinf(lpdf) || isnan(lpdf));
        if (invalid && (plainMis || !blendedSourcePdf))
        {
            if (blendedSourcePdf) *blendedSourcePdf = lightMisWeight * lightSelectionPdf;
            return 0.0f;
        }
        const f3 wi = (ls.lightType == LightLocalTriangle) ? normalize(ls.position - s.pos) : ls.position;
        f3 f; float pdf;
        disneyEvaluate(s.normal, s.geoNormal, wi, s.wo, s.albedo, s.metallic, s.roughness, f, pdf);
        if (blendedSourcePdf)
        {
            if (plainMis) *blendedSourcePdf = lightMisWeight * lightSelectionPdf;
            else
            {
                const float sourcePdfWrtSolidAngle = lightSelectionPdf * lpdf;
                const float blended = lightMisWeight * sourcePdfWrtSolidAngle + brdfMisWeight * pdf;
                *blendedSourcePdf = blended / lpdf;
            }
        }
        if (invalid) return 0.0f;
        const f3 refl = ls.radiance * f * fabsf(dot(wi, s.normal)) / ls.solidAnglePdf;
        return luminance(refl);
    }
    VPT_DEV float targetPdfForSurface(const LightSample &ls, const Surface &s) const { return evalCandidate(s, ls, 0.0f, 0.0f, 0.0f, nullptr); }
    template <bool kLights> VPT_FN bool lightSampleFromReservoir(LightSample &ls, const VptReservoir &r, f3 surfacePos) const
    {
        uint32_t li = r.lightData & kLightIndexMask;
        f2 uv = {float(r.uvData & 0xffff) / float(0xffff), float(r.uvData >> 16) / float(0xffff)};
        if (li == kSkyLight)
        {
            int x = clampi(int(uv.x * a.skyW), 0, a.skyW - 1), y = clampi(int(uv.y * a.skyH), 0, a.skyH - 1);
            ls = createSkyLightSample(y * a.skyW + x);
        }
        else if (li == kSunLight)
        {
            int x = clampi(int(uv.x * a.sunW), 0, a.sunW - 1), y = clampi(int(uv.y * a.sunH), 0, a.sunH - 1);
            ls = createSunLightSample(y * a.sunW + x);
        }
        else if (kLights && li < (uint32_t)a.lv.numLights) // hasLocalLights && lightIndex < numLights (Restir.h:404-410)
        {
            ls = triLightSample(createTriLight(a.lv.lights + li), uv, surfacePos);
            return true;
        }
        return li < kInvalidLight;
    }
    VPT_FN bool getPrevSurface(Surface &s, int x, int y) const
    {
        const VptCamera &pc = a.prevCam;
        if (x < 0 || y < 0 || x >= pc.resolution[0] || y >= pc.resolution[1]) return false;
        const size_t i = (size_t)y * a.width + x;
        // every plane of the pixel is requested before the depth test consumes the first: one round trip, not two
        const float depth = __ldg(a.prev.depth + i);
        const float4 nr = __ldg(a.prev.normalRoughness + i), gt = __ldg(a.prev.geoNormalThinfilm + i), mp = __ldg(a.prev.materialParameter + i);
        const float4 al = __ldg(a.prev.albedo + i);
        s.depth = depth;
        if (s.depth == kRayMax) return false;
        const float j0 = blueNoise(prevSampleIndex, 0), j1 = blueNoise(prevSampleIndex, 1);
        f2 prevUV = {(float(x) + j0) * pc.inversedResolution[0], (float(y) + j1) * pc.inversedResolution[1]};
        f3 viewDir = uvToWorldDirection(pc, prevUV);
        s.pos = F3(pc.pos[0], pc.pos[1], pc.pos[2]) + viewDir * s.depth;
        s.wo = -viewDir;
        s.normal = xyz(nr);
        s.geoNormal = xyz(gt);
        s.albedo = xyz(al);
        s.roughness = nr.w;
        s.metallic = (mp.x == 1.0f);
        s.translucency = mp.y;
        return true;
    }
};

VPT_DEV bool streamSample(VptReservoir &r, uint32_t lightIndex, f2 uv, float random, float targetPdf, float invSourcePdf)
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
VPT_DEV bool combineReservoirs(VptReservoir &r, const VptReservoir &nr, float random, float targetPdf)
{
    float risWeight = targetPdf * (nr.weightSum * nr.M);
    r.M += nr.M;
    r.weightSum += risWeight;
    bool sel = (random * r.weightSum < risWeight);
    if (sel) { r.lightData = nr.lightData; r.uvData = nr.uvData; r.targetPdf = targetPdf; }
    return sel;
}
VPT_DEV void finalizeResampling(VptReservoir &r, float num, float den)
{
    float d = r.targetPdf * den;
    r.weightSum = (d == 0.0f) ? 0.0f : (r.weightSum * num) / d;
}
VPT_DEV void clampIntoView(int &x, int &y, int width, int height)
{
    if (x < 0) x = -x;
    if (y < 0) y = -y;
    if (x >= width) x = 2 * width - x - 1;
    if (y >= height) y = 2 * height - y - 1;
}
VPT_DEV void storeReservoir(VptReservoir *dst, const VptReservoir &r)
{
    // 20-byte AoS record (RestirCommon.h): five scalar stores, 4-byte aligned
    dst->lightData = r.lightData; dst->uvData = r.uvData; dst->weightSum = r.weightSum; dst->targetPdf = r.targetPdf; dst->M = r.M;
}
VPT_DEV VptReservoir loadReservoir(const VptReservoir *src)
{
    VptReservoir r;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(src);
    r.lightData = __ldg(p); r.uvData = __ldg(p + 1);
    r.weightSum = __uint_as_float(__ldg(p + 2)); r.targetPdf = __uint_as_float(__ldg(p + 3)); r.M = __uint_as_float(__ldg(p + 4));
    return r;
}


// ------------------------------------------------------------------------------------------------ wavefront plumbing
enum : uint32_t
{
    F_LIVE = 1u << 0,     // slot holds a path
    F_RIS = 1u << 1,      // diffuse hit: RIS pending (S2, S5)
    F_RAY1 = 1u << 2,     // BSDF-candidate ray cast (result in vis1)
    F_RAY2 = 1u << 3,     // RIS visibility ray queued (result in vis2)
    F_VIS = 1u << 4,      // light visibility, once known
    F_RESTIR = 1u << 5,   // temporal ReSTIR path (sample 0, depth 0)
    F_RAY5 = 1u << 6,     // final visibility ray queued (result in vis4)
    F_CONT = 1u << 7,     // path continues to the next depth
    F_LIGHTOK = 1u << 9,  // selected light sample is valid
    F_VISHAVE1 = 1u << 10 // the RIS winner's visibility was resolved (direction = lightA)
};
constexpr int kRandShift = 16, kDepthShift = 24, kDiffuseShift = 28;
// Stage CTAs: 128 threads. Measured on B200 with the per-warp queue reservation (shading ms/frame, same register budgets):
// 256 -> 1.373, 128 -> 1.328, 64 -> 1.333, 32 -> 1.438 (with the CTA-wide reservation of round 1: 512 -> 1.710, 256 -> 1.504,
// 128 -> 1.443, 64 -> 1.439). Resident CTAs per SM (register budget) per stage, from the same kind of variant runs:
// S1/S2 8 (64 regs; 7 -> 1.476, 10 -> 1.566), S3 7 (73 regs; 5 -> 1.485, 6 -> 1.443, 7 -> 1.434), S5 16 (12 -> 1.443, 16 -> 1.439).
#ifndef VPT_SHADE_THREADS
#define VPT_SHADE_THREADS 128
#endif
constexpr int kShadeThreads = VPT_SHADE_THREADS;
#ifndef VPT_S3_MINB
#define VPT_S3_MINB 7
#endif
#ifndef VPT_S5_MINB
#define VPT_S5_MINB 16
#endif
#ifndef VPT_SHADE_MINB
#define VPT_SHADE_MINB 8
#endif
#ifndef VPT_S2_MINB
#define VPT_S2_MINB VPT_SHADE_MINB
#endif
constexpr int kCntWords = 256, kCntList = 128; // cnt[2k], cnt[2k+1] = count / cursor of the k-th DDA launch; cnt[128+d] = active paths at depth d

#ifndef VPT_RESERVE_WARP
#define VPT_RESERVE_WARP 1
#endif
// Queue reservation. Every thread of the warp calls it (convergent); n = entries wanted. One atomic per WARP and no barrier: the
// queue stays dense (what the DDA engine wants) and no warp waits for the CTA. VPT_RESERVE_WARP=0: one atomic per CTA behind a
// CTA-wide scan with two barriers — 12 % of shade1Kernel's stall samples sat in it (ncu r2p, tools/ncu_lines.py); it was the default
// while the stages ran 512-thread CTAs. Measured with 128-thread CTAs (stages, ms per frame): per CTA 1.360, per warp 1.333 — the
// 260 k same-address atomics per launch are not the bottleneck the per-CTA form was built to avoid.
template <int kSite = 0> VPT_DEV unsigned ctaReserve(unsigned n, unsigned *counter)
{
#if VPT_RESERVE_WARP
    const unsigned lane = threadIdx.x & 31;
    unsigned incl = n;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (unsigned)off) incl += v;
    }
    unsigned base = 0;
    if (lane == 31 && incl) base = atomicAdd(counter, incl);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - n;
#else
    __shared__ unsigned warpSum[kShadeThreads / 32];
    __shared__ unsigned ctaBase;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = n;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (unsigned)off) incl += v;
    }
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned tot = 0;
#pragma unroll
        for (int w = 0; w < kShadeThreads / 32; ++w) { const unsigned s = warpSum[w]; warpSum[w] = tot; tot += s; }
        ctaBase = tot ? atomicAdd(counter, tot) : 0u;
    }
    __syncthreads();
    // no trailing barrier: a kernel's second reservation (S5) is another instantiation with its own shared words
    return ctaBase + warpSum[warp] + incl - n;
#endif
}

struct PathId { int p, slot, sl, px, py, k; bool inImage; };
VPT_DEV PathId pathId(const TraceArgs &a, int p)
{
    PathId id;
    id.p = p;
    uint32_t sl, slot, ty, tx;
    a.divSlots.div((uint32_t)p, sl, slot);
    id.sl = (int)sl; id.slot = (int)slot;
    const int lane = id.slot & 31;
    a.divTilesX.div(slot >> 5, ty, tx);
    id.px = (int)tx * 8 + (lane & 7);
    id.py = (int)ty * 4 + (lane >> 3);
    id.k = a.sampleBegin + (a.waveFirst + id.sl) * a.sampleStep;
    id.inImage = id.px < a.width && id.py < a.height;
    return id;
}
// index within this part's launch -> path
VPT_DEV int partPath(const TraceArgs &a, int idx)
{
    uint32_t sl, ls;
    a.divPartSlots.div((uint32_t)idx, sl, ls);
    return (int)sl * a.nSlots + a.slotBase + (int)ls;
}
VPT_DEV Ctx makeCtx(const TraceArgs &a, const PathId &id, int randIdx)
{
    Ctx c{a, id.px, id.py, a.iterationIndex * a.spp + id.k, (a.iterationIndex - 1) * a.spp + a.ownerSample, randIdx, 0ull, 0ull, 0ull};
    c.loadKeys();
    return c;
}
VPT_DEV f3 camPos(const TraceArgs &a) { return F3(a.cam.pos[0], a.cam.pos[1], a.cam.pos[2]); }

// ---- textured materials (closesthit.cu:166-254): software trilinear over RGBA8 mip chains, wrap addressing
VPT_DEV f4 texelRGBA(const TraceArgs &a, int4 d, int level, int x, int y)
{
    const int n = d.y >> level;
    x &= n - 1; y &= n - 1;                                     // power-of-two wrap (two's complement handles negatives)
    const int w2 = d.y * d.y, n2 = n * n;
    const uint32_t v = __ldg(a.texels + (size_t)d.x + (size_t)((4 * w2 - 4 * n2) / 3) + (size_t)y * n + x);
    return {(float)(v & 0xffu) / 255.0f, (float)((v >> 8) & 0xffu) / 255.0f, (float)((v >> 16) & 0xffu) / 255.0f, (float)(v >> 24) / 255.0f};
}
VPT_DEV f4 texBilinear(const TraceArgs &a, int4 d, int level, float u, float v)
{
    const int n = d.y >> level;
    const float x = u * n - 0.5f, y = v * n - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float wa = x - fx, wb = y - fy;
    const int i = (int)fx, j = (int)fy;
    const f4 t00 = texelRGBA(a, d, level, i, j), t10 = texelRGBA(a, d, level, i + 1, j), t01 = texelRGBA(a, d, level, i, j + 1), t11 = texelRGBA(a, d, level, i + 1, j + 1);
    return ((1.0f - wa) * (1.0f - wb)) * t00 + (wa * (1.0f - wb)) * t10 + ((1.0f - wa) * wb) * t01 + (wa * wb) * t11;
}
VPT_DEV f4 tex2DLod(const TraceArgs &a, int tex, float u, float v, float lod)
{
    const int4 d = __ldg(a.texDescs + tex);
    const float maxLod = (float)(d.z - 1);
    lod = lod < 0.0f ? 0.0f : (lod > maxLod ? maxLod : lod);
    if (!(lod >= 0.0f)) lod = 0.0f;
    const int l0 = (int)floorf(lod);
    const int l1 = l0 + 1 < d.z ? l0 + 1 : l0;
    const float beta = lod - (float)l0;
    const f4 c0 = texBilinear(a, d, l0, u, v), c1 = texBilinear(a, d, l1, u, v);
    return c0 + beta * (c1 - c0);
}
// Camera::getRayConeWidth (Camera.h:133-149)
VPT_DEV float rayConeSpread(const VptCamera &c, int ix, int iy)
{
    const float pcx = ((float)ix + 0.5f) - c.resolution[0] / 2, pcy = ((float)iy + 0.5f) - c.resolution[1] / 2;
    const float ox = copysignf(0.5f, pcx), oy = copysignf(0.5f, pcy);
    const float nx = (pcx - ox) * c.inversedResolution[0] * 2 * c.tanHalfFov[0], ny = (pcy - oy) * c.inversedResolution[1] * 2 * c.tanHalfFov[1];
    const float fx = (pcx + ox) * c.inversedResolution[0] * 2 * c.tanHalfFov[0], fy = (pcy + oy) * c.inversedResolution[1] * 2 * c.tanHalfFov[1];
    return atanf(sqrtf(fx * fx + fy * fy)) - atanf(sqrtf(nx * nx + ny * ny));
}

// Surface of the current hit, rebuilt from the 32-byte record S1 wrote (closesthit.cu:160-256 without textures).
template <bool kTex> VPT_DEV Surface loadSurface(const TraceArgs &a, int p, int depth)
{
    const float4 sa = __ldg(a.wb.surfA + p), sb = __ldg(a.wb.surfB + p);
    const uint32_t bits = __float_as_uint(sb.w);
    const int face = (int)(bits & 7u);
    const VptMaterial *mat = a.materials + (bits >> 3);
    Surface s;
    s.pos = xyz(sa); s.depth = sa.w; s.wo = xyz(sb);
    s.geoNormal = faceNormal(face, -s.wo);
    s.translucency = __ldg(&mat->translucency);
    if (kTex)
    {
        // textured scene: S1 stored what it fetched
        const float4 sc = __ldg(a.wb.surfC + p), sd = __ldg(a.wb.surfD + p);
        s.normal = xyz(sc); s.roughness = sc.w; s.albedo = xyz(sd); s.metallic = sd.w != 0.0f;
        return s;
    }
    s.normal = lerp3(s.geoNormal, s.geoNormal, 0.2f);
    s.albedo = max3f(F3(__ldg(&mat->albedo[0]), __ldg(&mat->albedo[1]), __ldg(&mat->albedo[2])), F3(0.001f));
    s.roughness = __ldg(&mat->roughness);
    if (depth > 0) s.roughness = fminr(s.roughness * 2.0f + 0.1f, 1.0f);
    s.metallic = __ldg(&mat->metallic) != 0;
    return s;
}
VPT_DEV void storeLight(const TraceArgs &a, float4 *A, float4 *B, int i, const LightSample &l)
{
    A[i] = make_float4(l.position.x, l.position.y, l.position.z, l.solidAnglePdf);
    B[i] = make_float4(l.radiance.x, l.radiance.y, l.radiance.z, __int_as_float(l.lightType));
}
VPT_DEV LightSample loadLight(const float4 *A, const float4 *B, int i)
{
    const float4 la = __ldg(A + i), lb = __ldg(B + i);
    LightSample l;
    l.position = xyz(la); l.solidAnglePdf = la.w; l.radiance = xyz(lb); l.lightType = __float_as_int(lb.w);
    return l;
}
VPT_DEV void writeSkyGBuffer(const TraceArgs &a, size_t pix)
{
    a.cur.albedo[pix] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    a.cur.material[pix] = (float)0xFFFF;
    a.cur.normalRoughness[pix] = make_float4(0.0f, -1.0f, 0.0f, 0.0f);
    a.cur.geoNormalThinfilm[pix] = make_float4(0.0f, -1.0f, 0.0f, 0.0f);
    a.cur.materialParameter[pix] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// __miss__radiance (miss.cu:9-82): sky bicubic-smoothstep lookup + sun disk
VPT_DEV f3 missEmission(const Ctx &c, f3 rayDir)
{
    const TraceArgs &a = c.a;
    f3 emission = F3(0.0f);
    f2 uv = equalAreaSphereMapInv(rayDir);
    {
        f2 UV = {uv.x * a.skyW, uv.y * a.skyH};
        f2 tc = {floorf(UV.x - 0.5f) + 0.5f, floorf(UV.y - 0.5f) + 0.5f};
        f2 f = UV - tc;
        f2 f2_ = f * f, f3_ = f2_ * f;
        f2 w1 = {-2.0f * f3_.x + 3.0f * f2_.x, -2.0f * f3_.y + 3.0f * f2_.y};
        f2 w0 = {1.0f - w1.x, 1.0f - w1.y};
        int tx0 = (int)floorf(UV.x - 0.5f), ty0 = (int)floorf(UV.y - 0.5f);
        const int xs[4] = {tx0, tx0 + 1, tx0, tx0 + 1}, ys[4] = {ty0, ty0, ty0 + 1, ty0 + 1};
        const float ws[4] = {w0.x * w0.y, w1.x * w0.y, w0.x * w1.y, w1.x * w1.y};
        f3 out = F3(0.0f);
        float sumW = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            int x = xs[i], y = ys[i];
            if (x >= a.skyW) x %= a.skyW;
            if (x < 0) x = a.skyW - (-x) % a.skyW;
            if (y >= a.skyH) y = a.skyH - 1;
            if (y < 0) y = 0;
            sumW += ws[i];
            out += xyz(c.loadSky(x, y)) * ws[i];
        }
        out /= sumW;
        emission += out;
    }
    if (equalAreaMapConeInv(uv, c.sunDir(), rayDir, a.sunCosThetaMax))
    {
        int sx = (int)(uv.x * a.sunW), sy = (int)(uv.y * a.sunH);
        if (sx >= a.sunW) sx %= a.sunW;
        if (sx < 0) sx = a.sunW - (-sx) % a.sunW;
        emission += xyz(c.loadSun(sx, sy));
    }
    return emission;
}

// ------------------------------------------------------------------------------------------------ gen: raygen
// RayGen.cu:102-135: jittered primary ray, exact arithmetic up to the prepared DDA state.
__global__ void __launch_bounds__(kShadeThreads) genKernel(const __grid_constant__ TraceArgs a, unsigned *qCount)
{
    // Every path spawns exactly one primary ray, so its queue slot is its own index: no reservation, no barrier (the CTA-wide
    // reservation was 37 % of this kernel's stall samples, ncu r1k). A path without a ray to trace (a tile pixel outside the
    // image, a ray that misses the grid) queues a ray that starts on the shell corner (padded voxel 0): the engine retires
    // it on its first test as a miss, which is the result such a path needs.
    const int idx = blockIdx.x * kShadeThreads + threadIdx.x;
    if (idx == 0) atomicAdd(qCount, (unsigned)a.partPaths);
    if (idx >= a.partPaths) return;
    const int p = partPath(a, idx);
    const PathId id = pathId(a, p);
    uint32_t fl = 0;
    bool want = false;
    PreparedRay r;
    if (id.inImage)
    {
        Ctx c = makeCtx(a, id, 0);
        const f2 jitter = c.rnd2();
        const f2 sampleUv = {ex::mulf(ex::addf(float(id.px), jitter.x), a.cam.inversedResolution[0]),
                             ex::mulf(ex::addf(float(id.py), jitter.y), a.cam.inversedResolution[1])};
        const f3 dir = ex::normalize(ex::mulMat3(a.cam.uvToWorld, F3(sampleUv.x, sampleUv.y, 1.0f)));
        a.wb.dirT[p] = make_float4(dir.x, dir.y, dir.z, 0.0f);
        fl = F_LIVE | ((uint32_t)c.randIdx << kRandShift);
        want = prepareRay(a.grid, camPos(a), dir, 0.0f, (uint32_t)p, r);
    }
    a.wb.pflag[p] = fl;
    if (!want)
    {
        r.tMaxX = r.tMaxY = r.tMaxZ = r.tCur = 0.0f; r.tDeltaX = r.tDeltaY = r.tDeltaZ = r.tmin = 0.0f;
        r.lin = 0; r.meta = 6u << 4; r.result = (uint32_t)p; r.tmax = kRayMax;
    }
    storePreparedRay(a.wb.queue, (unsigned)idx, r);
}

// ------------------------------------------------------------------------------------------------ S1
// __miss__radiance, and __closesthit__radiance up to the BSDF-candidate ray (closesthit.cu:96-468).
template <bool kTex, bool kLights> __global__ void __launch_bounds__(kShadeThreads, VPT_SHADE_MINB) shade1Kernel(const __grid_constant__ TraceArgs a, int depth, const int *__restrict__ list,
                                                              const unsigned *__restrict__ listCount, unsigned *qCount)
{
    const int idx = blockIdx.x * kShadeThreads + threadIdx.x;
    const int n = list ? (int)__ldg(listCount) : a.partPaths;
    bool act = idx < n;
    const int p = act ? (list ? __ldg(list + idx) : partPath(a, idx)) : 0;
    const uint32_t fl = act ? a.wb.pflag[p] : 0u;
    act = act && (fl & F_LIVE);
    bool want = false;
    PreparedRay r;
    if (act)
    {
        const PathId id = pathId(a, p);
        const bool owns = (id.k == a.ownerSample);
        Ctx c = makeCtx(a, id, (int)((fl >> kRandShift) & 0xffu));
        const int diffuseBounce = (int)(fl >> kDiffuseShift);
        const size_t pix = (size_t)id.py * a.width + id.px;
        const f3 d = xyz(__ldg(a.wb.dirT + p));
        const f3 o = depth == 0 ? camPos(a) : xyz(__ldg(a.wb.org + p));
        const uint32_t hp = a.wb.hitPacked[p];
        const float t = a.wb.hitT[p];
        const bool gbufferPass = owns && depth == 0;
        float coneWidth = depth == 0 ? 0.0f : __ldg(a.wb.dirT + p).w; // rayData->rayConeWidth (RayGen.cu:134, closesthit.cu:195)
        uint32_t nf = F_LIVE | ((uint32_t)depth << kDepthShift);
        int newDiffuse = diffuseBounce;
        f3 radiance = F3(0.0f);
        float distance = kRayMax;

        if (hp == kHitMiss)
        {
            if (gbufferPass)
            {
                a.primaryHits[pix] = make_int4(-1, -1, -1, -1);
                storeReservoir(a.resCur + pix, emptyReservoir());
                writeSkyGBuffer(a, pix);
            }
            radiance = missEmission(c, d);
        }
        else
        {
            const int lin = (int)(hp >> 3), face = (int)(hp & 7u);
            uint32_t hxu, yzu, hzu, hyu;
            a.grid.divW.div((uint32_t)lin, yzu, hxu);
            a.grid.divD.div(yzu, hyu, hzu);
            const int hx = (int)hxu, hy = (int)hyu, hz = (int)hzu;
            if (gbufferPass) a.primaryHits[pix] = make_int4(hx, hy, hz, face);
            const int blockId = __ldg(a.grid.idsLinear + lin);
            distance = t;
            const f3 geoNormal = faceNormal(face, d);
            const f3 surfPos = hitPoint(hx, hy, hz, face, t, o, d);
            const f3 frontPos = surfPos + geoNormal * kSpawnEps;
            const uint32_t matIndex = __ldg(a.blockToMaterial + blockId);
            const VptMaterial *mat = a.materials + matIndex;
            const f3 matAlbedo = {__ldg(&mat->albedo[0]), __ldg(&mat->albedo[1]), __ldg(&mat->albedo[2])};
            if (__ldg(&mat->isEmissive))
            {
                if (depth == 0) // !hitFirstDiffuseSurface
                {
                    radiance = matAlbedo;
                    if (owns) writeSkyGBuffer(a, pix);
                }
            }
            else
            {
                Surface s;
                s.geoNormal = geoNormal;
                s.wo = -d;
                s.albedo = matAlbedo;
                s.roughness = __ldg(&mat->roughness);
                s.metallic = __ldg(&mat->metallic) != 0;
                s.translucency = __ldg(&mat->translucency);
                s.normal = geoNormal;
                if (kTex)
                {
                    // world-grid UV by the dominant normal axis + ray-cone LOD (closesthit.cu:166-200), then the four maps (:202-250)
                    const float uvScale = __ldg(&mat->uvScale);
                    f2 tc = {0.0f, 0.0f};
                    if (__ldg(&mat->useWorldGridUV))
                    {
                        if (fabsf(geoNormal.x) > 0.9f) tc = {fmodf(frontPos.z, uvScale), fmodf(frontPos.y, uvScale)};
                        else if (fabsf(geoNormal.y) > 0.9f) tc = {fmodf(frontPos.x, uvScale), fmodf(frontPos.z, uvScale)};
                        else if (fabsf(geoNormal.z) > 0.9f) tc = {fmodf(frontPos.x, uvScale), fmodf(frontPos.y, uvScale)};
                    }
                    tc = {tc.x / uvScale, tc.y / uvScale};
                    coneWidth += rayConeSpread(a.cam, id.px, id.py) * distance;
                    const int4 slots = __ldg(a.matTexSlots + matIndex);
                    const float lod = log2f(coneWidth / fmaxr(dot(geoNormal, s.wo), 0.2f) / uvScale * 2.0f * __ldg(a.matTexMip0Size + matIndex)) - 3.0f;
                    if (slots.x >= 0) s.albedo = s.albedo * xyz(tex2DLod(a, slots.x, tc.x, tc.y, lod));
                    if (slots.z >= 0) s.roughness = tex2DLod(a, slots.z, tc.x, tc.y, lod).x;
                    if (slots.w >= 0) s.metallic = tex2DLod(a, slots.w, tc.x, tc.y, lod).x > 0.5f;
                    if (slots.y >= 0)
                    {
                        f3 nm = normalize(xyz(tex2DLod(a, slots.y, tc.x, tc.y, lod)) - F3(0.5f));
                        nm.x = -nm.x; nm.y = -nm.y;
                        alignVector(geoNormal, nm);
                        s.normal = nm;
                    }
                }
                s.albedo = max3f(s.albedo, F3(0.001f));
                if (depth > 0) s.roughness = fminr(s.roughness * 2.0f + 0.1f, 1.0f);
                const bool isDiffuse = s.roughness > kRoughnessThreshold;
                s.normal = lerp3(geoNormal, s.normal, 0.2f);
                if (gbufferPass)
                {
                    a.cur.material[pix] = (float)__ldg(&mat->materialId);
                    a.cur.normalRoughness[pix] = make_float4(s.normal.x, s.normal.y, s.normal.z, s.roughness);
                    a.cur.geoNormalThinfilm[pix] = make_float4(s.normal.x, s.normal.y, s.normal.z, 0.0f);
                    a.cur.materialParameter[pix] = make_float4(s.metallic ? 1.0f : 0.0f, s.translucency, 0.0f, 0.0f);
                    a.cur.albedo[pix] = make_float4(s.albedo.x, s.albedo.y, s.albedo.z, 1.0f);
                }
                // The bounce sample (closesthit.cu:282-293) is only looked at when the path may continue; when the bounce limits
                // end it here anyway, its four RNG dimensions are consumed without evaluating the BSDF.
                const bool mayContinue = !(depth + 1 == a.totalBounceLimit || diffuseBounce + (isDiffuse ? 1 : 0) == a.diffuseBounceLimit);
                f3 bsdfWi = F3(0.0f), bsdfOverPdf = F3(0.0f); float bsdfPdf = 0.0f; bool transmission = false;
                if (mayContinue)
                    disneySample(c.rnd4(), s.normal, s.geoNormal, s.wo, s.albedo, s.metallic, s.translucency, s.roughness, bsdfWi, bsdfOverPdf, bsdfPdf, transmission);
                else
                    c.randIdx += 4;
                s.pos = frontPos;
                s.depth = distance;
                bool needSurf = false;
                if (!isDiffuse)
                {
                    if (gbufferPass) storeReservoir(a.resCur + pix, emptyReservoir());
                }
                else
                {
                    newDiffuse = diffuseBounce + 1;
                    needSurf = true;
                    const f3 sunD = c.sunDir();
                    const bool skipSun = (dot(s.normal, sunD) < 0.0f || dot(s.geoNormal, sunD) < 0.0f);
                    const int nSun = skipSun ? 0 : 1;
                    constexpr int nLocal = kLights ? 8 : 0; // closesthit.cu:330: 8 when the scene has lights (kLights: a separate instance, so scenes without lights carry none of this)
                    const int nMis = nLocal + nSun + 2;
                    const float localMisW = float(nLocal) / nMis, sunMisW = float(nSun) / nMis, skyMisW = 1.0f / nMis, brdfMisW = 1.0f / nMis;

                    // local-light candidates (closesthit.cu:347-375): alias-sampled light, point on its triangle, RIS stream
                    f2 candUv = {0.0f, 0.0f};
                    if (nLocal > 0)
                    {
                        VptReservoir localRes = emptyReservoir();
                        int localIdx = -1;
                        f2 localUv = {0.0f, 0.0f};
#pragma unroll 1
                        for (int i = 0; i < nLocal; ++i)
                        {
                            float sourcePdf;
                            const int li = (int)c.aliasSample(a.lv.alias, a.lv.numLights, c.rnd(), sourcePdf);
                            if (li >= a.lv.numLights) continue;
                            const f2 uv = c.rnd2();
                            const LightSample cand = triLightSample(createTriLight(a.lv.lights + li), uv, s.pos);
                            float blended;
                            const float targetPdf = c.evalCandidate(s, cand, sourcePdf, localMisW, brdfMisW, &blended);
                            const float risRnd = c.rnd();
                            if (blended != 0.0f && streamSample(localRes, (uint32_t)li, uv, risRnd, targetPdf, 1.0f / blended)) { localIdx = li; localUv = uv; }
                        }
                        finalizeResampling(localRes, 1.0f, (float)nMis);
                        a.wb.candC[p] = make_uint4((uint32_t)localIdx, __float_as_uint(localRes.weightSum), __float_as_uint(localRes.targetPdf), 0u); // integer plane: the index never passes through a float register (-ftz)
                        candUv = localUv;
                    }

                    // sun candidate (closesthit.cu:380-420; Restir.h:221-250)
                    VptReservoir sunRes = emptyReservoir();
                    int sunIdx = -1;
                    for (int i = 0; i < nSun; ++i)
                    {
                        float sourcePdf;
                        const int cidx = (int)c.aliasSample(a.sunAlias, a.sunW * a.sunH, c.rnd(), sourcePdf);
                        const LightSample cand = c.createSunLightSample(cidx);
                        uint32_t ixu, iyu;
                        a.divSunW.div((uint32_t)cidx, iyu, ixu);
                        const int ix = (int)ixu, iy = (int)iyu;
                        const f2 uv = {(ix + 0.5f) / float(a.sunW), (iy + 0.5f) / float(a.sunH)};
                        float blended;
                        const float targetPdf = c.evalCandidate(s, cand, sourcePdf, sunMisW, brdfMisW, &blended);
                        const float risRnd = c.rnd();
                        if (streamSample(sunRes, kSunLight, uv, risRnd, targetPdf, 1.0f / blended)) sunIdx = cidx;
                    }
                    finalizeResampling(sunRes, 1.0f, (float)nMis);
                    // sky candidate (closesthit.cu:422-450; Restir.h:252-283)
                    VptReservoir skyRes = emptyReservoir();
                    int skyIdx = -1;
                    {
                        float sourcePdf;
                        const int cidx = (int)c.aliasSample(a.skyAlias, a.skyW * a.skyH, c.rnd16(), sourcePdf);
                        const LightSample cand = c.createSkyLightSample(cidx);
                        uint32_t ixu, iyu;
                        a.divSkyW.div((uint32_t)cidx, iyu, ixu);
                        const int ix = (int)ixu, iy = (int)iyu;
                        const f2 uv = {(ix + 0.5f) / float(a.skyW), (iy + 0.5f) / float(a.skyH)};
                        float blended;
                        const float targetPdf = c.evalCandidate(s, cand, sourcePdf, skyMisW, brdfMisW, &blended);
                        const float risRnd = c.rnd();
                        if (streamSample(skyRes, kSkyLight, uv, risRnd, targetPdf, 1.0f / blended)) skyIdx = cidx;
                    }
                    finalizeResampling(skyRes, 1.0f, (float)nMis);
                    a.wb.candA[p] = make_float4(__int_as_float(sunIdx), sunRes.weightSum, sunRes.targetPdf, __int_as_float(skyIdx));
                    // one full 16-byte store (.zw: the local light's uv): two 8-byte halves leave every 32-byte sector half written,
                    // which costs a fill read of the sector at the L2 (133 MB per frame, ncu r2p: S1 read 374 MB for 232 MB of inputs)
                    a.wb.candB[p] = make_float4(skyRes.weightSum, skyRes.targetPdf, candUv.x, candUv.y);
                    // BSDF candidate: sample a direction and cast the BSDF-light ray (closesthit.cu:452-468)
                    f3 sampleDir, dummy; float brdfPdf; bool trans = false;
                    disneySample(c.rnd4(), s.normal, s.geoNormal, s.wo, s.albedo, s.metallic, s.translucency, s.roughness, sampleDir, dummy, brdfPdf, trans);
                    nf |= F_RIS;
                    if (brdfPdf > 0.0f)
                    {
                        nf |= F_RAY1;
                        a.wb.dir1[p] = make_float4(sampleDir.x, sampleDir.y, sampleDir.z, 0.0f);
                        want = prepareRay(a.grid, frontPos, sampleDir, 0.0f, (uint32_t)p, r);
                        // scenes with local lights trace this ray in closest-hit mode (which emissive face it hits matters,
                        // __closesthit__bsdf_light): its result then lands in hitPacked / hitT, which S1 has consumed by now
                        if (!want) { if (kLights) a.wb.hitPacked[p] = kHitMiss; else a.wb.vis1[p] = 0; }
                    }
                    if (a.enableRestir && gbufferPass) nf |= F_RESTIR;
                }
                // continuation (RayGen.cu:79-84, 146-165)
                const bool cont = mayContinue && !(bsdfPdf <= 0.0f || isNull(bsdfOverPdf));
                if (cont)
                {
                    nf |= F_CONT;
                    needSurf = true;
                    a.wb.nextD[p] = make_float4(bsdfWi.x, bsdfWi.y, bsdfWi.z, coneWidth);
                    a.wb.bop[p] = make_float4(bsdfOverPdf.x, bsdfOverPdf.y, bsdfOverPdf.z, 0.0f);
                }
                if (needSurf)
                {
                    a.wb.surfA[p] = make_float4(frontPos.x, frontPos.y, frontPos.z, distance);
                    a.wb.surfB[p] = make_float4(s.wo.x, s.wo.y, s.wo.z, __uint_as_float((uint32_t)face | (matIndex << 3)));
                    if (kTex)
                    {
                        a.wb.surfC[p] = make_float4(s.normal.x, s.normal.y, s.normal.z, s.roughness);
                        a.wb.surfD[p] = make_float4(s.albedo.x, s.albedo.y, s.albedo.z, s.metallic ? 1.0f : 0.0f);
                    }
                }
            }
        }
        if (depth == 0)
        {
            a.wb.rad[p] = make_float4(radiance.x, radiance.y, radiance.z, distance);
            if (nf & F_CONT) a.wb.thr[p] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        }
        else if (!isNull(radiance))
        {
            float4 acc = a.wb.rad[p];
            const float4 th = a.wb.thr[p];
            acc.x += th.x * radiance.x; acc.y += th.y * radiance.y; acc.z += th.z * radiance.z;
            a.wb.rad[p] = acc;
        }
        a.wb.pflag[p] = nf | ((uint32_t)c.randIdx << kRandShift) | ((uint32_t)newDiffuse << kDiffuseShift);
    }
    const unsigned pos = ctaReserve(want ? 1u : 0u, qCount);
    if (want) storePreparedRay(a.wb.queue, pos, r);
}

// ------------------------------------------------------------------------------------------------ S2
// RIS: classify the BSDF candidate, merge the three reservoirs, cast the winner's visibility ray (closesthit.cu:470-634).
template <bool kTex, bool kLights> __global__ void __launch_bounds__(kShadeThreads, VPT_S2_MINB) shade2Kernel(const __grid_constant__ TraceArgs a, int depth, const int *__restrict__ list,
                                                              const unsigned *__restrict__ listCount, unsigned *qCount)
{
    const int idx = blockIdx.x * kShadeThreads + threadIdx.x;
    const int n = list ? (int)__ldg(listCount) : a.partPaths;
    bool act = idx < n;
    const int p = act ? (list ? __ldg(list + idx) : partPath(a, idx)) : 0;
    uint32_t fl = act ? a.wb.pflag[p] : 0u;
    act = act && (fl & F_RIS);
    bool want = false;
    PreparedRay r;
    if (act)
    {
        const PathId id = pathId(a, p);
        Ctx c = makeCtx(a, id, (int)((fl >> kRandShift) & 0xffu));
        const Surface s = loadSurface<kTex>(a, p, depth);
        const f3 sunD = c.sunDir();
        const bool skipSun = (dot(s.normal, sunD) < 0.0f || dot(s.geoNormal, sunD) < 0.0f);
        const int nSun = skipSun ? 0 : 1;
        constexpr bool hasLights = kLights;
        const int nLocal = hasLights ? 8 : 0;
        const int nMis = nLocal + nSun + 2;
        const float localMisW = float(nLocal) / nMis, sunMisW = float(nSun) / nMis, skyMisW = 1.0f / nMis, brdfMisW = 1.0f / nMis;

        const float4 ca = __ldg(a.wb.candA + p), cb = __ldg(a.wb.candB + p);
        const int sunIdx = __float_as_int(ca.x), skyIdx = __float_as_int(ca.w);
        VptReservoir localRes = emptyReservoir();
        int localIdx = -1;
        if (hasLights)
        {
            const uint4 cc = __ldg(a.wb.candC + p);
            localIdx = (int)cc.x;
            localRes.weightSum = __uint_as_float(cc.y); localRes.targetPdf = __uint_as_float(cc.z); // already finalised by S1
            if (localIdx >= 0)
            {
                localRes.lightData = (uint32_t)localIdx | kLightValidBit;
                localRes.uvData = (uint32_t)(saturate(cb.z) * 0xffff) | ((uint32_t)(saturate(cb.w) * 0xffff) << 16);
            }
        }
        else
            finalizeResampling(localRes, 1.0f, (float)nMis);
        localRes.M = 1;
        VptReservoir sunRes = emptyReservoir();
        sunRes.weightSum = ca.y; sunRes.targetPdf = ca.z; sunRes.M = 1;
        if (sunIdx >= 0)
        {
            uint32_t ixu, iyu;
            a.divSunW.div((uint32_t)sunIdx, iyu, ixu);
            const int ix = (int)ixu, iy = (int)iyu;
            const f2 uv = {(ix + 0.5f) / float(a.sunW), (iy + 0.5f) / float(a.sunH)};
            sunRes.lightData = kSunLight | kLightValidBit;
            sunRes.uvData = (uint32_t)(saturate(uv.x) * 0xffff) | ((uint32_t)(saturate(uv.y) * 0xffff) << 16);
        }
        VptReservoir skyRes = emptyReservoir();
        skyRes.weightSum = cb.x; skyRes.targetPdf = cb.y; skyRes.M = 1;
        if (skyIdx >= 0)
        {
            uint32_t ixu, iyu;
            a.divSkyW.div((uint32_t)skyIdx, iyu, ixu);
            const int ix = (int)ixu, iy = (int)iyu;
            const f2 uv = {(ix + 0.5f) / float(a.skyW), (iy + 0.5f) / float(a.skyH)};
            skyRes.lightData = kSkyLight | kLightValidBit;
            skyRes.uvData = (uint32_t)(saturate(uv.x) * 0xffff) | ((uint32_t)(saturate(uv.y) * 0xffff) << 16);
        }

        // BSDF candidate
        VptReservoir brdfRes = emptyReservoir();
        LightSample brdfSample = noLight();
        const bool haveRay1 = (fl & F_RAY1) != 0;
        const f3 sampleDir = haveRay1 ? xyz(__ldg(a.wb.dir1 + p)) : F3(0.0f);
        const uint32_t hp1 = (haveRay1 && hasLights) ? a.wb.hitPacked[p] : kHitMiss;
        const bool ray1Hit = haveRay1 ? (hasLights ? hp1 != kHitMiss : a.wb.vis1[p] != 0) : true;
        {
            float lightSourcePdf = 0.0f;
            uint32_t lightIndex = kInvalidLight;
            f2 uv = {0, 0};
            LightSample cand = noLight();
            if (haveRay1 && ray1Hit && hasLights && (hp1 & 7u) < 6u)
            {
                // the BSDF ray ended on a voxel: an emissive one is a light (__closesthit__bsdf_light, closesthit.cu:854-901) — which of
                // the face's two triangles, and the hit's barycentrics
                const int lin = (int)(hp1 >> 3), face = (int)(hp1 & 7u);
                const VptMaterial *hm = a.materials + __ldg(a.blockToMaterial + __ldg(a.grid.idsLinear + lin));
                if (__ldg(&hm->isEmissive))
                {
                    uint32_t hxu, yzu, hzu, hyu;
                    a.grid.divW.div((uint32_t)lin, yzu, hxu);
                    a.grid.divD.div(yzu, hyu, hzu);
                    f3 A, eu, ev;
                    faceFrameDev(face, (int)hxu, (int)hyu, (int)hzu, A, eu, ev);
                    const f3 P = hitPoint((int)hxu, (int)hyu, (int)hzu, face, a.wb.hitT[p], s.pos, sampleDir);
                    const float fs = dot(P - A, eu), ft = dot(P - A, ev);
                    const int tri = (fs + ft <= 1.0f) ? 0 : 1;
                    const f2 bary = tri == 0 ? f2{fs, ft} : f2{1.0f - fs, 1.0f - ft};
                    const int li = findLight(a.lv, (uint32_t)lin, face, tri);
                    if (li >= 0 && li < a.lv.numLights)
                    {
                        lightIndex = (uint32_t)li;
                        const float sq = 1.0f - (1.0f - bary.x - bary.y); // InverseTriangleSample (LinearMath.h:2059-2064)
                        uv = {sq * sq, bary.y / sq};
                        cand = triLightSample(createTriLight(a.lv.lights + li), uv, s.pos);
                        lightSourcePdf = __ldg(&a.lv.alias[li].p);
                    }
                }
            }
            else if (haveRay1 && !ray1Hit)
            {
                if (equalAreaMapConeInv(uv, sunD, sampleDir, a.sunCosThetaMax))
                {
                    lightIndex = kSunLight;
                    int sx = (int)(uv.x * a.sunW - 0.5f), sy = (int)(uv.y * a.sunH - 0.5f);
                    if (sx >= a.sunW) sx %= a.sunW;
                    if (sx < 0) sx = a.sunW - ((-sx) % a.sunW);
                    sy = clampi(sy, 0, a.sunH - 1);
                    const int cidx = sy * a.sunW + sx;
                    cand = c.createSunLightSample(cidx);
                    cand.position = sampleDir;
                    lightSourcePdf = __ldg(&a.sunAlias[cidx].p);
                }
                else
                {
                    lightIndex = kSkyLight;
                    uv = equalAreaSphereMapInv(sampleDir);
                    const int kx = (int)(uv.x * a.skyW - 0.5f), ky = (int)(uv.y * a.skyH - 0.5f);
                    int cidx = ky * a.skyW + kx;
                    cidx = clampi(cidx, 0, a.skyW * a.skyH - 1);
                    cand = c.createSkyLightSample(cidx);
                    cand.position = sampleDir;
                    lightSourcePdf = __ldg(&a.skyAlias[cidx].p);
                }
            }
            if (lightSourcePdf != 0.0f)
            {
                const float misW = (lightIndex == kSkyLight) ? skyMisW : ((lightIndex == kSunLight) ? sunMisW : localMisW);
                float blended;
                const float targetPdf = c.evalCandidate(s, cand, lightSourcePdf, misW, brdfMisW, &blended);
                const float risRnd = c.rnd();
                if (streamSample(brdfRes, lightIndex, uv, risRnd, targetPdf, 1.0f / blended)) brdfSample = cand;
            }
        }
        finalizeResampling(brdfRes, 1.0f, (float)nMis);
        brdfRes.M = 1;

        VptReservoir ris = emptyReservoir();
        combineReservoirs(ris, localRes, 0.5f, localRes.targetPdf);
        const float r0 = c.rnd(); const bool selSun = combineReservoirs(ris, sunRes, r0, sunRes.targetPdf);
        const float r1 = c.rnd(); const bool selSky = combineReservoirs(ris, skyRes, r1, skyRes.targetPdf);
        const float r2 = c.rnd(); const bool selBrdf = combineReservoirs(ris, brdfRes, r2, brdfRes.targetPdf);
        finalizeResampling(ris, 1.0f, 1.0f);
        ris.M = 1;
        LightSample lightSample = noLight();
        if (selBrdf) lightSample = brdfSample;
        else if (selSky) { if (skyIdx >= 0) lightSample = c.createSkyLightSample(skyIdx); }
        else if (selSun) { if (sunIdx >= 0) lightSample = c.createSunLightSample(sunIdx); }
        else if (localIdx >= 0) lightSample = triLightSample(createTriLight(a.lv.lights + localIdx), f2{cb.z, cb.w}, s.pos); // localSample, recomputed

        fl &= ~(0xffu << kRandShift);
        fl |= ((uint32_t)c.randIdx << kRandShift);
        if (lightSample.lightType != LightInvalid)
        {
            fl |= F_LIGHTOK;
            if (isValidReservoir(ris))
            {
                fl |= F_VISHAVE1;
                // identical (origin, direction) -> identical result: reuse the BSDF-candidate ray instead of re-tracing it
                if (haveRay1 && lightSample.lightType != LightLocalTriangle && sameDir(sampleDir, lightSample.position))
                {
                    if (!ray1Hit) fl |= F_VIS;
                    else { ris.lightData = 0; ris.weightSum = 0; }
                }
                else
                {
                    want = prepareRay(a.grid, s.pos, lightRayDir<kLights>(lightSample, s.pos), 0.0f, (uint32_t)p, r, lightRayTmax<kLights>(lightSample, s.pos, 0.0f));
                    if (want) fl |= F_RAY2;
                    else fl |= F_VIS; // never enters the grid: visible
                }
            }
        }
        a.wb.ris[p] = make_uint4(ris.lightData, ris.uvData, __float_as_uint(ris.weightSum), __float_as_uint(ris.targetPdf));
        storeLight(a, a.wb.lightA, a.wb.lightB, p, lightSample);
        a.wb.pflag[p] = fl;
    }
    const unsigned pos = ctaReserve(want ? 1u : 0u, qCount);
    if (want) storePreparedRay(a.wb.queue, pos, r);
}

VPT_DEV VptReservoir loadRis(const TraceArgs &a, int p)
{
    const uint4 v = __ldg(a.wb.ris + p);
    VptReservoir r;
    r.lightData = v.x; r.uvData = v.y; r.weightSum = __uint_as_float(v.z); r.targetPdf = __uint_as_float(v.w); r.M = 1;
    return r;
}
template <bool kLights> VPT_DEV VptReservoir loadPrevReservoir(const TraceArgs &a, int ix, int iy, float mCap)
{
    VptReservoir pr = loadReservoir(a.resPrev + (size_t)iy * a.width + ix);
    if (kLights && a.lv.stateDirty) // LoadDIReservoir (Restir.h:48-79): the light list changed since the reservoir was written
    {
        const uint32_t prevIdx = pr.lightData & kLightIndexMask;
        if (prevIdx < kSunLight && a.lv.prevNumLights > 0 && prevIdx < (uint32_t)a.lv.prevNumLights)
        {
            const int curIdx = __ldg(a.lv.prevToCur + prevIdx);
            if (curIdx < 0 || curIdx >= a.lv.numLights) pr = emptyReservoir();
            else pr.lightData = (pr.lightData & ~kLightIndexMask) | (uint32_t)curIdx;
        }
    }
    if (isnan(pr.weightSum) || isinf(pr.weightSum)) pr = emptyReservoir();
    if (pr.M > mCap) pr.M = mCap;
    return pr;
}

// ------------------------------------------------------------------------------------------------ S3
// Temporal ReSTIR: candidates from the previous frame + the bias-correction rays (closesthit.cu:636-760).
#ifndef VPT_S3_UNROLL
#define VPT_S3_UNROLL 0
#endif
// Measured and not kept: prefetch.global.L2 / .L1 of the three candidates' 18 scattered sectors before the candidate loop (the loop
// visits them one after the other: three dependent round trips): S3 327 -> 372 us — the prefetches are extra LSU work and the
// sectors are evicted or still in flight when the loop gets to them; unrolling the loop: slower as well (register pressure).
#if VPT_S3_UNROLL
#define VPT_S3_LOOP _Pragma("unroll")
#else
#define VPT_S3_LOOP _Pragma("unroll 1")
#endif
template <bool kTex, bool kLights> __global__ void __launch_bounds__(kShadeThreads, VPT_S3_MINB) shade3Kernel(const __grid_constant__ TraceArgs a, unsigned *qCount)
{
    const int idx = blockIdx.x * kShadeThreads + threadIdx.x;
    const int p = a.slotBase + idx; // sample 0 of the wave: path == slot
    bool act = idx < a.partSlots;
    uint32_t fl = act ? a.wb.pflag[p] : 0u;
    act = act && (fl & F_RIS) && (fl & F_RESTIR);
    unsigned nWant = 0;
    PreparedRay rays[3];
    if (act)
    {
        const PathId id = pathId(a, p);
        Ctx c = makeCtx(a, id, (int)((fl >> kRandShift) & 0xffu));
        const Surface s = loadSurface<kTex>(a, p, 0);
        VptReservoir ris = loadRis(a, p);
        LightSample lightSample = loadLight(a.wb.lightA, a.wb.lightB, p);
        // resolve the RIS winner's visibility (closesthit.cu:602-634)
        if (fl & F_RAY2)
        {
            if (a.wb.vis2[p] == 0) fl |= F_VIS;
            else { ris.lightData = 0; ris.weightSum = 0; }
        }
        const VptCamera &pc = a.prevCam;
        const f3 pcPos = {pc.pos[0], pc.pos[1], pc.pos[2]};
        VptReservoir restir = emptyReservoir();
        combineReservoirs(restir, ris, 0.5f, ris.targetPdf);
        const f3 prevWorldPos = s.pos;
        const f2 prevUV = c.worldDirectionToUV(pc, normalize(prevWorldPos - pcPos));
        const int prevPx = (int)(prevUV.x * pc.resolution[0]), prevPy = (int)(prevUV.y * pc.resolution[1]);
        const float expectedPrevDepth = distance(prevWorldPos, pcPos);
        constexpr int nTemporal = 3;
        constexpr float mCap = 20.0f;
        int offx[nTemporal], offy[nTemporal];
        offx[0] = prevPx - id.px; offy[0] = prevPy - id.py;
        {
            const f2 dsk = concentricSampleDisk(c.rnd2()) * 64.0f;
            offx[1] = prevPx - id.px + (int)dsk.x; offy[1] = prevPy - id.py + (int)dsk.y;
        }
        {
            const f2 dsk = concentricSampleDisk(c.rnd2()) * 64.0f;
            offx[2] = (int)dsk.x; offy[2] = (int)dsk.y;
        }
        unsigned cached = 0, rayMask = 0;
        int selectedLoopIdx = -1;
        float pm[3] = {0.0f, 0.0f, 0.0f};
VPT_S3_LOOP
        for (int i = 0; i < nTemporal; ++i)
        {
            int ix = id.px + offx[i], iy = id.py + offy[i];
            clampIntoView(ix, iy, a.width, a.height);
            Surface ts;
            // the candidate's reservoir travels with its surface planes (same pixel): requested before the validation
            const bool inView = ix >= 0 && iy >= 0 && ix < a.width && iy < a.height; // getPrevSurface's own test
            VptReservoir pr = inView ? loadPrevReservoir<kLights>(a, ix, iy, mCap) : emptyReservoir();
            const float mIn = pr.M;
            if (!c.getPrevSurface(ts, ix, iy)) continue;
            const bool nOk = dot(s.normal, ts.geoNormal) >= 0.5f;
            const bool dOk = fabsf(expectedPrevDepth - ts.depth) <= 0.1f * fmaxr(expectedPrevDepth, ts.depth);
            const bool rOk = fabsf(s.roughness - ts.roughness) <= 0.5f * fmaxr(s.roughness, ts.roughness);
            if (!(nOk && dOk && rOk)) continue;
            cached |= (1u << i);
            float neighborWeight = 0;
            LightSample cand = noLight();
            if (isValidReservoir(pr))
            {
                if (!c.template lightSampleFromReservoir<kLights>(cand, pr, s.pos)) pr = emptyReservoir();
                neighborWeight = c.targetPdfForSurface(cand, s);
            }
            pm[i] = mIn; // the candidate's (capped) M as loaded: the bias-correction pass below needs nothing else of the reservoir
            if (combineReservoirs(restir, pr, c.rnd(), neighborWeight)) { lightSample = cand; selectedLoopIdx = i; }
        }
        float ps[3] = {0.0f, 0.0f, 0.0f};
        const bool valid = isValidReservoir(restir);
        if (valid)
        {
VPT_S3_LOOP
            for (int i = 0; i < nTemporal; ++i)
            {
                if ((cached & (1u << i)) == 0) continue;
                int ix = id.px + offx[i], iy = id.py + offy[i];
                clampIntoView(ix, iy, a.width, a.height);
                Surface ts;
                c.getPrevSurface(ts, ix, iy);
                LightSample atNeighbor = noLight();
                c.template lightSampleFromReservoir<kLights>(atNeighbor, restir, ts.pos);
                ps[i] = c.targetPdfForSurface(atNeighbor, ts);
                if (ps[i] > 0 && !(i == 0 && i == selectedLoopIdx))
                {
                    const float extraRayOffset = 0.01f + 0.01f * ts.depth;
                    PreparedRay pr_;
                    // towards `lightSample`, the sample at the CURRENT surface, as in the reference (closesthit.cu:743-744)
                    if (prepareRay(a.grid, ts.pos, lightRayDir<kLights>(lightSample, ts.pos), extraRayOffset, (uint32_t)(p * 3 + i), pr_, lightRayTmax<kLights>(lightSample, ts.pos, extraRayOffset)))
                    {
                        rays[nWant++] = pr_;
                        rayMask |= (1u << i);
                    }
                    else a.wb.vis3[p * 3 + i] = 0;
                }
            }
        }
        a.wb.rstA[p] = make_uint4(restir.lightData, restir.uvData, __float_as_uint(restir.weightSum), __float_as_uint(restir.targetPdf));
        a.wb.rstB[p] = make_uint4(__float_as_uint(restir.M), cached | ((uint32_t)(selectedLoopIdx + 1) << 3) | (rayMask << 5) | (valid ? 256u : 0u), 0u, 0u);
        a.wb.psA[p] = make_float4(ps[0], ps[1], ps[2], 0.0f);
        a.wb.psB[p] = make_float4(pm[0], pm[1], pm[2], 0.0f);
        storeLight(a, a.wb.light2A, a.wb.light2B, p, lightSample);
        fl &= ~(0xffu << kRandShift);
        fl |= ((uint32_t)c.randIdx << kRandShift);
        a.wb.pflag[p] = fl;
    }
    const unsigned pos = ctaReserve(nWant, qCount);
    for (unsigned i = 0; i < nWant; ++i) storePreparedRay(a.wb.queue, pos + i, rays[i]);
}

// ------------------------------------------------------------------------------------------------ S4
// Bias-corrected normalisation and the final visibility ray (closesthit.cu:760-820).
template <bool kLights> __global__ void __launch_bounds__(kShadeThreads) shade4Kernel(const __grid_constant__ TraceArgs a, unsigned *qCount)
{
    const int idx = blockIdx.x * kShadeThreads + threadIdx.x;
    const int p = a.slotBase + idx;
    bool act = idx < a.partSlots;
    uint32_t fl = act ? a.wb.pflag[p] : 0u;
    act = act && (fl & F_RIS) && (fl & F_RESTIR);
    bool want = false;
    PreparedRay r;
    if (act)
    {
        const uint4 ra = a.wb.rstA[p], rb = a.wb.rstB[p];
        VptReservoir restir;
        restir.lightData = ra.x; restir.uvData = ra.y; restir.weightSum = __uint_as_float(ra.z); restir.targetPdf = __uint_as_float(ra.w);
        restir.M = __uint_as_float(rb.x);
        const uint32_t meta = rb.y;
        const unsigned cached = meta & 7u, rayMask = (meta >> 5) & 7u;
        const int selectedLoopIdx = (int)((meta >> 3) & 3u) - 1;
        const LightSample lightSample = loadLight(a.wb.light2A, a.wb.light2B, p);
        if (meta & 256u)
        {
            const float4 psv = a.wb.psA[p], pmv = a.wb.psB[p];
            const float psArr[3] = {psv.x, psv.y, psv.z}, pmArr[3] = {pmv.x, pmv.y, pmv.z};
            float pi = restir.targetPdf, piSum = restir.targetPdf * 1;
#pragma unroll
            for (int i = 0; i < 3; ++i)
            {
                if ((cached & (1u << i)) == 0) continue;
                float ps = psArr[i];
                if ((rayMask & (1u << i)) && a.wb.vis3[p * 3 + i] != 0) ps = 0.0f;
                if (selectedLoopIdx == i) pi = ps;
                piSum += ps * pmArr[i];
            }
            finalizeResampling(restir, pi, piSum);
        }
        const bool risVis = (fl & F_VIS) != 0;
        fl &= ~F_VIS; // from here on F_VIS is the visibility of the ReSTIR sample
        if (lightSample.lightType != LightInvalid)
        {
            const f3 dirRis = xyz(__ldg(a.wb.lightA + p));
            const bool haveRay1 = (fl & F_RAY1) != 0;
            bool known = false, visible = false;
            // (a local light's `position` is a point, an environment light's a direction: equal bits = the same ray from this surface)
            if ((fl & F_VISHAVE1) && sameDir(dirRis, lightSample.position)) { known = true; visible = risVis; }
            else if (haveRay1 && lightSample.lightType != LightLocalTriangle && sameDir(xyz(__ldg(a.wb.dir1 + p)), lightSample.position))
            {
                known = true;
                visible = kLights ? a.wb.hitPacked[p] == kHitMiss : a.wb.vis1[p] == 0; // ray #1 ran in closest-hit mode when lights exist
            }
            if (known)
            {
                if (visible) fl |= F_VIS;
                else { restir.lightData = 0; restir.weightSum = 0; }
            }
            else
            {
                const float4 sa = __ldg(a.wb.surfA + p);
                want = prepareRay(a.grid, xyz(sa), lightRayDir<kLights>(lightSample, xyz(sa)), 0.0f, (uint32_t)p, r, lightRayTmax<kLights>(lightSample, xyz(sa), 0.0f));
                if (want) fl |= F_RAY5;
                else fl |= F_VIS;
            }
        }
        a.wb.rstA[p] = make_uint4(restir.lightData, restir.uvData, __float_as_uint(restir.weightSum), __float_as_uint(restir.targetPdf));
        a.wb.pflag[p] = fl;
    }
    const unsigned pos = ctaReserve(want ? 1u : 0u, qCount);
    if (want) storePreparedRay(a.wb.queue, pos, r);
}

// ------------------------------------------------------------------------------------------------ S5
// Shade with the surviving reservoir, store it, accumulate, spawn the continuation ray (closesthit.cu:822-851, RayGen.cu:71-84).
template <bool kTex, bool kLights> __global__ void __launch_bounds__(kShadeThreads, VPT_S5_MINB) shade5Kernel(const __grid_constant__ TraceArgs a, int depth, const int *__restrict__ list,
                                                              const unsigned *__restrict__ listCount, int *__restrict__ nextList,
                                                              unsigned *nextCount, unsigned *qCount)
{
    const int idx = blockIdx.x * kShadeThreads + threadIdx.x;
    const int n = list ? (int)__ldg(listCount) : a.partPaths;
    bool act = idx < n;
    const int p = act ? (list ? __ldg(list + idx) : partPath(a, idx)) : 0;
    uint32_t fl = act ? a.wb.pflag[p] : 0u;
    act = act && (fl & (F_RIS | F_CONT));
    bool cont = false, want = false;
    PreparedRay r;
    if (act)
    {
        const PathId id = pathId(a, p);
        if (fl & F_RIS)
        {
            const Surface s = loadSurface<kTex>(a, p, depth);
            const bool useRestir = (fl & F_RESTIR) != 0;
            VptReservoir shading;
            LightSample lightSample;
            bool visible;
            if (useRestir)
            {
                const uint4 ra = a.wb.rstA[p], rb = a.wb.rstB[p];
                shading.lightData = ra.x; shading.uvData = ra.y; shading.weightSum = __uint_as_float(ra.z); shading.targetPdf = __uint_as_float(ra.w);
                shading.M = __uint_as_float(rb.x);
                lightSample = loadLight(a.wb.light2A, a.wb.light2B, p);
                visible = (fl & F_VIS) != 0;
                if (fl & F_RAY5)
                {
                    visible = a.wb.vis4[p] == 0;
                    if (!visible) { shading.lightData = 0; shading.weightSum = 0; }
                }
            }
            else
            {
                shading = loadRis(a, p);
                lightSample = loadLight(a.wb.lightA, a.wb.lightB, p);
                visible = (fl & F_VIS) != 0;
                if (fl & F_RAY2)
                {
                    visible = a.wb.vis2[p] == 0;
                    if (!visible) { shading.lightData = 0; shading.weightSum = 0; }
                }
            }
            if (lightSample.lightType != LightInvalid && isValidReservoir(shading) && visible)
            {
                const f3 sampleDir = lightRayDir<kLights>(lightSample, s.pos);
                const f3 albedo = (depth == 0) ? F3(1.0f) : s.albedo; // first hit is demodulated (closesthit.cu:301)
                f3 bsdf; float pdf;
                disneyEvaluate(s.normal, s.geoNormal, sampleDir, s.wo, albedo, s.metallic, s.roughness, bsdf, pdf);
                const float cosTheta = fmaxf(0.0f, dot(sampleDir, s.normal));
                const f3 shadowRad = bsdf * cosTheta * lightSample.radiance * shading.weightSum / lightSample.solidAnglePdf;
                float4 acc = a.wb.rad[p];
                if (depth == 0) { acc.x += shadowRad.x; acc.y += shadowRad.y; acc.z += shadowRad.z; }
                else
                {
                    const float4 th = a.wb.thr[p];
                    acc.x += th.x * shadowRad.x; acc.y += th.y * shadowRad.y; acc.z += th.z * shadowRad.z;
                }
                a.wb.rad[p] = acc;
            }
            if (depth == 0 && id.k == a.ownerSample)
                storeReservoir(a.resCur + (size_t)id.py * a.width + id.px, a.enableRestir ? shading : emptyReservoir());
        }
        if ((fl & F_CONT) && nextList) // the last depth round traces nothing further: no continuation state, no reservations
        {
            cont = true;
            const float4 sa = __ldg(a.wb.surfA + p), nd = __ldg(a.wb.nextD + p), bo = __ldg(a.wb.bop + p);
            float4 th = a.wb.thr[p];
            th.x *= bo.x; th.y *= bo.y; th.z *= bo.z;
            a.wb.thr[p] = th;
            a.wb.org[p] = make_float4(sa.x, sa.y, sa.z, 0.0f);
            a.wb.dirT[p] = make_float4(nd.x, nd.y, nd.z, nd.w); // .w = accumulated ray-cone width
            want = prepareRay(a.grid, xyz(sa), xyz(nd), 0.0f, (uint32_t)p, r);
            if (!want) { a.wb.hitPacked[p] = kHitMiss; a.wb.hitT[p] = kRayMax; }
            a.wb.pflag[p] = F_LIVE | (fl & (0xffu << kRandShift)) | (fl & (0xfu << kDiffuseShift));
        }
    }
    if (nextList) // launch-uniform
    {
        const unsigned lpos = ctaReserve<1>(cont ? 1u : 0u, nextCount);
        if (cont) nextList[lpos] = p;
        const unsigned pos = ctaReserve(want ? 1u : 0u, qCount);
        if (want) storePreparedRay(a.wb.queue, pos, r);
    }
}

// ------------------------------------------------------------------------------------------------ accumulate
// Per pixel: sum of the wave's samples in sample order (RayGen.cu:175-181 NaN guard per sample), depth from sample 0.
__global__ void __launch_bounds__(kShadeThreads) accumulateKernel(const __grid_constant__ TraceArgs a)
{
    const int idx = blockIdx.x * kShadeThreads + threadIdx.x;
    if (idx >= a.partSlots) return;
    const int slot = a.slotBase + idx;
    const PathId id = pathId(a, slot);
    if (!id.inImage) return;
    const size_t pix = (size_t)id.py * a.width + id.px;
    const bool first = a.waveFirst == 0;
    const bool haveDepth = first && a.sampleBegin == a.ownerSample;
    float4 prev = first ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : a.illumination[pix];
    f3 sum = {prev.x, prev.y, prev.z};
    float depth0 = prev.w;
    for (int sl = 0; sl < a.samplesInWave; ++sl)
    {
        const float4 v = __ldg(a.wb.rad + (size_t)sl * a.nSlots + slot);
        f3 r = {v.x, v.y, v.z};
        if (isnan(r.x) || isnan(r.y) || isnan(r.z)) r = F3(0.5f);
        sum += r;
        if (sl == 0 && haveDepth) depth0 = v.w;
    }
    if (haveDepth) a.cur.depth[pix] = depth0;
    // vpt_render (unsharded): the division by spp of vpt_resolve, fused into the last wave's accumulate
    if (a.resolveSpp > 0.0f) { sum.x = sum.x / a.resolveSpp; sum.y = sum.y / a.resolveSpp; sum.z = sum.z / a.resolveSpp; }
    a.illumination[pix] = make_float4(sum.x, sum.y, sum.z, depth0);
}

__global__ void resolveKernel(float4 *illum, int npix, float spp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 v = illum[i];
    v.x = v.x / spp; v.y = v.y / spp; v.z = v.z / spp;
    illum[i] = v;
}

// ------------------------------------------------------------------------------------------------ host
static size_t alignUp(size_t v) { return (v + 255) & ~(size_t)255; }
template <typename T> static void carve(char *&cur, T *&ptr, size_t count) { ptr = reinterpret_cast<T *>(cur); cur += alignUp(count * sizeof(T)); }
static void carveAll(char *base, WaveBuffers &wb, int nSlots, int samples)
{
    const size_t N = (size_t)nSlots * samples, S = (size_t)nSlots;
    char *cur = base;
    carve(cur, wb.dirT, N); carve(cur, wb.org, N); carve(cur, wb.hitT, N); carve(cur, wb.hitPacked, N);
    carve(cur, wb.surfA, N); carve(cur, wb.surfB, N); carve(cur, wb.surfC, N); carve(cur, wb.surfD, N); carve(cur, wb.pflag, N);
    carve(cur, wb.candA, N); carve(cur, wb.candB, N); carve(cur, wb.candC, N); carve(cur, wb.dir1, N);
    carve(cur, wb.ris, N); carve(cur, wb.lightA, N); carve(cur, wb.lightB, N);
    carve(cur, wb.vis1, N); carve(cur, wb.vis2, N); carve(cur, wb.vis4, N);
    carve(cur, wb.rad, N); carve(cur, wb.thr, N); carve(cur, wb.nextD, N); carve(cur, wb.bop, N);
    carve(cur, wb.rstA, S); carve(cur, wb.rstB, S); carve(cur, wb.light2A, S); carve(cur, wb.light2B, S);
    carve(cur, wb.psA, S); carve(cur, wb.psB, S); carve(cur, wb.vis3, S * 3);
    const size_t q = (N > S * 3 ? N : S * 3) + 64; // parts use disjoint sub-ranges
    carve(cur, wb.queue, q * 3);
    carve(cur, wb.listA, N); carve(cur, wb.listB, N);
    carve(cur, wb.cnt, (size_t)kCntWords * 2);
}
size_t waveWorkspaceBytes(int nSlots, int samplesInWave)
{
    WaveBuffers wb;
    carveAll(nullptr, wb, nSlots, samplesInWave);
    return (size_t)(reinterpret_cast<char *>(wb.cnt) - (char *)nullptr) + alignUp(2 * kCntWords * sizeof(unsigned));
}
void waveCarve(WaveWorkspace &ws, int nSlots, int samplesInWave)
{
    carveAll(static_cast<char *>(ws.arena), ws.wb, nSlots, samplesInWave);
    ws.nSlots = nSlots; ws.maxSamplesInWave = samplesInWave;
}

cudaError_t launchTrace(TraceArgs &a, int maxSamplesInWave, cudaStream_t s, const TraceStreams *ts, int smCount, size_t smemOptIn, int *launches,
                        TraceProfile *prof)
{
    int nl = 0;
    const bool timing = prof && prof->enabled;
    // two parts on two side streams (overlap), or one part after the other on the caller's stream when every launch is timed
    const bool overlap = ts && ts->part[0] && !timing;
    if (timing) { prof->n = 0; cudaEventRecord(prof->ev[0], s); }
    auto mark = [&](int kind, cudaStream_t st) {
        if (timing && prof->n < TraceProfile::kMax) { prof->kind[prof->n] = kind; cudaEventRecord(prof->ev[prof->n + 1], st); ++prof->n; }
    };
    int shardSamples = a.sampleBegin < a.spp ? (a.spp - a.sampleBegin + a.sampleStep - 1) / a.sampleStep : 0;
    if (a.sampleLimit > 0 && shardSamples > a.sampleLimit) shardSamples = a.sampleLimit;
    a.occInSmem = ((size_t)a.grid.occWords * 4 + 1024 <= smemOptIn) ? 1 : 0;
    const bool smem = a.occInSmem != 0, stats = a.countSteps != 0;
    const WaveBuffers wb0 = a.wb;
    const int nTiles = a.nSlots / 32;
    // Parts exist so two streams can overlap one part's DDA with the other's shading. Measured on B200 (r1 w5): no
    // gain — a shading kernel's thousands of pending CTAs keep back-filling the SMs, so the other part's 1024-thread
    // DDA CTA (173 KiB of shared memory) only gets in at the tail. One part unless the caller passes side streams.
    const int nParts = (overlap && nTiles >= 2) ? 2 : 1;
    const bool tex = a.nTextures > 0, lights = a.lv.numLights > 0;
#define KTEX(k) (tex ? (lights ? k<true, true> : k<true, false>) : (lights ? k<false, true> : k<false, false>))
#define VPT_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
    for (int first = 0; first < shardSamples; first += maxSamplesInWave)
    {
        a.waveFirst = first;
        a.samplesInWave = shardSamples - first < maxSamplesInWave ? shardSamples - first : maxSamplesInWave;
        a.nPaths = a.nSlots * a.samplesInWave;
        const bool restirWave = a.enableRestir && a.sampleBegin == a.ownerSample && first == 0;
        VPT_TRY(cudaMemsetAsync(wb0.cnt, 0, 2 * kCntWords * sizeof(unsigned), s));
        if (overlap) VPT_TRY(cudaEventRecord(ts->fork, s));
        size_t queueBase = 0, listBase = 0;
        for (int part = 0; part < nParts; ++part)
        {
            cudaStream_t st = overlap ? ts->part[part] : s;
            if (overlap) VPT_TRY(cudaStreamWaitEvent(st, ts->fork, 0));
            const int tile0 = part == 0 ? 0 : (nTiles + 1) / 2, tile1 = (part == 0 && nParts == 2) ? (nTiles + 1) / 2 : nTiles;
            a.slotBase = tile0 * 32; a.partSlots = (tile1 - tile0) * 32; a.partPaths = a.partSlots * a.samplesInWave;
            a.divPartSlots = makeFastDiv((uint32_t)a.partSlots);
            a.wb = wb0;
            a.wb.queue = wb0.queue + queueBase * 3;
            a.wb.listA = wb0.listA + listBase; a.wb.listB = wb0.listB + listBase;
            a.wb.cnt = wb0.cnt + part * kCntWords;
            queueBase += (size_t)(a.partPaths > a.partSlots * 3 ? a.partPaths : a.partSlots * 3);
            listBase += (size_t)a.partPaths;
            unsigned *cnt = a.wb.cnt;
            auto dda = [&](int pair, bool closest, uint8_t *vis, bool prevWorld = false, bool nearEnd = false) -> cudaError_t {
                DdaArgs d;
                d.queue = a.wb.queue; d.count = cnt + 2 * pair; d.cursor = cnt + 2 * pair + 1;
                d.hitT = a.wb.hitT; d.hitPacked = a.wb.hitPacked; d.vis = vis; d.grid = a.grid; d.counters = a.counters;
                if (prevWorld && a.occPrev) { d.grid.occ = a.occPrev; d.grid.upH = a.upHPrev; } // closesthit.cu:736-755: prevTopObject
                ++nl;
                cudaError_t e = launchDda(d, closest, smem, stats, a.lv.numLights > 0, nearEnd, st, smCount);
                mark(0, st);
                return e;
            };
            const unsigned gridPaths = (unsigned)((a.partPaths + kShadeThreads - 1) / kShadeThreads);
            const unsigned gridSlots = (unsigned)((a.partSlots + kShadeThreads - 1) / kShadeThreads);
            int pair = 0;
            genKernel<<<gridPaths, kShadeThreads, 0, st>>>(a, cnt + 2 * pair); ++nl; mark(1, st);
            for (int depth = 0; depth < a.depthRounds; ++depth)
            {
                const int *list = depth == 0 ? nullptr : (depth & 1 ? a.wb.listA : a.wb.listB);
                int *nextList = (depth + 1 < a.depthRounds) ? ((depth + 1) & 1 ? a.wb.listA : a.wb.listB) : nullptr;
                const unsigned *listCount = cnt + kCntList + depth;
                VPT_TRY(dda(pair, true, nullptr)); ++pair;
                KTEX(shade1Kernel)<<<gridPaths, kShadeThreads, 0, st>>>(a, depth, list, listCount, cnt + 2 * pair); ++nl; mark(1, st);
                VPT_TRY(dda(pair, lights, a.wb.vis1)); ++pair; // BSDF-candidate ray: which emissive face it hits matters when lights exist
                KTEX(shade2Kernel)<<<gridPaths, kShadeThreads, 0, st>>>(a, depth, list, listCount, cnt + 2 * pair); ++nl; mark(1, st);
                VPT_TRY(dda(pair, false, a.wb.vis2)); ++pair;
                if (depth == 0 && restirWave)
                {
                    KTEX(shade3Kernel)<<<gridSlots, kShadeThreads, 0, st>>>(a, cnt + 2 * pair); ++nl; mark(1, st);
                    VPT_TRY(dda(pair, false, a.wb.vis3, true, true)); ++pair; // the bias rays start extraRayOffset along the ray (tmin > 0)
                    (lights ? shade4Kernel<true> : shade4Kernel<false>)<<<gridSlots, kShadeThreads, 0, st>>>(a, cnt + 2 * pair); ++nl; mark(1, st);
                    VPT_TRY(dda(pair, false, a.wb.vis4)); ++pair;
                }
                KTEX(shade5Kernel)<<<gridPaths, kShadeThreads, 0, st>>>(a, depth, list, listCount, nextList, cnt + kCntList + depth + 1, cnt + 2 * pair); ++nl; mark(1, st);
            }
            {
                TraceArgs acc = a;
                if (first + maxSamplesInWave < shardSamples) acc.resolveSpp = 0.0f; // only the last wave resolves
                accumulateKernel<<<gridSlots, kShadeThreads, 0, st>>>(acc); ++nl; mark(1, st);
            }
            VPT_TRY(cudaGetLastError());
            if (overlap)
            {
                VPT_TRY(cudaEventRecord(ts->join[part], st));
                VPT_TRY(cudaStreamWaitEvent(s, ts->join[part], 0));
            }
        }
        a.wb = wb0;
    }
#undef VPT_TRY
    if (launches) *launches = nl;
    return cudaGetLastError();
}
cudaError_t launchResolve(float4 *illum, int npix, float spp, cudaStream_t s)
{
    resolveKernel<<<(npix + 255) / 256, 256, 0, s>>>(illum, npix, spp);
    return cudaGetLastError();
}

} // namespace vpt
