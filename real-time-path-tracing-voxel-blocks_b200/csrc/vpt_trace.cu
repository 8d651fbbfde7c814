// Per-pixel path tracing kernel for sm_100a: replaces the reference's OptiX pipeline
// (__raygen__pathtracer / __closesthit__radiance / __miss__radiance,
//  /root/reference/renderer/shaders/RayGen.cu:102-181, closesthit.cu:10-852, miss.cu:9-82)
// with hand-written SIMT traversal — B200 has no RT cores.
//
//  * Traversal: Amanatides–Woo DDA over the voxel grid with the comparison order and the
//    tMax += tDelta accumulation of VoxelEngine::performRayTraversal (voxelengine/VoxelEngine.cu:1040-1166),
//    generalised with an entry clip, [tmin,tmax) and the entered face id.
//  * Grid residency: the 1-bit/voxel occupancy mask (64 KiB for the 16-chunk world) is staged whole in
//    shared memory once per CTA; a DDA step is one LDS + bit test, the 1-byte block id is fetched from
//    L2 only on a hit. Worlds whose mask exceeds shared memory walk the mask through L1/L2 instead.
//  * Scheduling: persistent CTAs (2 x 512 threads per SM = 32 warps at 64 registers), each warp claims 8x4-pixel tiles from a global atomic
//    counter (warp-granular dynamic load balance, no CTA barrier in the loop). A warp's lanes are an
//    8x4 pixel block so primary rays stay coherent and every G-buffer float4 row store is a full 128-byte line.
//  * Shading: Disney BSDF (Bsdf.h:371-617), RIS over sun / sky / BSDF candidates and temporal ReSTIR with
//    bias-correction rays (closesthit.cu:318-851, Restir.h). This TU is compiled -fmad=false so every
//    + - * / sqrt matches the CPU oracle bit for bit; only libm transcendentals differ (ulps).
#include "vpt_kernels.h"
#include "vpt_math.cuh"

namespace vpt {

constexpr float kSpawnEps = 0.0009765625f; // 2^-10
constexpr uint32_t kLightValidBit = 0x80000000u, kLightIndexMask = 0x7FFFFFFFu;
constexpr uint32_t kInvalidLight = 0x7FFFFFFFu, kSkyLight = 0x7FFFFFFEu, kSunLight = 0x7FFFFFFDu;
enum { LightInvalid = 0, LightSky = 1, LightSun = 2, LightLocalTriangle = 3 };
constexpr float kRoughnessThreshold = 0.00001f, kTranslucencyThreshold = 0.001f;
constexpr float kDisneyMinPdf = 1e-5f, kDisneyMaxThroughput = 32.0f, kDisneyMinLobeProb = 0.05f;

struct Hit { int hit, x, y, z, face, id; float t; int steps; };

// ------------------------------------------------------------------------------------------------ DDA
template <bool kSmemOcc>
__device__ __noinline__ Hit ddaTrace(const GridView &g, const uint32_t *__restrict__ occ, f3 o, f3 d, float tmin, float tmax)
{
    Hit h;
    h.hit = 0; h.x = h.y = h.z = 0; h.face = 6; h.id = 0; h.t = kRayMax; h.steps = 0;
    const int W = g.W, H = g.H, D = g.D;
    int x = (int)floorf(o.x), y = (int)floorf(o.y), z = (int)floorf(o.z);
    int hitAxis = -1;
    float tCur = 0.0f;

    if (x < 0 || x >= W || y < 0 || y >= H || z < 0 || z >= D)
    {
        const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
        const float dim[3] = {(float)W, (float)H, (float)D};
        float tEnter = -FLT_MAX, tExit = FLT_MAX;
        int axis = -1;
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            if (fabsf(dd[a]) < 1e-8f)
            {
                if (oo[a] < 0.0f || oo[a] >= dim[a]) return h;
                continue;
            }
            float ta = ex::divf(ex::subf(0.0f, oo[a]), dd[a]), tb = ex::divf(ex::subf(dim[a], oo[a]), dd[a]);
            float tn = fminr(ta, tb), tf = fmaxr(ta, tb);
            if (tn > tEnter) { tEnter = tn; axis = a; }
            if (tf < tExit) tExit = tf;
        }
        if (axis < 0 || tEnter > tExit || tExit < 0.0f || tEnter < 0.0f) return h;
        f3 p = ex::pointAt(o, d, tEnter);
        x = clampi((int)floorf(p.x), 0, W - 1);
        y = clampi((int)floorf(p.y), 0, H - 1);
        z = clampi((int)floorf(p.z), 0, D - 1);
        if (axis == 0) x = dd[0] > 0.0f ? 0 : W - 1;
        if (axis == 1) y = dd[1] > 0.0f ? 0 : H - 1;
        if (axis == 2) z = dd[2] > 0.0f ? 0 : D - 1;
        hitAxis = axis;
        tCur = tEnter;
    }

    const int stepX = (d.x > 0.0f) ? 1 : -1, stepY = (d.y > 0.0f) ? 1 : -1, stepZ = (d.z > 0.0f) ? 1 : -1;
    const float tDeltaX = (fabsf(d.x) < 1e-8f) ? FLT_MAX : ex::divf(1.0f, fabsf(d.x));
    const float tDeltaY = (fabsf(d.y) < 1e-8f) ? FLT_MAX : ex::divf(1.0f, fabsf(d.y));
    const float tDeltaZ = (fabsf(d.z) < 1e-8f) ? FLT_MAX : ex::divf(1.0f, fabsf(d.z));
    const float nbX = (stepX > 0) ? (float)(x + 1) : (float)x;
    const float nbY = (stepY > 0) ? (float)(y + 1) : (float)y;
    const float nbZ = (stepZ > 0) ? (float)(z + 1) : (float)z;
    float tMaxX = (fabsf(d.x) < 1e-8f) ? FLT_MAX : ex::divf(ex::subf(nbX, o.x), d.x);
    float tMaxY = (fabsf(d.y) < 1e-8f) ? FLT_MAX : ex::divf(ex::subf(nbY, o.y), d.y);
    float tMaxZ = (fabsf(d.z) < 1e-8f) ? FLT_MAX : ex::divf(ex::subf(nbZ, o.z), d.z);

    // Walk. Same decisions as the reference loop (bounds test, solid test, tMaxX<tMaxY / <tMaxZ selection with its
    // tie order, tMax += tDelta) in a branch-light form:
    //  * the cell is a linear voxel index: word = lin>>5, bit = lin&31 (W is a multiple of 32);
    //  * "if X<Y then (X<Z ? X : Z) else (Y<Z ? Y : Z)" == "A = X<Y ? X : Y; A<Z ? A : Z" (two compares);
    //  * the bounds test is a packed count of steps left inside the grid per axis (11-bit fields + guard bits:
    //    a borrow out of a field clears its guard), so leaving the grid costs one subtract and one test;
    //  * only the selected axis' tMax is advanced, by an exact __fadd_rn.
    int lin = (y * D + z) * W + x;
    const int dLinX = stepX, dLinY = stepY * W * D, dLinZ = stepZ * W;
    const uint32_t remX = (uint32_t)(stepX > 0 ? W - 1 - x : x), remY = (uint32_t)(stepY > 0 ? H - 1 - y : y),
                   remZ = (uint32_t)(stepZ > 0 ? D - 1 - z : z);
    // fields: X bits 0-9 (guard 10), Z bits 11-20 (guard 21), Y bits 22-30 (guard 31): W,D <= 1024, H <= 512 (checked by vpt_set_grid)
    constexpr uint32_t kGuards = (1u << 10) | (1u << 21) | (1u << 31);
    constexpr uint32_t kDecX = 1u, kDecZ = 1u << 11, kDecY = 1u << 22;
    const uint32_t rem0 = kGuards | remX | (remZ << 11) | (remY << 22);
    uint32_t rem = rem0, lastDec = 0;
    uint32_t occShared = 0;
    if (kSmemOcc) occShared = (uint32_t)__cvta_generic_to_shared(occ);
    bool found = false, left = false;
    for (;;)
    {
        if (tCur >= tmax) break;
        uint32_t word;
        if (kSmemOcc) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(occShared + (((uint32_t)lin >> 3) & ~3u)));
        else word = __ldg(occ + (lin >> 5));
        if (((word >> (lin & 31)) & 1u) && tCur >= tmin) { found = true; break; }
        const bool xy = tMaxX < tMaxY;
        const float tA = xy ? tMaxX : tMaxY;
        const bool az = tA < tMaxZ;
        const uint32_t dec = az ? (xy ? kDecX : kDecY) : kDecZ;
        rem -= dec;
        lastDec = dec;
        if ((rem & kGuards) != kGuards) { left = true; break; } // this step leaves the grid
        tCur = az ? tA : tMaxZ;
        if (az && xy) tMaxX = ex::addf(tMaxX, tDeltaX);
        if (az && !xy) tMaxY = ex::addf(tMaxY, tDeltaY);
        if (!az) tMaxZ = ex::addf(tMaxZ, tDeltaZ);
        lin += az ? (xy ? dLinX : dLinY) : dLinZ;
    }
    if (lastDec == kDecX) hitAxis = 0; else if (lastDec == kDecY) hitAxis = 1; else if (lastDec == kDecZ) hitAxis = 2;
    {
        // steps walked (statistics): initial minus remaining counts; the step that left the grid counts too
        const uint32_t r = left ? rem + lastDec : rem;
        h.steps = (int)((remX - (r & 1023u)) + (remZ - ((r >> 11) & 1023u)) + (remY - ((r >> 22) & 511u))) + (left ? 1 : 0);
    }
    if (found)
    {
        h.hit = 1; h.t = tCur;
        h.x = lin % W; const int yz = lin / W; h.z = yz % D; h.y = yz / D;
        h.id = __ldg(g.idsLinear + lin);
        if (hitAxis == 0) h.face = stepX > 0 ? 2 : 3;
        else if (hitAxis == 1) h.face = stepY > 0 ? 1 : 0;
        else if (hitAxis == 2) h.face = stepZ > 0 ? 5 : 4;
        else h.face = 6;
    }
    return h;
}

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
VPT_DEV f3 hitPoint(const Hit &h, f3 o, f3 d)
{
    f3 p = ex::pointAt(o, d, h.t);
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

__device__ __noinline__ void disneySample(f4 u, f3 n, f3 ng, f3 wo, f3 albedo, bool metallic, float translucency, float roughness,
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

__device__ __noinline__ void disneyEvaluate(f3 n, f3 ng, f3 wi, f3 wo, f3 albedo, bool metallic, float roughness, f3 &bsdf, float &pdf)
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

struct RayData
{
    f3 pos; float distance; f3 wo, wi; unsigned depth; f3 radiance, bsdfOverPdf; float pdf;
    bool hitFirstDiffuseSurface, shouldTerminate, isCurrentBounceDiffuse;
};

template <bool kSmemOcc>
struct Tracer
{
    const TraceArgs &a;
    const uint32_t *occ;
    int px, py, sampleIndex, prevSampleIndex, randIdx;
    unsigned rays, steps;

    VPT_DEV float blueNoise(int sIdx, int dim) const
    {
        // BlueNoiseRandGenerator::rand (RandGen.h:21-45); ranking table is padded with 256 zero bytes
        const int pi = px & 127, pj = py & 127;
        sIdx &= 255;
        const int base = (pi + pj * 128) * 8;
        const int ranked = sIdx ^ __ldg(a.ranking + dim + base);
        int value = __ldg(a.sobol + dim + ranked * 256);
        value ^= __ldg(a.scrambling + (dim % 8) + base);
        return value / 256.0f;
    }
    VPT_DEV float rnd() { return blueNoise(sampleIndex, randIdx++); }
    VPT_DEV f2 rnd2() { float x = rnd(); float y = rnd(); return {x, y}; }
    VPT_DEV f4 rnd4() { float x = rnd(); float y = rnd(); float z = rnd(); float w = rnd(); return {x, y, z, w}; }
    VPT_DEV float rnd16() { f2 u = rnd2(); return u.x + u.y / 256.0f; }
    VPT_DEV Hit trace(f3 o, f3 d, float tmin, float tmax)
    {
        Hit h = ddaTrace<kSmemOcc>(a.grid, occ, o, d, tmin, tmax);
        ++rays; steps += (unsigned)h.steps;
        return h;
    }
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
    __device__ __noinline__ LightSample createSunLightSample(int idx) const
    {
        int ix = idx % a.sunW, iy = idx / a.sunW;
        f2 uv = {(ix + 0.5f) / float(a.sunW), (iy + 0.5f) / float(a.sunH)};
        LightSample ls;
        ls.solidAnglePdf = (a.sunW * a.sunH) / (kTwoPi * (1.0f - a.sunCosThetaMax));
        ls.position = equalAreaMapCone(sunDir(), uv.x, uv.y, a.sunCosThetaMax);
        ls.radiance = xyz(loadSun(ix, iy));
        ls.lightType = LightSun;
        return ls;
    }
    __device__ __noinline__ LightSample createSkyLightSample(int idx) const
    {
        int ix = idx % a.skyW, iy = idx / a.skyW;
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
    __device__ __noinline__ float evalCandidate(const Surface &s, const LightSample &ls, float lightSelectionPdf, float lightMisWeight,
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
    __device__ __noinline__ bool lightSampleFromReservoir(LightSample &ls, const VptReservoir &r) const
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
        return li < kInvalidLight;
    }
    __device__ __noinline__ bool getPrevSurface(Surface &s, int x, int y) const
    {
        const VptCamera &pc = a.prevCam;
        if (x < 0 || y < 0 || x >= pc.resolution[0] || y >= pc.resolution[1]) return false;
        const size_t i = (size_t)y * a.width + x;
        s.depth = __ldg(a.prev.depth + i);
        if (s.depth == kRayMax) return false;
        const float4 nr = __ldg(a.prev.normalRoughness + i), gt = __ldg(a.prev.geoNormalThinfilm + i), mp = __ldg(a.prev.materialParameter + i);
        const float j0 = blueNoise(prevSampleIndex, 0), j1 = blueNoise(prevSampleIndex, 1);
        f2 prevUV = {(float(x) + j0) * pc.inversedResolution[0], (float(y) + j1) * pc.inversedResolution[1]};
        f3 viewDir = uvToWorldDirection(pc, prevUV);
        s.pos = F3(pc.pos[0], pc.pos[1], pc.pos[2]) + viewDir * s.depth;
        s.wo = -viewDir;
        s.normal = xyz(nr);
        s.geoNormal = xyz(gt);
        s.albedo = xyz(__ldg(a.prev.albedo + i));
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

// ------------------------------------------------------------------------------------------------ miss
template <bool kSmemOcc>
__device__ __noinline__ void missRadiance(Tracer<kSmemOcc> &c, RayData &rd, bool ownsGBuffer)
{
    const TraceArgs &a = c.a;
    const size_t pix = (size_t)c.py * a.width + c.px;
    if (rd.depth == 0 && ownsGBuffer)
    {
        storeReservoir(a.resCur + pix, emptyReservoir());
        a.cur.albedo[pix] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        a.cur.material[pix] = (float)0xFFFF;
        a.cur.normalRoughness[pix] = make_float4(0.0f, -1.0f, 0.0f, 0.0f);
        a.cur.geoNormalThinfilm[pix] = make_float4(0.0f, -1.0f, 0.0f, 0.0f);
        a.cur.materialParameter[pix] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    f3 emission = F3(0.0f);
    const f3 rayDir = rd.wi;
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
    rd.radiance = emission;
    rd.distance = kRayMax;
    rd.shouldTerminate = true;
}

// ------------------------------------------------------------------------------------------------ closest hit
template <bool kSmemOcc>
__device__ __noinline__ void closestHit(Tracer<kSmemOcc> &c, RayData &rd, const Hit &h, f3 rayOrig, bool ownsGBuffer)
{
    const TraceArgs &a = c.a;
    const size_t pix = (size_t)c.py * a.width + c.px;
    const bool gbufferPass = ownsGBuffer && rd.depth == 0;

    rd.distance = h.t;
    const f3 geoNormal = faceNormal(h.face, rd.wi);
    const f3 surfPos = hitPoint(h, rayOrig, rd.wi);
    const f3 frontPos = surfPos + geoNormal * kSpawnEps;
    rd.pos = frontPos;
    const VptMaterial *mat = a.materials + __ldg(a.blockToMaterial + h.id);
    const int isEmissive = __ldg(&mat->isEmissive);
    const f3 matAlbedo = {__ldg(&mat->albedo[0]), __ldg(&mat->albedo[1]), __ldg(&mat->albedo[2])};

    if (isEmissive)
    {
        if (!rd.hitFirstDiffuseSurface)
        {
            rd.radiance = matAlbedo;
            if (ownsGBuffer)
            {
                a.cur.albedo[pix] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
                a.cur.material[pix] = (float)0xFFFF;
                a.cur.normalRoughness[pix] = make_float4(0.0f, -1.0f, 0.0f, 0.0f);
                a.cur.geoNormalThinfilm[pix] = make_float4(0.0f, -1.0f, 0.0f, 0.0f);
                a.cur.materialParameter[pix] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        rd.shouldTerminate = true;
        return;
    }

    Surface s;
    s.geoNormal = geoNormal;
    s.wo = rd.wo;
    s.albedo = max3f(matAlbedo, F3(0.001f));
    s.roughness = __ldg(&mat->roughness);
    if (rd.hitFirstDiffuseSurface) s.roughness = fminr(s.roughness * 2.0f + 0.1f, 1.0f);
    const bool isDiffuse = s.roughness > kRoughnessThreshold;
    s.metallic = __ldg(&mat->metallic) != 0;
    s.translucency = __ldg(&mat->translucency);
    s.normal = lerp3(geoNormal, geoNormal, 0.2f);
    rd.isCurrentBounceDiffuse = isDiffuse;
    const int materialId = __ldg(&mat->materialId);

    if (gbufferPass)
    {
        a.cur.material[pix] = (float)materialId;
        a.cur.normalRoughness[pix] = make_float4(s.normal.x, s.normal.y, s.normal.z, s.roughness);
        a.cur.geoNormalThinfilm[pix] = make_float4(s.normal.x, s.normal.y, s.normal.z, 0.0f);
        a.cur.materialParameter[pix] = make_float4(s.metallic ? 1.0f : 0.0f, s.translucency, 0.0f, 0.0f);
    }

    f3 bsdfWi, bsdfOverPdf; float bsdfPdf; bool transmission = false;
    disneySample(c.rnd4(), s.normal, s.geoNormal, s.wo, s.albedo, s.metallic, s.translucency, s.roughness, bsdfWi, bsdfOverPdf, bsdfPdf, transmission);
    if (bsdfPdf <= 0.0f) rd.shouldTerminate = true;
    rd.pos = frontPos;
    rd.wi = bsdfWi;
    rd.bsdfOverPdf = bsdfOverPdf;
    rd.pdf = bsdfPdf;

    bool skipAlbedoInShadowRay = false;
    if (rd.depth == 0)
    {
        rd.hitFirstDiffuseSurface = true;
        if (ownsGBuffer) a.cur.albedo[pix] = make_float4(s.albedo.x, s.albedo.y, s.albedo.z, 1.0f);
        skipAlbedoInShadowRay = true;
    }
    const bool enableReSTIR = a.enableRestir && rd.depth == 0 && ownsGBuffer;
    VptReservoir *storeSlot = (rd.depth == 0 && ownsGBuffer) ? a.resCur + pix : nullptr;

    if (!isDiffuse)
    {
        if (storeSlot) storeReservoir(storeSlot, emptyReservoir());
        return;
    }
    s.pos = rd.pos;
    s.depth = rd.distance;

    const f3 sunD = c.sunDir();
    LightSample lightSample = noLight();
    VptReservoir ris = emptyReservoir();
    const bool skipSun = (dot(s.normal, sunD) < 0.0f || dot(s.geoNormal, sunD) < 0.0f);
    const int nLocal = 0;
    const int nSun = skipSun ? 0 : 1, nSky = 1, nBrdf = 1;
    const int nMis = nLocal + nSun + nSky + nBrdf;
    const float localMisW = float(nLocal) / nMis, sunMisW = float(nSun) / nMis, skyMisW = float(nSky) / nMis, brdfMisW = float(nBrdf) / nMis;

    VptReservoir localRes = emptyReservoir();
    LightSample localSample = noLight();
    finalizeResampling(localRes, 1.0f, (float)nMis);
    localRes.M = 1;

    VptReservoir sunRes = emptyReservoir();
    LightSample sunSample = noLight();
    for (int i = 0; i < nSun; ++i)
    {
        float sourcePdf;
        int idx = (int)c.aliasSample(a.sunAlias, a.sunW * a.sunH, c.rnd(), sourcePdf);
        LightSample cand = c.createSunLightSample(idx);
        int ix = idx % a.sunW, iy = idx / a.sunW;
        f2 uv = {(ix + 0.5f) / float(a.sunW), (iy + 0.5f) / float(a.sunH)};
        float blended;
        float targetPdf = c.evalCandidate(s, cand, sourcePdf, sunMisW, brdfMisW, &blended);
        float risRnd = c.rnd();
        if (streamSample(sunRes, kSunLight, uv, risRnd, targetPdf, 1.0f / blended)) sunSample = cand;
    }
    finalizeResampling(sunRes, 1.0f, (float)nMis);
    sunRes.M = 1;

    VptReservoir skyRes = emptyReservoir();
    LightSample skySample = noLight();
    for (int i = 0; i < nSky; ++i)
    {
        float sourcePdf;
        int idx = (int)c.aliasSample(a.skyAlias, a.skyW * a.skyH, c.rnd16(), sourcePdf);
        LightSample cand = c.createSkyLightSample(idx);
        int ix = idx % a.skyW, iy = idx / a.skyW;
        f2 uv = {(ix + 0.5f) / float(a.skyW), (iy + 0.5f) / float(a.skyH)};
        float blended;
        float targetPdf = c.evalCandidate(s, cand, sourcePdf, skyMisW, brdfMisW, &blended);
        float risRnd = c.rnd();
        if (streamSample(skyRes, kSkyLight, uv, risRnd, targetPdf, 1.0f / blended)) skySample = cand;
    }
    finalizeResampling(skyRes, 1.0f, (float)nMis);
    skyRes.M = 1;

    // Shadow rays from frontPos over an unbounded range depend only on the direction: the reference re-traces the
    // BSDF-sampled direction when RIS selects it (closesthit.cu:616) and re-traces the initial sample for the final
    // visibility when temporal resampling keeps it (:801). Identical (origin, direction) -> identical result, so the
    // result is reused instead of walking the grid again (rays are counted only when actually traced).
    f3 visDir0 = F3(0.0f), visDir1 = F3(0.0f);
    bool visRes0 = false, visRes1 = false, visHave0 = false, visHave1 = false;
    auto sameDir = [](f3 a, f3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; };

    VptReservoir brdfRes = emptyReservoir();
    LightSample brdfSample = noLight();
    for (int i = 0; i < nBrdf; ++i)
    {
        float lightSourcePdf = 0.0f;
        f3 sampleDir;
        uint32_t lightIndex = kInvalidLight;
        f2 uv = {0, 0};
        LightSample cand = noLight();
        float brdfPdf; bool trans = false; f3 dummy;
        disneySample(c.rnd4(), s.normal, s.geoNormal, s.wo, s.albedo, s.metallic, s.translucency, s.roughness, sampleDir, dummy, brdfPdf, trans);
        if (brdfPdf > 0.0f)
        {
            Hit sh = c.trace(frontPos, sampleDir, 0.0f, FLT_MAX);
            visDir0 = sampleDir; visRes0 = !sh.hit; visHave0 = true;
            if (!sh.hit)
            {
                if (equalAreaMapConeInv(uv, sunD, sampleDir, a.sunCosThetaMax))
                {
                    lightIndex = kSunLight;
                    int sx = (int)(uv.x * a.sunW - 0.5f), sy = (int)(uv.y * a.sunH - 0.5f);
                    if (sx >= a.sunW) sx %= a.sunW;
                    if (sx < 0) sx = a.sunW - ((-sx) % a.sunW);
                    sy = clampi(sy, 0, a.sunH - 1);
                    int idx = sy * a.sunW + sx;
                    cand = c.createSunLightSample(idx);
                    cand.position = sampleDir;
                    lightSourcePdf = __ldg(&a.sunAlias[idx].p);
                }
                else
                {
                    lightIndex = kSkyLight;
                    uv = equalAreaSphereMapInv(sampleDir);
                    int kx = (int)(uv.x * a.skyW - 0.5f), ky = (int)(uv.y * a.skyH - 0.5f);
                    int idx = ky * a.skyW + kx;
                    idx = clampi(idx, 0, a.skyW * a.skyH - 1);
                    cand = c.createSkyLightSample(idx);
                    cand.position = sampleDir;
                    lightSourcePdf = __ldg(&a.skyAlias[idx].p);
                }
            }
        }
        if (lightSourcePdf == 0.0f) continue;
        float misW = (lightIndex == kSkyLight) ? skyMisW : ((lightIndex == kSunLight) ? sunMisW : localMisW);
        float blended;
        float targetPdf = c.evalCandidate(s, cand, lightSourcePdf, misW, brdfMisW, &blended);
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
        if (visHave0 && sameDir(visDir0, lightSample.position)) isLightVisible = visRes0;
        else
        {
            Hit vh = c.trace(frontPos, lightSample.position, 0.0f, kRayMax);
            isLightVisible = !vh.hit;
        }
        visDir1 = lightSample.position; visRes1 = isLightVisible; visHave1 = true;
        if (!isLightVisible) { ris.lightData = 0; ris.weightSum = 0; }
    }

    VptReservoir restir = emptyReservoir();
    if (enableReSTIR)
    {
        const VptCamera &pc = a.prevCam;
        const f3 pcPos = {pc.pos[0], pc.pos[1], pc.pos[2]};
        combineReservoirs(restir, ris, 0.5f, ris.targetPdf);
        const f3 prevWorldPos = s.pos;
        f2 prevUV = c.worldDirectionToUV(pc, normalize(prevWorldPos - pcPos));
        const int prevPx = (int)(prevUV.x * pc.resolution[0]), prevPy = (int)(prevUV.y * pc.resolution[1]);
        const float expectedPrevDepth = distance(prevWorldPos, pcPos);
        constexpr int nTemporal = 3;
        constexpr float mCap = 20.0f;
        int offx[nTemporal], offy[nTemporal];
        offx[0] = prevPx - c.px; offy[0] = prevPy - c.py;
        {
            f2 dsk = concentricSampleDisk(c.rnd2()) * 64.0f;
            offx[1] = prevPx - c.px + (int)dsk.x; offy[1] = prevPy - c.py + (int)dsk.y;
        }
        {
            f2 dsk = concentricSampleDisk(c.rnd2()) * 64.0f;
            offx[2] = (int)dsk.x; offy[2] = (int)dsk.y;
        }
        unsigned cached = 0;
        int selectedLoopIdx = -1;
#pragma unroll 1
        for (int i = 0; i < nTemporal; ++i)
        {
            int ix = c.px + offx[i], iy = c.py + offy[i];
            clampIntoView(ix, iy, a.width, a.height);
            Surface ts;
            if (!c.getPrevSurface(ts, ix, iy)) continue;
            bool nOk = dot(s.normal, ts.geoNormal) >= 0.5f;
            bool dOk = fabsf(expectedPrevDepth - ts.depth) <= 0.1f * fmaxr(expectedPrevDepth, ts.depth);
            bool rOk = fabsf(s.roughness - ts.roughness) <= 0.5f * fmaxr(s.roughness, ts.roughness);
            if (!(nOk && dOk && rOk)) continue;
            cached |= (1u << i);
            VptReservoir pr = loadReservoir(a.resPrev + (size_t)iy * a.width + ix);
            if (isnan(pr.weightSum) || isinf(pr.weightSum)) pr = emptyReservoir();
            if (pr.M > mCap) pr.M = mCap;
            float neighborWeight = 0;
            LightSample cand = noLight();
            if (isValidReservoir(pr))
            {
                if (!c.lightSampleFromReservoir(cand, pr)) pr = emptyReservoir();
                neighborWeight = c.targetPdfForSurface(cand, s);
            }
            if (combineReservoirs(restir, pr, c.rnd(), neighborWeight)) { lightSample = cand; selectedLoopIdx = i; }
        }
        if (isValidReservoir(restir))
        {
            float pi = restir.targetPdf, piSum = restir.targetPdf * 1;
#pragma unroll 1
            for (int i = 0; i < nTemporal; ++i)
            {
                if ((cached & (1u << i)) == 0) continue;
                int ix = c.px + offx[i], iy = c.py + offy[i];
                clampIntoView(ix, iy, a.width, a.height);
                Surface ts;
                c.getPrevSurface(ts, ix, iy);
                LightSample atNeighbor = noLight();
                c.lightSampleFromReservoir(atNeighbor, restir);
                float ps = c.targetPdfForSurface(atNeighbor, ts);
                if (ps > 0 && !(i == 0 && i == selectedLoopIdx))
                {
                    const float extraRayOffset = 0.01f + 0.01f * ts.depth;
                    Hit nh = c.trace(ts.pos, lightSample.position, extraRayOffset, kRayMax);
                    if (nh.hit) ps = 0.0f;
                }
                VptReservoir pr = loadReservoir(a.resPrev + (size_t)iy * a.width + ix);
                if (isnan(pr.weightSum) || isinf(pr.weightSum)) pr = emptyReservoir();
                if (pr.M > mCap) pr.M = mCap;
                if (selectedLoopIdx == i) pi = ps;
                piSum += ps * pr.M;
            }
            finalizeResampling(restir, pi, piSum);
        }
        if (lightSample.lightType != LightInvalid)
        {
            if (visHave1 && sameDir(visDir1, lightSample.position)) isLightVisible = visRes1;
            else if (visHave0 && sameDir(visDir0, lightSample.position)) isLightVisible = visRes0;
            else
            {
                Hit vh = c.trace(frontPos, lightSample.position, 0.0f, kRayMax);
                isLightVisible = !vh.hit;
            }
            if (!isLightVisible) { restir.lightData = 0; restir.weightSum = 0; }
        }
    }

    const VptReservoir shading = enableReSTIR ? restir : ris;
    if (lightSample.lightType != LightInvalid && isValidReservoir(shading) && isLightVisible)
    {
        f3 sampleDir = lightSample.position;
        const f3 albedo = skipAlbedoInShadowRay ? F3(1.0f) : s.albedo;
        f3 bsdf; float pdf;
        disneyEvaluate(s.normal, s.geoNormal, sampleDir, s.wo, albedo, s.metallic, s.roughness, bsdf, pdf);
        float cosTheta = fmaxf(0.0f, dot(sampleDir, s.normal));
        f3 shadowRad = bsdf * cosTheta * lightSample.radiance * shading.weightSum / lightSample.solidAnglePdf;
        rd.radiance += shadowRad;
    }
    if (storeSlot) storeReservoir(storeSlot, enableReSTIR ? restir : emptyReservoir());
}

// ------------------------------------------------------------------------------------------------ raygen
template <bool kSmemOcc>
VPT_DEV f3 tracePath(Tracer<kSmemOcc> &c, bool ownsGBuffer, float &primaryDist)
{
    const TraceArgs &a = c.a;
    RayData rd;
    c.randIdx = 0;
    f2 jitter = c.rnd2();
    // exact arithmetic up to the primary hit (bit-exact voxel/face vs the oracle)
    f2 sampleUv = {ex::mulf(ex::addf(float(c.px), jitter.x), a.cam.inversedResolution[0]), ex::mulf(ex::addf(float(c.py), jitter.y), a.cam.inversedResolution[1])};
    rd.pos = F3(a.cam.pos[0], a.cam.pos[1], a.cam.pos[2]);
    rd.wi = ex::normalize(ex::mulMat3(a.cam.uvToWorld, F3(sampleUv.x, sampleUv.y, 1.0f)));
    f3 radiance = F3(0.0f), throughput = F3(1.0f);
    rd.depth = 0;
    rd.isCurrentBounceDiffuse = false;
    rd.hitFirstDiffuseSurface = false;
    primaryDist = kRayMax;
    bool terminated = false;
    int totalBounce = 0, diffuseBounce = 0;
    while (!terminated)
    {
        rd.bsdfOverPdf = F3(1.0f); rd.pdf = 0.0f; rd.radiance = F3(0.0f); rd.wo = -rd.wi; rd.distance = kRayMax;
        rd.shouldTerminate = false;
        rd.isCurrentBounceDiffuse = false;
        const f3 orig = rd.pos;
        Hit h = c.trace(orig, rd.wi, 0.0f, kRayMax);
        if (rd.depth == 0 && ownsGBuffer)
            a.primaryHits[(size_t)c.py * a.width + c.px] = h.hit ? make_int4(h.x, h.y, h.z, h.face) : make_int4(-1, -1, -1, -1);
        if (h.hit) closestHit(c, rd, h, orig, ownsGBuffer);
        else missRadiance(c, rd, ownsGBuffer);
        radiance += throughput * rd.radiance;
        bool cont = !(rd.shouldTerminate || rd.pdf <= 0.0f || isNull(rd.bsdfOverPdf));
        if (cont) throughput *= rd.bsdfOverPdf;
        terminated = !cont;
        ++totalBounce;
        if (rd.isCurrentBounceDiffuse) ++diffuseBounce;
        if (totalBounce == a.totalBounceLimit || diffuseBounce == a.diffuseBounceLimit) terminated = true;
        if (rd.depth == 0) primaryDist = rd.distance;
        ++rd.depth;
    }
    if (isnan(radiance.x) || isnan(radiance.y) || isnan(radiance.z)) radiance = F3(0.5f);
    return radiance;
}

#ifndef VPT_TRACE_THREADS
#define VPT_TRACE_THREADS 512
#endif
constexpr int kTraceThreads = VPT_TRACE_THREADS;
#ifndef VPT_TRACE_CTAS_PER_SM
#define VPT_TRACE_CTAS_PER_SM 2
#endif

template <bool kSmemOcc>
__global__ void __launch_bounds__(kTraceThreads, VPT_TRACE_CTAS_PER_SM) traceKernel(const __grid_constant__ TraceArgs a)
{
    extern __shared__ uint32_t occS[];
    const uint32_t *occ = a.grid.occ;
    if (kSmemOcc)
    {
        // stage the whole occupancy mask: 16-byte vector loads, coalesced
        const uint4 *src = reinterpret_cast<const uint4 *>(a.grid.occ);
        uint4 *dst = reinterpret_cast<uint4 *>(occS);
        const int n4 = a.grid.occWords >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
        for (int i = (n4 << 2) + threadIdx.x; i < a.grid.occWords; i += blockDim.x) occS[i] = __ldg(a.grid.occ + i);
        __syncthreads();
        occ = occS;
    }
    const int lane = threadIdx.x & 31;
    const int tilesX = (a.width + 7) >> 3, tilesY = (a.height + 3) >> 2;
    const int nTiles = tilesX * tilesY;
    unsigned long long *sched = a.counters + 2;
    unsigned raysAcc = 0, stepsAcc = 0;
    for (;;)
    {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(sched, 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= nTiles) break;
        const int px = (tile % tilesX) * 8 + (lane & 7);
        const int py = (tile / tilesX) * 4 + (lane >> 3);
        if (px < a.width && py < a.height)
        {
            Tracer<kSmemOcc> c{a, occ, px, py, 0, 0, 0, 0u, 0u};
            f3 sum = F3(0.0f);
            float depth0 = kRayMax;
            bool haveDepth = false;
            for (int k = a.sampleBegin; k < a.spp; k += a.sampleStep)
            {
                c.sampleIndex = a.iterationIndex * a.spp + k;
                c.prevSampleIndex = (a.iterationIndex - 1) * a.spp;
                float pd;
                f3 r = tracePath(c, k == 0, pd);
                sum += r;
                if (k == 0) { depth0 = pd; haveDepth = true; }
            }
            const size_t pix = (size_t)py * a.width + px;
            if (haveDepth) a.cur.depth[pix] = depth0;
            a.illumination[pix] = make_float4(sum.x, sum.y, sum.z, haveDepth ? depth0 : 0.0f);
            raysAcc += c.rays; stepsAcc += c.steps;
        }
    }
    // statistics: one atomic pair per warp
    unsigned long long r64 = raysAcc, s64 = stepsAcc;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        r64 += __shfl_down_sync(0xffffffffu, r64, off);
        s64 += __shfl_down_sync(0xffffffffu, s64, off);
    }
    if (lane == 0) { atomicAdd(a.counters + 0, r64); atomicAdd(a.counters + 1, s64); }
}

__global__ void resolveKernel(float4 *illum, int npix, float spp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 v = illum[i];
    v.x = v.x / spp; v.y = v.y / spp; v.z = v.z / spp;
    illum[i] = v;
}

cudaError_t launchTrace(const TraceArgs &a, cudaStream_t s, int smCount, size_t smemOptIn)
{
    const size_t occBytes = (size_t)a.grid.occWords * 4;
    const bool smem = a.occInSmem != 0;
    const int grid = smCount * VPT_TRACE_CTAS_PER_SM;
    if (smem)
    {
        if (occBytes > smemOptIn) return cudaErrorInvalidValue;
        cudaError_t e = cudaFuncSetAttribute(traceKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)occBytes);
        if (e != cudaSuccess) return e;
        traceKernel<true><<<grid, kTraceThreads, occBytes, s>>>(a);
    }
    else
        traceKernel<false><<<grid, kTraceThreads, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launchResolve(float4 *illum, int npix, float spp, cudaStream_t s)
{
    resolveKernel<<<(npix + 255) / 256, 256, 0, s>>>(illum, npix, spp);
    return cudaGetLastError();
}

} // namespace vpt
