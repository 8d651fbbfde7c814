// Shared device helpers of the denoiser translation units (vpt_denoise.cu: fast arithmetic class; vpt_temporal.cu:
// exact class). See vpt_denoise.cu for the pass map.
#pragma once
#include "vpt_kernels.h"
#include "vpt_math.cuh"

namespace vpt {


constexpr float kDenoisingRange = 500000.0f;
constexpr float kSkyZs = 1.0e20f; // zs above this = sky (depth is kRayMax = 1e27 there)
#ifndef VPT_DN_BY
#define VPT_DN_BY 8
#endif
#ifndef VPT_ATROUS_BY
#define VPT_ATROUS_BY 8
#endif
constexpr int kBX = 32, kBY = VPT_DN_BY, kAtrousBY = VPT_ATROUS_BY; // (32x32 a-trous CTAs were measured slower: 71 vs 61 us)

// a*a - b*b style cancellations: no FMA contraction (the oracle is built -ffp-contract=off)
VPT_DEV float subSq(float a, float b) { return __fsub_rn(a, __fmul_rn(b, b)); }
VPT_DEV float diffSq(float a, float b) { return __fsub_rn(__fmul_rn(a, a), __fmul_rn(b, b)); }

struct Cam
{
    f3 pos, dir; float invResX, invResY, tanHalfFovX, resX;
    mat3 uvToWorld, worldToUv;
};
VPT_DEV Cam loadCam(const VptCamera &c)
{
    Cam k;
    k.pos = F3(c.pos[0], c.pos[1], c.pos[2]); k.dir = F3(c.dir[0], c.dir[1], c.dir[2]);
    k.invResX = c.inversedResolution[0]; k.invResY = c.inversedResolution[1];
    k.tanHalfFovX = c.tanHalfFov[0]; k.resX = c.resolution[0];
    k.uvToWorld = mat3From(c.uvToWorld); k.worldToUv = mat3From(c.worldToUv);
    return k;
}
VPT_DEV f3 uvToWorldDirection(const Cam &c, f2 uv) { return normalize(mul(c.uvToWorld, F3(uv.x, uv.y, 1.0f))); }
VPT_DEV f2 worldDirectionToUV(const Cam &c, f3 d) { f3 h = mul(c.worldToUv, d); return {h.x / h.z, h.y / h.z}; }
VPT_DEV f3 worldPosFromPixel(const Cam &c, int x, int y, float depth)
{
    f2 uv = {(float(x) + 0.5f) * c.invResX, (float(y) + 0.5f) * c.invResY};
    return c.pos + uvToWorldDirection(c, uv) * depth;
}
// view vector of pixel (x,y): M*(u,v,1) = M0 + x*Mx + y*My (DnView is filled on the host from the camera)
VPT_DEV f3 viewVec(const DnView &v, float x, float y)
{
    return {fmaf(y, v.My[0], fmaf(x, v.Mx[0], v.M0[0])), fmaf(y, v.My[1], fmaf(x, v.Mx[1], v.M0[1])), fmaf(y, v.My[2], fmaf(x, v.Mx[2], v.M0[2]))};
}
VPT_DEV f4 ld4(const float4 *b, int W, int H, int x, int y)
{
    x = clampi(x, 0, W - 1); y = clampi(y, 0, H - 1);
    return F4(__ldg(b + (size_t)y * W + x));
}
VPT_DEV float ld1(const float *b, int W, int H, int x, int y)
{
    x = clampi(x, 0, W - 1); y = clampi(y, 0, H - 1);
    return __ldg(b + (size_t)y * W + x);
}
// Load2DUshort1 on the float material surface (Sampler.h:102-107 used at HistoryFix.h:61,87; Atrous.h:47,110)
VPT_DEV float matU16(const float *mat, int W, int H, int x, int y)
{
    y = clampi(y, 0, H - 1);
    x = clampi(x, 0, 2 * W - 1);
    const unsigned short *row = reinterpret_cast<const unsigned short *>(mat + (size_t)y * W);
    return (float)__ldg(row + x);
}
VPT_DEV float linearStep(float a, float b, float x) { return saturate((x - a) / (b - a)); }
VPT_DEV float smoothStep(float a, float b, float x) { float t = linearStep(a, b, x); return t * t * (3.0f - 2.0f * t); }
VPT_DEV float acosApprox(float x) { return sqrtf(2.0f) * sqrtf(saturate(1.0f - x)); }
VPT_DEV float nonExpWeight(float x, float px, float py) { return smoothStep(1.0f, 0.0f, fabsf(x * px + py)); }
// nonExpWeight(acosApprox(d), p, 0) with the constants folded: t = 1 - min(sqrt(2*sat(1-d))*p, 1); t*t*(3-2t)
VPT_DEV float normalWeight(float d, float p)
{
    const float t = 1.0f - fminf(sqrtf(2.0f * saturate(1.0f - d)) * p, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
VPT_DEV float specLobeTanHalfAngle(float roughness, float percentOfVolume)
{
    roughness = saturate(roughness); percentOfVolume = saturate(percentOfVolume);
    return roughness * roughness * percentOfVolume / (1.0f - percentOfVolume + 1e-6f);
}
VPT_DEV float normalWeightParam2(float roughness, float angleFraction)
{
    float angle = atanf(specLobeTanHalfAngle(roughness, angleFraction));
    return 1.0f / fmaxr(angle, 1e-6f);
}
VPT_DEV float planeDistWeightAtrous(f3 cpos, f3 cn, f3 spos, float thr) { return fabsf(dot(spos - cpos, cn)) < thr ? 1.0f : 0.0f; }
VPT_DEV f3 rgbToYCoCg(f3 c) { return {0.25f * (c.x + 2.0f * c.y + c.z), c.x - c.z, c.y - 0.5f * (c.x + c.z)}; }
VPT_DEV f3 yCoCgToRgb(f3 c) { return {c.x + 0.5f * (c.y - c.z), c.x + 0.5f * c.z, c.x - 0.5f * (c.y + c.z)}; }
VPT_DEV uint32_t seqHash(uint32_t x) { x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16; return x; }
VPT_DEV uint32_t seqExplode(uint32_t x)
{
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// ---- argument blocks shared by the gather kernels (vpt_denoise.cu) and the shared-memory tile kernels (vpt_dn_tiles.cu)
struct AtrousArgs
{
    int W, H, rowBegin, rowEnd;
    DnView view;
    float phiLuminance, depthThreshold, lobeAngleFraction;
    unsigned frameIndex, step;
    float nParamFull; // GetNormalWeightParam2(1, lobe fraction) of a pixel with historyLength >= 5 (launch-uniform), from the host
    const float4 *in, *G;
    const uint32_t *MQ;
    const float *histLen;
    const float4 *albedo; // composite variant
    float4 *out;
};
// per-centre constants of the tangent-plane test: dist(tap) = | zs_tap * (A0 + x*Ax + y*Ay) - c0 |
struct PlaneTest { float A0, Ax, Ay, c0, thr; };
VPT_DEV PlaneTest planeTest(const DnView &v, int x, int y, f3 cn, float zs, float depthThreshold)
{
    const f3 vc = viewVec(v, (float)x, (float)y);
    PlaneTest p;
    p.A0 = dot(F3(v.M0[0], v.M0[1], v.M0[2]), cn); p.Ax = dot(F3(v.Mx[0], v.Mx[1], v.Mx[2]), cn); p.Ay = dot(F3(v.My[0], v.My[1], v.My[2]), cn);
    p.c0 = zs * dot(vc, cn);
    p.thr = depthThreshold * (zs * sqrtf(dot(vc, vc))); // depthThreshold * z
    return p;
}
VPT_DEV bool planeNear(const PlaneTest &p, float zsTap, float x, float y)
{
    return fabsf(fmaf(zsTap, fmaf(y, p.Ay, fmaf(x, p.Ax, p.A0)), -p.c0)) < p.thr;
}
#ifndef VPT_ATAP_BRANCH
#define VPT_ATAP_BRANCH 0 // skipping the normal weight where the tap's normal equals the centre's (weight exactly 1): measured slower both as a
                          // per-thread branch (three passes 142 vs 128 us at 1080p) and as a warp-uniform vote (128 vs 116 us)
#endif
// One edge-stopped tap of an a-trous pass (Atrous.h:76-140), shared by the gather and the tile kernel so that both data paths
// run the very same arithmetic. kern = the 3x3 kernel weight, ok = the tap is inside the image.
//  * the weighted sums are explicit FMAs (the f4 operators left 4 FMUL + 4 FADD per tap: 7 % of the pass, ncu r2e).
VPT_DEV void atrousTap(const PlaneTest &pt, f3 cn, uint32_t cMat, float nParam, float cLum, float phiInv, bool ok, float kern, float4 sg, uint32_t sm,
                       float4 sv, float fx, float fy, float &sumW, f4 &sum)
{
    float w = (ok && sg.w < kSkyZs && (sm & 0xffffu) == cMat && planeNear(pt, sg.w, fx, fy)) ? kern : 0.0f;
    const float d = dot(cn, F3(sg.x, sg.y, sg.z));
#if VPT_ATAP_BRANCH
    if (d < 1.0f) w *= normalWeight(d, nParam);
#else
    w *= normalWeight(d, nParam);
#endif
    if (w > 1e-4f)
    {
        const float lumW = fabsf(cLum - luminance(xyz(sv))) * phiInv;
        w *= __expf(-lumW);
        sumW += w;
        sum.x = fmaf(w, sv.x, sum.x); sum.y = fmaf(w, sv.y, sum.y); sum.z = fmaf(w, sv.z, sum.z); sum.w = fmaf(w * w, sv.w, sum.w);
    }
}

struct ClampArgs
{
    int W, H, rowBegin, rowEnd;
    const float4 *illum, *ping, *pong;
    const float *depth, *histLen;
    float4 *prevIllum, *prevFast;
    float *prevHistLen;
};
AtrousArgs makeAtrousArgs(const DenoiseLaunch &d, const float4 *in, float4 *out, unsigned frameIndex, unsigned step); // host (vpt_denoise.cu)

#define PIXEL_GUARD(W_, rowBegin_, rowEnd_)                       \
    const int x = blockIdx.x * kBX + threadIdx.x;                 \
    const int y = (rowBegin_) + blockIdx.y * kBY + threadIdx.y;   \
    if (x >= (W_) || y >= (rowEnd_)) return;                      \
    const size_t pix = (size_t)y * (W_) + x;


} // namespace vpt
