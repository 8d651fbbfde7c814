// TemporalAccumulation (/root/reference/renderer/denoising/TemporalAccumulation.h:228-449, 29-215; samplers
// shaders/Sampler.h:396-498, 576-698).
//
// Mixed arithmetic classes. This pass produces historyLength, which is a CONTROL variable of the rest of the chain —
// HistoryFix and HistoryClamping branch on hl <= 4, AtrousSmem on hl >= 3 — and in the first frames hl sits exactly on
// those integers (n frames of history -> n +- an ulp from the bilinear weight normalisation and the footprint quality).
// One ulp of difference flips the branch for half the image. So the chain that decides hl — view vector, world position,
// reprojected uv, N.V of both frames, the parallax and disocclusion thresholds of the tap-validity tests, the averaged and
// re-projected normals, the bilinear
// weights, the custom-weight history-length fetch and the footprint arithmetic — is the EXACT class (vpt::ex::, explicit round-to-nearest intrinsics, compensated dot / Mat3*v like the
// reference's LinearMath): historyLength is bit-identical to the oracle, every implementation takes the same branches
// (the same reasoning as the exact class for primary hits) — including the normal sign test: an axis-aligned voxel scene
// is full of exact cancellations (dot == 0), which fast arithmetic turns into +-1e-9. Only the 12-tap history fetches
// and the blends (no decisions) are the fast class.
#include "vpt_denoise_common.cuh"
#include <cmath>
#include <cstring>

namespace vpt {

// ------------------------------------------------------------------------------------------------ taps
// kIn: the caller has established that every tap of the pixel's footprint lies inside the image, so the clamp-to-edge of the
// reference's surface reads (cudaBoundaryModeClamp) is the identity and is skipped: a tap is then base + compile-time offset.
// 52 taps per pixel -> ~260 fewer integer instructions on all but the border ring (ncu r1l: the pass is issue-bound).
template <bool kIn> VPT_DEV f4 tap4(const float4 *b, int W, int H, int x, int y)
{
    if (!kIn) { x = clampi(x, 0, W - 1); y = clampi(y, 0, H - 1); }
    return F4(__ldg(b + (y * W + x)));
}
template <bool kIn> VPT_DEV float tap1(const float *b, int W, int H, int x, int y)
{
    if (!kIn) { x = clampi(x, 0, W - 1); y = clampi(y, 0, H - 1); }
    return __ldg(b + (y * W + x));
}

// ------------------------------------------------------------------------------------------------ samplers
VPT_DEV void bilinearSetup(f2 uv, int W, int H, f2 &f, int &tx0, int &ty0)
{
    f2 UV = {uv.x * W, uv.y * H};
    f2 tc = {floorf(UV.x - 0.5f) + 0.5f, floorf(UV.y - 0.5f) + 0.5f};
    f = UV - tc;
    tx0 = (int)floorf(UV.x - 0.5f); ty0 = (int)floorf(UV.y - 0.5f);
}
VPT_DEV f4 bilinearWeight(f2 uv, int W, int H)
{
    f2 f; int a, b; bilinearSetup(uv, W, H, f, a, b);
    f2 w1 = f, w0 = {1.0f - f.x, 1.0f - f.y};
    return {w0.x * w0.y, w1.x * w0.y, w0.x * w1.y, w1.x * w1.y};
}
template <bool kIn> VPT_DEV f4 sampleBilinearCustom4(const float4 *tex, f2 uv, int W, int H, f4 cw)
{
    f2 f; int x0, y0; bilinearSetup(uv, W, H, f, x0, y0);
    f2 w1 = f, w0 = {1.0f - f.x, 1.0f - f.y};
    const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
    const float ws[4] = {w0.x * w0.y * cw.x, w1.x * w0.y * cw.y, w0.x * w1.y * cw.z, w1.x * w1.y * cw.w};
    f4 out = F4(0.0f); float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        f4 v = tap4<kIn>(tex, W, H, xs[i], ys[i]);
        float w = max1f(ws[i], 1e-6f);
        sum += w; out += v * w;
    }
    return out / sum;
}
VPT_DEV float sampleBilinearCustom1(const float *tex, f2 uv, int W, int H, f4 cw)
{
    f2 f; int x0, y0; bilinearSetup(uv, W, H, f, x0, y0);
    f2 w1 = f, w0 = {1.0f - f.x, 1.0f - f.y};
    const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
    const float ws[4] = {w0.x * w0.y * cw.x, w1.x * w0.y * cw.y, w0.x * w1.y * cw.z, w1.x * w1.y * cw.w};
    float out = 0.0f, sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        float v = ld1(tex, W, H, xs[i], ys[i]);
        float w = max1f(ws[i], 1e-6f);
        sum += w; out += v * w;
    }
    return out / sum;
}
template <bool kIn> VPT_DEV f4 sampleBicubic12(const float4 *tex, f2 uv, int W, int H)
{
    f2 f; int x1, y1; bilinearSetup(uv, W, H, f, x1, y1);
    f2 f2_ = f * f, f3_ = f2_ * f;
    f2 w0 = {f2_.x - 0.5f * (f3_.x + f.x), f2_.y - 0.5f * (f3_.y + f.y)};
    f2 w1 = {1.5f * f3_.x - 2.5f * f2_.x + 1.0f, 1.5f * f3_.y - 2.5f * f2_.y + 1.0f};
    f2 w3 = {0.5f * (f3_.x - f2_.x), 0.5f * (f3_.y - f2_.y)};
    f2 w2 = {1.0f - w0.x - w1.x - w3.x, 1.0f - w0.y - w1.y - w3.y};
    const int x0 = x1 - 1, x2 = x1 + 1, x3 = x1 + 2, y0 = y1 - 1, y2 = y1 + 1, y3 = y1 + 2;
    const int xs[12] = {x1, x2, x0, x1, x2, x3, x0, x1, x2, x3, x1, x2};
    const int ys[12] = {y0, y0, y1, y1, y1, y1, y2, y2, y2, y2, y3, y3};
    const float ws[12] = {w1.x * w0.y, w2.x * w0.y, w0.x * w1.y, w1.x * w1.y, w2.x * w1.y, w3.x * w1.y,
                          w0.x * w2.y, w1.x * w2.y, w2.x * w2.y, w3.x * w2.y, w1.x * w3.y, w2.x * w3.y};
    f4 out = F4(0.0f); float sum = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) { sum += ws[i]; out += tap4<kIn>(tex, W, H, xs[i], ys[i]) * ws[i]; }
    return out / sum;
}
VPT_DEV f3 sampleSmoothStep3(const float4 *tex, f2 uv, int W, int H)
{
    f2 f; int x0, y0; bilinearSetup(uv, W, H, f, x0, y0);
    f2 f2_ = f * f, f3_ = f2_ * f;
    f2 w1 = {-2.0f * f3_.x + 3.0f * f2_.x, -2.0f * f3_.y + 3.0f * f2_.y};
    f2 w0 = {1.0f - w1.x, 1.0f - w1.y};
    const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
    const float ws[4] = {w0.x * w0.y, w1.x * w0.y, w0.x * w1.y, w1.x * w1.y};
    f3 out = F3(0.0f); float sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { sum += ws[i]; out += xyz(ld4(tex, W, H, xs[i], ys[i])) * ws[i]; }
    return out / sum;
}
VPT_DEV float parallaxInPixels(f3 X, f2 uvZero, const Cam &cam, f2 rectSize)
{
    f2 uv = worldDirectionToUV(cam, normalize(X - cam.pos));
    f2 d = (uv - uvZero) * rectSize;
    return sqrtf(d.x * d.x + d.y * d.y);
}

// ------------------------------------------------------------------------------------------------ exact-class pieces
VPT_DEV f3 exSub3(f3 a, f3 b) { return {ex::subf(a.x, b.x), ex::subf(a.y, b.y), ex::subf(a.z, b.z)}; }
VPT_DEV float exLerp(float a, float b, float w) { return ex::addf(a, ex::mulf(w, ex::subf(b, a))); }
struct ExBilinear { float fx, fy; int x0, y0; };
VPT_DEV ExBilinear exBilinearSetup(f2 uv, int W, int H)
{
    const float U = ex::mulf(uv.x, (float)W), V = ex::mulf(uv.y, (float)H);
    const float flx = floorf(ex::subf(U, 0.5f)), fly = floorf(ex::subf(V, 0.5f));
    return {ex::subf(U, ex::addf(flx, 0.5f)), ex::subf(V, ex::addf(fly, 0.5f)), (int)flx, (int)fly};
}
// bilinearWeight (Sampler.h:328-348)
VPT_DEV f4 exBilinearWeight(const ExBilinear &b)
{
    const float w1x = b.fx, w1y = b.fy, w0x = ex::subf(1.0f, b.fx), w0y = ex::subf(1.0f, b.fy);
    return {ex::mulf(w0x, w0y), ex::mulf(w1x, w0y), ex::mulf(w0x, w1y), ex::mulf(w1x, w1y)};
}
// sampleBilinearCustom1 (Sampler.h:452-498): custom tap weights floored at 1e-6, normalised
template <bool kIn> VPT_DEV float exSampleBilinearCustom1(const float *tex, const ExBilinear &b, int W, int H, f4 cw)
{
    const f4 bw = exBilinearWeight(b);
    const int xs[4] = {b.x0, b.x0 + 1, b.x0, b.x0 + 1}, ys[4] = {b.y0, b.y0, b.y0 + 1, b.y0 + 1};
    const float ws[4] = {ex::mulf(bw.x, cw.x), ex::mulf(bw.y, cw.y), ex::mulf(bw.z, cw.z), ex::mulf(bw.w, cw.w)};
    float out = 0.0f, sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        const float v = tap1<kIn>(tex, W, H, xs[i], ys[i]);
        const float w = max1f(ws[i], 1e-6f);
        sum = ex::addf(sum, w); out = ex::addf(out, ex::mulf(v, w));
    }
    return ex::divf(out, sum);
}

VPT_DEV f3 exAdd3(f3 a, f3 b) { return {ex::addf(a.x, b.x), ex::addf(a.y, b.y), ex::addf(a.z, b.z)}; }
VPT_DEV f3 exScale3(f3 a, float s) { return {ex::mulf(a.x, s), ex::mulf(a.y, s), ex::mulf(a.z, s)}; }
VPT_DEV f3 exDiv3(f3 a, float s) { const ex::rcpx k = ex::rcpPrepare(s); return {ex::divBy(k, a.x), ex::divBy(k, a.y), ex::divBy(k, a.z)}; }
// Quat (LinearMath.h:1311-1366) in the exact class: an axis-aligned voxel scene is full of EXACT cancellations (a 3x3 normal
// average perpendicular to the history normal gives dot == 0), so the sign test below must round like the oracle
VPT_DEV quat exQmul(quat p, quat q)
{
    return {exAdd3(exAdd3(exScale3(q.v, p.w), exScale3(p.v, q.w)), ex::cross(p.v, q.v)), ex::subf(ex::mulf(p.w, q.w), ex::dot(p.v, q.v))};
}
VPT_DEV f3 exQrotate(quat q, f3 v) { return exQmul(exQmul(q, quat{v, 0.0f}), quat{-q.v, q.w}).v; }
// parallaxInPixels (TemporalAccumulation.h:29-40): uv of X seen from camPos minus uvZero, in pixels
VPT_DEV float exParallaxInPixels(f3 X, f2 uvZero, f3 camPos, const float *worldToUv, f2 rectSize)
{
    const f3 h = ex::mulMat3(worldToUv, ex::normalize(exSub3(X, camPos)));
    const ex::rcpx k = ex::rcpPrepare(h.z);
    const float dx = ex::mulf(ex::subf(ex::divBy(k, h.x), uvZero.x), rectSize.x), dy = ex::mulf(ex::subf(ex::divBy(k, h.y), uvZero.y), rectSize.y);
    return __fsqrt_rn(ex::addf(ex::mulf(dx, dx), ex::mulf(dy, dy)));
}

// SampleBicubicSmoothStep (Sampler.h:652-698), xyz only
template <bool kIn> VPT_DEV f3 exSampleSmoothStep3(const float4 *tex, const ExBilinear &b, int W, int H)
{
    const float fx2 = ex::mulf(b.fx, b.fx), fy2 = ex::mulf(b.fy, b.fy), fx3 = ex::mulf(fx2, b.fx), fy3 = ex::mulf(fy2, b.fy);
    const float w1x = ex::addf(ex::mulf(-2.0f, fx3), ex::mulf(3.0f, fx2)), w1y = ex::addf(ex::mulf(-2.0f, fy3), ex::mulf(3.0f, fy2));
    const float w0x = ex::subf(1.0f, w1x), w0y = ex::subf(1.0f, w1y);
    const int xs[4] = {b.x0, b.x0 + 1, b.x0, b.x0 + 1}, ys[4] = {b.y0, b.y0, b.y0 + 1, b.y0 + 1};
    const float ws[4] = {ex::mulf(w0x, w0y), ex::mulf(w1x, w0y), ex::mulf(w0x, w1y), ex::mulf(w1x, w1y)};
    f3 out = F3(0.0f); float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { sum = ex::addf(sum, ws[i]); out = exAdd3(out, exScale3(xyz(tap4<kIn>(tex, W, H, xs[i], ys[i])), ws[i])); }
    return exDiv3(out, sum);
}

// ------------------------------------------------------------------------------------------------ temporal
struct TemporalArgs
{
    int W, H, rowBegin, rowEnd;
    VptCamera cam, prevCam;
    float denoisingRange, disocclusionThreshold, disocclusionThresholdAlternate, maxAccum, maxFastAccum;
    const float *depth, *prevDepth, *prevHistLen;
    const float4 *normalRough, *prevNormalRough, *illum, *prevIllum, *prevFast;
    float4 *ping, *pong;
    float *histLen;
    unsigned *fixCount; int *fixList; // pixels that end with historyLength <= 4: HistoryFix's work list
    float prevToCur[4];               // Quat rotationBetween(prevCam.dir, cam.dir), xyz + w
    float invW, invH, disThr, unproject; // launch-uniform IEEE quotients / sums, evaluated once on the host
    int identityRotation, staticCamera;
};
#ifndef VPT_TEMPORAL_MINB
#define VPT_TEMPORAL_MINB 4 // measured: 2 -> 223 us, 3 (80 regs) -> 176, 4 (64 regs) -> 162, 5 -> 162
#endif
// Everything after the reprojected footprint is known. kIn: every tap of the pixel (its 3x3 normal neighbourhood and the 4x4
// previous-frame footprint around (bx, by)) lies inside the image — warp-uniform, see the kernel.
struct TemporalPre
{
    int x, y; size_t pix; float z;
    f3 n, viewVec, worldPos, Vprev; float NoV; f2 prevUV; ExBilinear bil;
};
template <bool kIn> VPT_DEV void temporalRest(const TemporalArgs &a, const TemporalPre &t)
{
    const int W = a.W, H = a.H, x = t.x, y = t.y;
    const size_t pix = t.pix;
    const float z = t.z;
    const f3 n = t.n, worldPos = t.worldPos, Vprev = t.Vprev;
    const float NoV = t.NoV;
    const f2 prevUV = t.prevUV;
    const ExBilinear bil = t.bil;
    const f3 camPosV = F3(a.cam.pos[0], a.cam.pos[1], a.cam.pos[2]), prevCamPos = F3(a.prevCam.pos[0], a.prevCam.pos[1], a.prevCam.pos[2]);
    // launch-uniform: rotation between the previous and current view directions, evaluated on the host (hostRotationBetween)
    const quat prevToCur = {F3(a.prevToCur[0], a.prevToCur[1], a.prevToCur[2]), a.prevToCur[3]};
    // ---- fast class from here on, except where noted
    const f3 illum = xyz(__ldg(a.illum + pix));
    f3 nAvg = n;
#pragma unroll
    for (int i = -1; i <= 1; ++i)
#pragma unroll
        for (int j = -1; j <= 1; ++j)
        {
            if (i == 0 && j == 0) continue;
            nAvg = exAdd3(nAvg, xyz(tap4<kIn>(a.normalRough, W, H, x + i, y + j)));
        }
    nAvg = exDiv3(nAvg, 9.0f);
    const float m1 = luminance(illum), m2 = m1 * m1;
    // ---- exact class: everything a tap-validity decision depends on (parallax, disocclusion thresholds, expected depth)
    const f2 pixelUv = {ex::mulf(ex::addf(float(x), 0.5f), a.invW), ex::mulf(ex::addf(float(y), 0.5f), a.invH)};
    const f3 prevWorldPos = worldPos;
    const f3 camDelta = exSub3(prevCamPos, camPosV);
    const f2 rect = {(float)W, (float)H};
    float par1, par2;
    if (a.staticCamera)
    {
        // prevCam == cam bit for bit: camDelta = 0, both parallax terms re-project the same point with the same camera, i.e.
        // they repeat the operations that produced prevUV. par2 = |prevUV - prevUV| = 0, par1 = |prevUV - pixelUv| in pixels:
        // the same values without two more normalize / Mat3 / divide chains (launch-uniform branch).
        const float dx = ex::mulf(ex::subf(prevUV.x, pixelUv.x), rect.x), dy = ex::mulf(ex::subf(prevUV.y, pixelUv.y), rect.y);
        par1 = __fsqrt_rn(ex::addf(ex::mulf(dx, dx), ex::mulf(dy, dy)));
        par2 = 0.0f;
    }
    else
    {
        par1 = exParallaxInPixels(exAdd3(prevWorldPos, camDelta), pixelUv, prevCamPos, a.prevCam.worldToUv, rect);
        par2 = exParallaxInPixels(exSub3(prevWorldPos, camDelta), prevUV, camPosV, a.cam.worldToUv, rect);
    }
    const float parMax = fmaxr(par1, par2);
    const float disThr = a.disThr; // lerp(threshold + 1.5/H, alternate + 1.5/H, 0): launch-uniform, from the host
    const f3 toPrev = exSub3(prevWorldPos, prevCamPos);
    const float estPrevDepth = __fsqrt_rn(ex::dot(toPrev, toPrev));
    const int bx = bil.x0, by = bil.y0;
    const float pixelSize = ex::mulf(a.unproject, z); // unproject = tanHalfFov.x / (resolution.x / 2), from the host
    const float frustumSize = ex::mulf(pixelSize, (float)min(W, H));
    const float slopeScale = ex::divf(1.0f, exLerp(exLerp(0.05f, 1.0f, NoV), 1.0f, saturate(ex::divf(parMax, 30.0f))));
    float thr[4];
    {
        const float base = ex::mulf(saturate(ex::mulf(disThr, slopeScale)), frustumSize);
        const int px0 = bx, py0 = by, px1 = bx + 1, py1 = by + 1;
        float rx0 = (px0 >= 0) ? 1.0f : 0.0f, ry0 = (py0 >= 0) ? 1.0f : 0.0f, rx1 = (px1 >= 0) ? 1.0f : 0.0f, ry1 = (py1 >= 0) ? 1.0f : 0.0f;
        rx0 *= (px0 < W) ? 1.0f : 0.0f; ry0 *= (py0 < H) ? 1.0f : 0.0f; rx1 *= (px1 < W) ? 1.0f : 0.0f; ry1 *= (py1 < H) ? 1.0f : 0.0f;
        const float inScreen[4] = {rx0 * ry0, rx1 * ry0, rx0 * ry1, rx1 * ry1};
#pragma unroll
        for (int i = 0; i < 4; ++i) thr[i] = ex::subf(ex::mulf(base, inScreen[i]), 1e-6f);
    }
    const f3 curNormalAvg = ex::normalize(nAvg);
    const int bic[4][2][2] = {{{0, -1}, {-1, 0}}, {{1, -1}, {2, 0}}, {{-1, 1}, {0, 2}}, {{2, 1}, {1, 2}}};
    const int bilTap[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
    float bicubicValid = 1.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
        {
            float pz = tap1<kIn>(a.prevDepth, W, H, bx + bic[i][j][0], by + bic[i][j][1]);
            bicubicValid *= fabsf(ex::subf(pz, estPrevDepth)) > thr[i] ? 0.0f : 1.0f;
        }
    float tv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        float pz = tap1<kIn>(a.prevDepth, W, H, bx + bilTap[i][0], by + bilTap[i][1]);
        float v = fabsf(ex::subf(pz, estPrevDepth)) > thr[i] ? 0.0f : 1.0f;
        bicubicValid *= v; tv[i] = v;
    }
    f4 tapsValid = {tv[0], tv[1], tv[2], tv[3]};
    const f3 prevNFlat = ex::normalize(exSampleSmoothStep3<kIn>(a.prevNormalRough, bil, W, H));
    // identity rotation (static camera: q = (0,0,0,1) exactly): q v q^-1 = v up to the sign of a zero, which the sign test
    // below cannot see; launch-uniform branch
    const f3 prevNRot = ex::normalize(a.identityRotation ? prevNFlat : exQrotate(prevToCur, prevNFlat));
    if (ex::dot(curNormalAvg, prevNRot) < 0.0f) { tapsValid = F4(0.0f); bicubicValid = 0.0f; }
    // ---- fast class: the history fetches and blends
    const bool useBicubic = bicubicValid > 0;
    f4 prevIllum; f3 prevFast;
    if (useBicubic)
    {
        prevIllum = sampleBicubic12<kIn>(a.prevIllum, prevUV, W, H);
        prevFast = xyz(sampleBicubic12<kIn>(a.prevFast, prevUV, W, H));
    }
    else
    {
        prevIllum = sampleBilinearCustom4<kIn>(a.prevIllum, prevUV, W, H, tapsValid);
        prevFast = xyz(sampleBilinearCustom4<kIn>(a.prevFast, prevUV, W, H, tapsValid));
    }
    prevIllum = max4f(prevIllum, F4(0.0f));
    prevFast = max3f(prevFast, F3(0.0f));
    float reprojFound = (bicubicValid > 0.0f) ? 2.0f : 1.0f;
    // ---- exact class: historyLength
    const f4 bw = exBilinearWeight(bil);
    float footprintQuality = (bicubicValid > 0) ? 1.0f : ex::addf(ex::addf(ex::addf(bw.x, bw.y), bw.z), bw.w);
    float historyLength;
    if (dot4(tapsValid, F4(1.0f)) == 0.0f) { reprojFound = 0.0f; footprintQuality = 0.0f; historyLength = 0.0f; }
    else historyLength = exSampleBilinearCustom1<kIn>(a.prevHistLen, bil, W, H, tapsValid);
    historyLength = ex::addf(historyLength, 1.0f);
    const float NoVprev = fabsf(ex::dot(n, Vprev));
    float sizeQuality = ex::divf(ex::addf(NoVprev, 1e-3f), ex::addf(NoV, 1e-3f));
    sizeQuality = ex::mulf(sizeQuality, sizeQuality); sizeQuality = ex::mulf(sizeQuality, sizeQuality);
    footprintQuality = ex::mulf(footprintQuality, exLerp(0.1f, 1.0f, saturate(sizeQuality)));
    if (footprintQuality < 1.0f) { historyLength = ex::mulf(historyLength, __fsqrt_rn(footprintQuality)); historyLength = fmaxr(historyLength, 1.0f); }
    historyLength = fminr(historyLength, a.maxAccum);
    // ---- fast class
    const float alpha = (reprojFound > 0) ? fmaxr(1.0f / (a.maxAccum + 1.0f), 1.0f / historyLength) : 1.0f;
    const float alphaFast = (reprojFound > 0) ? fmaxr(1.0f / (a.maxFastAccum + 1.0f), 1.0f / historyLength) : 1.0f;
    a.ping[pix] = toFloat4(lerp4(prevIllum, F4(illum, m2), alpha));
    a.pong[pix] = toFloat4(F4(lerp3(prevFast, illum, alphaFast), 0.0f));
    a.histLen[pix] = historyLength;
    if (historyLength <= 4.0f)
    {
        // warp-aggregated append (one atomic per warp)
        const unsigned m = __activemask();
        const int lane = (threadIdx.y * kBX + threadIdx.x) & 31, leader = __ffs(m) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(a.fixCount, (unsigned)__popc(m));
        base = __shfl_sync(m, base, leader);
        a.fixList[base + __popc(m & ((1u << lane) - 1u))] = (int)pix;
    }
}

__global__ void __launch_bounds__(kBX *kBY, VPT_TEMPORAL_MINB) temporalKernel(const __grid_constant__ TemporalArgs a)
{
    const int W = a.W, H = a.H;
    PIXEL_GUARD(W, a.rowBegin, a.rowEnd)
    const float z = __ldg(a.depth + pix);
    if (z > a.denoisingRange) return;
    TemporalPre t;
    t.x = x; t.y = y; t.pix = pix; t.z = z;
    t.n = xyz(__ldg(a.normalRough + pix));
    // ---- exact class: the chain that decides historyLength
    const f3 camPosV = F3(a.cam.pos[0], a.cam.pos[1], a.cam.pos[2]), prevCamPos = F3(a.prevCam.pos[0], a.prevCam.pos[1], a.prevCam.pos[2]);
    const f2 curUV = {ex::mulf(ex::addf(float(x), 0.5f), a.cam.inversedResolution[0]), ex::mulf(ex::addf(float(y), 0.5f), a.cam.inversedResolution[1])};
    t.viewVec = ex::normalize(ex::mulMat3(a.cam.uvToWorld, F3(curUV.x, curUV.y, 1.0f)));
    t.worldPos = ex::pointAt(camPosV, t.viewVec, z);
    const f3 V = -ex::normalize(t.viewVec);
    t.NoV = fabsf(ex::dot(t.n, V));
    t.Vprev = ex::normalize(exSub3(t.worldPos, prevCamPos));
    const f3 hUv = ex::mulMat3(a.prevCam.worldToUv, t.Vprev);
    {
        const ex::rcpx k = ex::rcpPrepare(hUv.z);
        t.prevUV = {ex::divBy(k, hUv.x), ex::divBy(k, hUv.y)};
    }
    t.bil = exBilinearSetup(t.prevUV, W, H);
    // Every tap in range (the pixel's 3x3 normals and the 4x4 footprint of the reprojection)? Then no tap needs the clamp-to-edge:
    // decided per WARP so the two instances never diverge (border warps and large reprojections take the clamped one).
    const int bx = t.bil.x0, by = t.bil.y0;
    // (one pixel of slack: the fast-class bilinearSetup of the 12-tap fetch may floor to the neighbouring texel of the exact-class one)
    const bool in = x >= 1 && x + 1 < W && y >= 1 && y + 1 < H && bx >= 2 && bx + 3 < W && by >= 2 && by + 3 < H;
    if (__all_sync(__activemask(), in)) temporalRest<true>(a, t);
    else temporalRest<false>(a, t);
}

// Quat::rotationBetween (LinearMath.h:1311-1366) on the host (same IEEE operations: products and sums rounded one by one, error-free transforms through
// fmaf, correctly rounded sqrt and division), so the launch-uniform quaternion costs no barrier in the kernel. It used to be
// computed by thread 0 of every CTA behind a __syncthreads (12 % of the kernel's stall samples, ncu r1k).
namespace {
struct HC { float v, err; };
inline HC hTwoProd(float a, float b) { volatile float ab = a * b; return {ab, std::fmaf(a, b, -ab)}; }
inline float hDop(float a, float b, float c, float d)
{
    volatile float cd = c * d;
    const float err = std::fmaf(-c, d, cd);
    const float r = std::fmaf(a, b, -cd);
    volatile float s = r + err;
    return s;
}
inline HC hTwoSum(float a, float b)
{
    volatile float s = a + b, delta = s - a;
    volatile float t0 = s - delta, t1 = a - t0, t2 = b - delta, e = t1 + t2;
    return {s, e};
}
inline float hInner3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    const HC p0 = hTwoProd(a0, b0), p1 = hTwoProd(a1, b1), p2 = hTwoProd(a2, b2);
    const HC s12 = hTwoSum(p1.v, p2.v);
    volatile float e0 = p2.err + s12.err, e1 = p1.err + e0;
    const HC s = hTwoSum(p0.v, s12.v);
    volatile float e2 = e1 + s.err, e3 = p0.err + e2, r = s.v + e3;
    return r;
}
void hostRotationBetween(const float *p, const float *q, float *out4)
{
    const float cx = hDop(p[1], q[2], p[2], q[1]), cy = hDop(p[2], q[0], p[0], q[2]), cz = hDop(p[0], q[1], p[1], q[0]);
    const float pp = hInner3(p[0], p[0], p[1], p[1], p[2], p[2]), qq = hInner3(q[0], q[0], q[1], q[1], q[2], q[2]);
    const float pq = hInner3(p[0], q[0], p[1], q[1], p[2], q[2]);
    volatile float ppqq = pp * qq;
    volatile float w = std::sqrt((float)ppqq) + pq;
    volatile float xx = cx * cx, yy = cy * cy, zz = cz * cz, ww = w * w;
    volatile float s0 = xx + yy, s1 = s0 + zz, s2 = s1 + ww;
    const float n = std::sqrt((float)s2);
    out4[0] = cx / n; out4[1] = cy / n; out4[2] = cz / n; out4[3] = w / n;
}
} // namespace

cudaError_t launchTemporal(const DenoiseLaunch &d)
{
    const dim3 grid((d.width + kBX - 1) / kBX, (d.rowEnd - d.rowBegin + kBY - 1) / kBY), block(kBX, kBY);
    TemporalArgs a;
    a.W = d.width; a.H = d.height; a.rowBegin = d.rowBegin; a.rowEnd = d.rowEnd;
    a.cam = d.cam; a.prevCam = d.prevCam;
    a.denoisingRange = d.p.denoisingRange; a.disocclusionThreshold = d.p.disocclusionThreshold;
    a.disocclusionThresholdAlternate = d.p.disocclusionThresholdAlternate;
    a.maxAccum = d.p.maxAccumulatedFrameNum; a.maxFastAccum = d.p.maxFastAccumulatedFrameNum;
    a.depth = d.b.cur.depth; a.prevDepth = d.b.prev.depth; a.prevHistLen = d.b.prevHistoryLength;
    a.normalRough = d.b.cur.normalRoughness; a.prevNormalRough = d.b.prev.normalRoughness;
    a.illum = d.b.illumination; a.prevIllum = d.b.prevIllum; a.prevFast = d.b.prevFastIllum;
    a.ping = d.b.ping; a.pong = d.b.pong; a.histLen = d.b.historyLength; a.fixCount = d.counters + 1; a.fixList = d.fixList;
    hostRotationBetween(d.prevCam.dir, d.cam.dir, a.prevToCur);
    a.staticCamera = (std::memcmp(d.cam.pos, d.prevCam.pos, sizeof d.cam.pos) == 0 && std::memcmp(d.cam.worldToUv, d.prevCam.worldToUv, sizeof d.cam.worldToUv) == 0) ? 1 : 0;
    a.identityRotation = (a.prevToCur[0] == 0.0f && a.prevToCur[1] == 0.0f && a.prevToCur[2] == 0.0f && a.prevToCur[3] == 1.0f) ? 1 : 0;
    {
        volatile float invW = 1.0f / (float)d.width, invH = 1.0f / (float)d.height, q = 1.5f / (float)d.height;
        volatile float thrBonus = d.p.disocclusionThreshold + q, thrAlt = d.p.disocclusionThresholdAlternate + q;
        volatile float diff = thrAlt - thrBonus, scaled = 0.0f * diff, dis = thrBonus + scaled; // exLerp(thrBonus, thrAlt, 0)
        volatile float half = d.cam.resolution[0] / 2.0f, unproject = d.cam.tanHalfFov[0] / half;
        a.invW = invW; a.invH = invH; a.disThr = dis; a.unproject = unproject;
    }
    temporalKernel<<<grid, block, 0, d.stream>>>(a);
    return cudaGetLastError();
}

} // namespace vpt
