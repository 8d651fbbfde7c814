// ORACLE — test infrastructure only (see orc_math.h header). CPU restatement of the denoiser pass chain.
//
//   host chain          /root/reference/renderer/denoising/Denoiser.cu:24-408
//   FireflyBoilingFilter /root/reference/renderer/denoising/FireflyFilter.h:9-251
//   BufferCopySky/NonSky /root/reference/renderer/denoising/BufferCopy.h:6-34, 36-116
//   TemporalAccumulation /root/reference/renderer/denoising/TemporalAccumulation.h:8-449
//   HistoryFix           /root/reference/renderer/denoising/HistoryFix.h:6-120
//   HistoryClamping      /root/reference/renderer/denoising/HistoryClamping.h:6-219
//   AtrousSmem           /root/reference/renderer/denoising/AtrousSmem.h:9-303
//   Atrous               /root/reference/renderer/denoising/Atrous.h:6-158
//   helpers              /root/reference/renderer/denoising/DenoiserCommon.h, shaders/Sampler.h:134-188,
//                        :328-348 (GetBilinearWeight), :396-498 (custom bilinear), :576-650 (12-tap), :652-698
// All surface reads use clamp addressing (cudaBoundaryModeClamp). HitDistReconstruction and PrePass are off in
// the shipped settings and are not restated.
//
// Reference quirks kept on purpose (and mirrored by the CUDA path):
//  * Load2DUshort1 on the FLOAT material surface (HistoryFix.h:61,87; Atrous.h:47,110): reads 16-bit halves
//    of the row, i.e. u16 element x of row y — a half of the float of pixel x/2.  -> matU16().
//  * HistoryFix overwrites the responsive history in Pong for pixels with historyLength <= 4 (Denoiser.cu:216).
//  * Float3 min/max in HistoryClamping compare the x (luma) component only (LinearMath.h:526-529).
//  * TemporalAccumulation indexes the disocclusion threshold of the 8 outer bicubic taps with the loop
//    variable of the tap GROUP (TemporalAccumulation.h:129).
// Reference behaviour that is undefined / racy and is DEFINED here:
//  * FireflyBoilingFilter reads and writes illumination + reservoirs in place while neighbours read them
//    (FireflyFilter.h:151,236) and shuffles with exited lanes: the oracle reads a pre-pass snapshot and the
//    8x4 tile statistics run over the in-screen, non-sky pixels of the tile.
//  * Atrous has no out-of-range guard for pixelPos (Atrous.h:29-45); out-of-range threads have no effect.
#pragma once
#include "orc_trace.h"

namespace orc {

struct DenoisingParams // GlobalSettings.h:82-141, field order of the C ABI struct VptDenoisingParams
{
    int32_t enableHitDistanceReconstruction = 0, enablePrePass = 0, enableTemporalAccumulation = 1, enableHistoryFix = 1,
            enableHistoryClamping = 1, enableSpatialFiltering = 1, enableFireflyFilter = 1;
    float maxAccumulatedFrameNum = 30.0f, maxFastAccumulatedFrameNum = 6.0f;
    float phiLuminance = 2.0f, lobeAngleFraction = 0.5f, roughnessFraction = 0.15f, depthThreshold = 0.003f;
    int32_t atrousIterationNum = 5;
    float disocclusionThreshold = 0.01f, disocclusionThresholdAlternate = 0.05f, denoisingRange = 500000.0f;
};
static_assert(sizeof(DenoisingParams) == 68, "DenoisingParams POD");

struct DenoiseState
{
    std::vector<f4> illumOutput, ping, pong, prevIllum, prevFastIllum;
    std::vector<float> historyLength, prevHistoryLength;
    void resize(size_t n)
    {
        illumOutput.assign(n, F4(0.0f)); ping.assign(n, F4(0.0f)); pong.assign(n, F4(0.0f));
        prevIllum.assign(n, F4(0.0f)); prevFastIllum.assign(n, F4(0.0f));
        historyLength.assign(n, 0.0f); prevHistoryLength.assign(n, 0.0f);
    }
};

constexpr float kDenoisingRange = 500000.0f;

template <typename T>
inline T ldc(const std::vector<T> &b, int w, int h, int x, int y)
{
    x = clampi(x, 0, w - 1); y = clampi(y, 0, h - 1);
    return b[(size_t)y * w + x];
}
inline float matU16(const std::vector<float> &mat, int w, int h, int x, int y)
{
    y = clampi(y, 0, h - 1);
    x = clampi(x, 0, 2 * w - 1);
    const uint16_t *row = reinterpret_cast<const uint16_t *>(mat.data() + (size_t)y * w);
    return (float)row[x];
}
inline float linearStep(float a, float b, float x) { return saturate((x - a) / (b - a)); }
inline float smoothStep(float a, float b, float x) { float t = linearStep(a, b, x); return t * t * (3.0f - 2.0f * t); }
inline float acosApprox(float x) { return sqrtf(2.0f) * sqrtf(saturate(1.0f - x)); }
inline float nonExpWeight(float x, float px, float py) { return smoothStep(1.0f, 0.0f, fabsf(x * px + py)); }
inline float specLobeTanHalfAngle(float roughness, float percentOfVolume)
{
    roughness = saturate(roughness); percentOfVolume = saturate(percentOfVolume);
    return roughness * roughness * percentOfVolume / (1.0f - percentOfVolume + 1e-6f);
}
inline float normalWeightParam2(float roughness, float angleFraction)
{
    float angle = atanf(specLobeTanHalfAngle(roughness, angleFraction));
    return 1.0f / fmaxr(angle, 1e-6f);
}
inline float planeDistWeightAtrous(f3 cpos, f3 cn, f3 spos, float thr) { return fabsf(dot(spos - cpos, cn)) < thr ? 1.0f : 0.0f; }
inline f3 worldPosFromPixel(const Camera &cam, int x, int y, float depth)
{
    f2 uv = {(float(x) + 0.5f) * cam.inversedResolution.x, (float(y) + 0.5f) * cam.inversedResolution.y};
    return cam.pos + uvToWorldDirection(cam, uv) * depth;
}
inline f3 rgbToYCoCg(f3 c) { return {0.25f * (c.x + 2.0f * c.y + c.z), c.x - c.z, c.y - 0.5f * (c.x + c.z)}; }
inline f3 yCoCgToRgb(f3 c) { return {c.x + 0.5f * (c.y - c.z), c.x + 0.5f * c.z, c.x - 0.5f * (c.y + c.z)}; }
inline uint32_t seqHash(uint32_t x) { x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16; return x; }
inline uint32_t seqExplode(uint32_t x)
{
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// ------------------------------------------------------------------ FireflyBoilingFilter
inline void fireflyFilter(Scene &sc, const Camera &cam, int parity, float weightThreshold, float minWeight, float normalThreshold,
                          float depthSigma, float phiLuminance)
{
    const int W = sc.width, H = sc.height;
    const size_t npix = (size_t)W * H;
    GBufferSet &g = sc.gb[sc.cur];
    Reservoir *res = sc.reservoirs.data() + (size_t)parity * npix;
    const std::vector<f4> illumIn = sc.illumination;                 // pre-pass snapshot
    const std::vector<Reservoir> resIn(res, res + npix);
    auto valid = [&](const Reservoir &r) { return r.lightData != 0 && std::isfinite(r.weightSum) && r.weightSum > 0.0f; };
#pragma omp parallel for schedule(dynamic, 4)
    for (int ty = 0; ty < (H + 3) / 4; ++ty)
        for (int tx = 0; tx < (W + 7) / 8; ++tx)
        {
            float tileSum = 0.0f; unsigned tileCount = 0;
            // lane order = threadIdx.y*8 + threadIdx.x; the shuffle tree's summation order is restated
            float lane[32]; unsigned cnt[32];
            for (int l = 0; l < 32; ++l)
            {
                int x = tx * 8 + (l & 7), y = ty * 4 + (l >> 3);
                lane[l] = 0.0f; cnt[l] = 0;
                if (x >= W || y >= H) continue;
                if (g.depth[(size_t)y * W + x] > kDenoisingRange) continue;
                const Reservoir &r = resIn[(size_t)y * W + x];
                if (valid(r)) { lane[l] = r.weightSum; cnt[l] = 1; }
            }
            for (int off = 16; off > 0; off >>= 1)
                for (int l = 0; l < off; ++l) { lane[l] += lane[l + off]; cnt[l] += cnt[l + off]; }
            tileSum = lane[0]; tileCount = cnt[0];
            for (int l = 0; l < 32; ++l)
            {
                int x = tx * 8 + (l & 7), y = ty * 4 + (l >> 3);
                if (x >= W || y >= H) continue;
                size_t pix = (size_t)y * W + x;
                const float centerDepth = g.depth[pix];
                if (centerDepth > kDenoisingRange) continue;
                const Reservoir reservoir = resIn[pix];
                if (!valid(reservoir)) continue;
                const float currentWeight = reservoir.weightSum;
                const float neighborWeightSum = tileSum - currentWeight;
                const int neighborValidCount = (int)tileCount - 1;
                bool isFirefly = false;
                if (currentWeight >= minWeight)
                {
                    if (neighborValidCount <= 0) isFirefly = true;
                    else
                    {
                        const float avg = neighborWeightSum / float(neighborValidCount);
                        if (avg > 0.0f && currentWeight > avg * weightThreshold) isFirefly = true;
                    }
                }
                if (!isFirefly) continue;
                const f4 centerColor4 = illumIn[pix];
                const float centerLum = luminance(xyz(centerColor4));
                f3 centerNormal = xyz(g.normalRoughness[pix]);
                const float cnLen = length(centerNormal);
                if (cnLen > 0.0f) centerNormal /= cnLen; else centerNormal = {0, 1, 0};
                const float centerMaterial = g.material[pix];
                const f3 centerWorldPos = worldPosFromPixel(cam, x, y, centerDepth);
                const float gaussian[3] = {1.0f, 2.0f, 1.0f};
                f4 filteredColor = centerColor4; float filteredWeight = 1.0f;
                f4 fallbackColor = centerColor4 * (gaussian[0] * gaussian[0]); float fallbackWeight = gaussian[0] * gaussian[0];
                const float depthScale = fmaxf(fabsf(centerDepth), 1.0f);
                const float normalWeightParam = normalWeightParam2(1.0f, 0.25f);
                Reservoir best = reservoir; float bestScore = FLT_MAX; bool hasReplacement = false;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx)
                    {
                        if (dx == 0 && dy == 0) continue;
                        const int sx = x + dx, sy = y + dy;
                        if (sx < 0 || sy < 0 || sx >= W || sy >= H) continue;
                        const size_t sp = (size_t)sy * W + sx;
                        const float gw = gaussian[abs(dx)] * gaussian[abs(dy)];
                        const f4 sc4 = illumIn[sp];
                        fallbackColor += sc4 * gw; fallbackWeight += gw;
                        const float sd = g.depth[sp];
                        if (sd > kDenoisingRange) continue;
                        f3 sn = xyz(g.normalRoughness[sp]);
                        const float snLen = length(sn);
                        if (snLen <= 0.0f) continue;
                        sn /= snLen;
                        const float nd = dot(centerNormal, sn);
                        if (nd < normalThreshold) continue;
                        if (fabsf(g.material[sp] - centerMaterial) > 0.5f) continue;
                        const f3 swp = worldPosFromPixel(cam, sx, sy, sd);
                        const float geomW = planeDistWeightAtrous(centerWorldPos, centerNormal, swp, depthSigma * depthScale);
                        if (geomW <= 0.0f) continue;
                        const float normalW = nonExpWeight(acosApprox(clampf(nd, -1.0f, 1.0f)), normalWeightParam, 0.0f);
                        const float depthW = expf(-fabsf(sd - centerDepth) / (depthScale * depthSigma + 1e-6f));
                        const float lumW = expf(-fabsf(luminance(xyz(sc4)) - centerLum) * phiLuminance);
                        const float total = gw * geomW * normalW * depthW * lumW;
                        if (total > 1e-5f) { filteredColor += sc4 * total; filteredWeight += total; }
                        const Reservoir nr = resIn[sp];
                        const bool nValid = nr.lightData != 0 && std::isfinite(nr.weightSum) && nr.weightSum > 0.0f && nr.weightSum < currentWeight;
                        if (nValid)
                        {
                            const float depthTerm = fabsf(sd - centerDepth) / (depthScale + 1e-6f);
                            const float normalTerm = 1.0f - clampf(nd, 0.0f, 1.0f);
                            const float weightDiff = fabsf(nr.weightSum - currentWeight);
                            const float score = depthTerm + normalTerm + 0.25f * weightDiff;
                            if (score < bestScore) { bestScore = score; best = nr; hasReplacement = true; }
                        }
                    }
                f4 outColor;
                if (filteredWeight > 0.0f) outColor = filteredColor / filteredWeight;
                else if (fallbackWeight > 0.0f) outColor = fallbackColor / fallbackWeight;
                else outColor = centerColor4;
                sc.illumination[pix] = outColor;
                if (hasReplacement) res[pix] = best;
                else
                {
                    Reservoir cl = reservoir;
                    float avg = (neighborValidCount > 0) ? (neighborWeightSum / float(neighborValidCount)) : minWeight;
                    float target = (neighborValidCount > 0) ? (avg * weightThreshold) : minWeight;
                    target = fmaxf(target, minWeight);
                    cl.weightSum = fminf(cl.weightSum, target);
                    res[pix] = cl;
                }
            }
        }
}

// ------------------------------------------------------------------ sampling helpers (Sampler.h)
struct BilinearTaps { int x0, y0; float w[4]; };
inline void bilinearSetup(f2 uv, int W, int H, f2 &f, int &tx0, int &ty0)
{
    f2 UV = {uv.x * W, uv.y * H};
    f2 tc = {std::floor(UV.x - 0.5f) + 0.5f, std::floor(UV.y - 0.5f) + 0.5f};
    f = UV - tc;
    tx0 = (int)std::floor(UV.x - 0.5f); ty0 = (int)std::floor(UV.y - 0.5f);
}
inline f4 bilinearWeight(f2 uv, int W, int H)
{
    f2 f; int a, b; bilinearSetup(uv, W, H, f, a, b);
    f2 w1 = f, w0 = {1.0f - f.x, 1.0f - f.y};
    return {w0.x * w0.y, w1.x * w0.y, w0.x * w1.y, w1.x * w1.y};
}
inline f4 sampleBilinearCustom4(const std::vector<f4> &tex, f2 uv, int W, int H, f4 cw)
{
    f2 f; int x0, y0; bilinearSetup(uv, W, H, f, x0, y0);
    f2 w1 = f, w0 = {1.0f - f.x, 1.0f - f.y};
    const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
    const float ws[4] = {w0.x * w0.y * cw.x, w1.x * w0.y * cw.y, w0.x * w1.y * cw.z, w1.x * w1.y * cw.w};
    f4 out = F4(0.0f); float sum = 0.0f;
    for (int i = 0; i < 4; ++i)
    {
        f4 v = ldc(tex, W, H, xs[i], ys[i]);
        float w = max1f(ws[i], 1e-6f);
        sum += w; out += v * w;
    }
    return out / sum;
}
inline float sampleBilinearCustom1(const std::vector<float> &tex, f2 uv, int W, int H, f4 cw)
{
    f2 f; int x0, y0; bilinearSetup(uv, W, H, f, x0, y0);
    f2 w1 = f, w0 = {1.0f - f.x, 1.0f - f.y};
    const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
    const float ws[4] = {w0.x * w0.y * cw.x, w1.x * w0.y * cw.y, w0.x * w1.y * cw.z, w1.x * w1.y * cw.w};
    float out = 0.0f, sum = 0.0f;
    for (int i = 0; i < 4; ++i)
    {
        float v = ldc(tex, W, H, xs[i], ys[i]);
        float w = max1f(ws[i], 1e-6f);
        sum += w; out += v * w;
    }
    return out / sum;
}
// SampleBicubic12Taps with BoundaryFuncClamp (Sampler.h:576-650)
inline f4 sampleBicubic12(const std::vector<f4> &tex, f2 uv, int W, int H)
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
    for (int i = 0; i < 12; ++i) { sum += ws[i]; out += ldc(tex, W, H, xs[i], ys[i]) * ws[i]; }
    return out / sum;
}
// SampleBicubicSmoothStep with BoundaryFuncClamp (Sampler.h:652-698), xyz only
inline f3 sampleSmoothStep3(const std::vector<f4> &tex, f2 uv, int W, int H)
{
    f2 f; int x0, y0; bilinearSetup(uv, W, H, f, x0, y0);
    f2 f2_ = f * f, f3_ = f2_ * f;
    f2 w1 = {-2.0f * f3_.x + 3.0f * f2_.x, -2.0f * f3_.y + 3.0f * f2_.y};
    f2 w0 = {1.0f - w1.x, 1.0f - w1.y};
    const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
    const float ws[4] = {w0.x * w0.y, w1.x * w0.y, w0.x * w1.y, w1.x * w1.y};
    f3 out = F3(0.0f); float sum = 0;
    for (int i = 0; i < 4; ++i) { sum += ws[i]; out += xyz(ldc(tex, W, H, xs[i], ys[i])) * ws[i]; }
    return out / sum;
}

// ------------------------------------------------------------------ TemporalAccumulation
inline float parallaxInPixels(f3 X, f2 uvZero, const Camera &cam, f2 rectSize)
{
    f2 uv = worldDirectionToUV(cam, normalize(X - cam.pos));
    f2 d = (uv - uvZero) * rectSize;
    return sqrtf(d.x * d.x + d.y * d.y);
}
inline void temporalAccumulation(Scene &sc, DenoiseState &ds, const Camera &cam, const Camera &prevCam, const DenoisingParams &p)
{
    const int W = sc.width, H = sc.height;
    const GBufferSet &g = sc.gb[sc.cur], &pg = sc.gb[sc.cur ^ 1];
    const quat prevToCur = rotationBetween(prevCam.dir, cam.dir);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            const size_t pix = (size_t)y * W + x;
            const float z = g.depth[pix];
            if (z > p.denoisingRange) continue;
            const f2 pixelUv = {(float(x) + 0.5f) * (1.0f / (float)W), (float(y) + 0.5f) * (1.0f / (float)H)};
            const float curMat = g.material[pix]; (void)curMat;
            const f3 n = xyz(g.normalRoughness[pix]);
            const f2 curUV = {(float(x) + 0.5f) * cam.inversedResolution.x, (float(y) + 0.5f) * cam.inversedResolution.y};
            const f3 viewVec = uvToWorldDirection(cam, curUV);
            const f3 worldPos = worldPosFromPixel(cam, x, y, z);
            const f3 V = -normalize(viewVec);
            const float NoV = fabsf(dot(n, V));
            const f3 prevWorldPos = worldPos; // + motionWS (== 0)
            const f2 prevUV = worldDirectionToUV(prevCam, normalize(prevWorldPos - prevCam.pos));
            const f3 illum = xyz(sc.illumination[pix]);
            f3 nAvg = n;
            for (int i = -1; i <= 1; ++i)
                for (int j = -1; j <= 1; ++j)
                {
                    if (i == 0 && j == 0) continue;
                    nAvg += xyz(ldc(g.normalRoughness, W, H, x + i, y + j));
                }
            nAvg /= 9.0f;
            const float m1 = luminance(illum), m2 = m1 * m1;
            const f3 camDelta = prevCam.pos - cam.pos;
            const f2 rect = {(float)W, (float)H};
            const float par1 = parallaxInPixels(prevWorldPos + camDelta, pixelUv, prevCam, rect);
            const float par2 = parallaxInPixels(prevWorldPos - camDelta, prevUV, cam, rect);
            const float parMax = fmaxr(par1, par2);
            const float thrBonus = p.disocclusionThreshold + (1.5f / H);
            const float thrAltBonus = p.disocclusionThresholdAlternate + (1.5f / H);
            const float disThr = lerpf(thrBonus, thrAltBonus, 0.0f);

            // ---- loadSurfaceMotionBasedPrevData
            const f3 curNormalAvg = normalize(nAvg);
            const float estPrevDepth = length(prevWorldPos - prevCam.pos);
            const f2 prevPixF = {prevUV.x * W, prevUV.y * H};
            const int bx = (int)std::floor(prevPixF.x - 0.5f), by = (int)std::floor(prevPixF.y - 0.5f);
            const float pixelSize = pixelWorldSizeScaleToDepth(cam) * z;
            const float frustumSize = pixelSize * (float)std::min(W, H);
            const float slopeScale = 1.0f / lerpf(lerpf(0.05f, 1.0f, NoV), 1.0f, saturate(parMax / 30.0f));
            float thr[4];
            {
                const float base = saturate(disThr * slopeScale) * frustumSize;
                const int px0 = bx, py0 = by, px1 = bx + 1, py1 = by + 1;
                float rx0 = (px0 >= 0) ? 1.0f : 0.0f, ry0 = (py0 >= 0) ? 1.0f : 0.0f, rx1 = (px1 >= 0) ? 1.0f : 0.0f, ry1 = (py1 >= 0) ? 1.0f : 0.0f;
                rx0 *= (px0 < W) ? 1.0f : 0.0f; ry0 *= (py0 < H) ? 1.0f : 0.0f; rx1 *= (px1 < W) ? 1.0f : 0.0f; ry1 *= (py1 < H) ? 1.0f : 0.0f;
                const float inScreen[4] = {rx0 * ry0, rx1 * ry0, rx0 * ry1, rx1 * ry1};
                for (int i = 0; i < 4; ++i) { thr[i] = base * inScreen[i]; thr[i] -= 1e-6f; }
            }
            static const int bic[4][2][2] = {{{0, -1}, {-1, 0}}, {{1, -1}, {2, 0}}, {{-1, 1}, {0, 2}}, {{2, 1}, {1, 2}}};
            static const int bil[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
            float bicubicValid = 1.0f;
            f4 tapsValid = F4(0.0f);
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 2; ++j)
                {
                    float pz = ldc(pg.depth, W, H, bx + bic[i][j][0], by + bic[i][j][1]);
                    bicubicValid *= fabsf(pz - estPrevDepth) > thr[i] ? 0.0f : 1.0f;
                }
            float tv[4];
            for (int i = 0; i < 4; ++i)
            {
                float pz = ldc(pg.depth, W, H, bx + bil[i][0], by + bil[i][1]);
                float v = fabsf(pz - estPrevDepth) > thr[i] ? 0.0f : 1.0f;
                bicubicValid *= v; tv[i] = v;
            }
            tapsValid = {tv[0], tv[1], tv[2], tv[3]};
            const f3 prevNFlat = normalize(sampleSmoothStep3(pg.normalRoughness, prevUV, W, H));
            const f3 prevNRot = normalize(qrotate(prevToCur, prevNFlat));
            if (dot(curNormalAvg, prevNRot) < 0.0f) { tapsValid = F4(0.0f); bicubicValid = 0.0f; }
            const bool useBicubic = bicubicValid > 0;
            f4 prevIllum; f3 prevFast;
            if (useBicubic)
            {
                prevIllum = sampleBicubic12(ds.prevIllum, prevUV, W, H);
                prevFast = xyz(sampleBicubic12(ds.prevFastIllum, prevUV, W, H));
            }
            else
            {
                prevIllum = sampleBilinearCustom4(ds.prevIllum, prevUV, W, H, tapsValid);
                prevFast = xyz(sampleBilinearCustom4(ds.prevFastIllum, prevUV, W, H, tapsValid));
            }
            prevIllum = max4f(prevIllum, F4(0.0f));
            prevFast = max3f(prevFast, F3(0.0f));
            float reprojFound = (bicubicValid > 0.0f) ? 2.0f : 1.0f;
            const f4 bw = bilinearWeight(prevUV, W, H);
            float footprintQuality = (bicubicValid > 0) ? 1.0f : dot4(bw, F4(1.0f));
            float historyLength;
            if (dot4(tapsValid, F4(1.0f)) == 0.0f) { reprojFound = 0.0f; footprintQuality = 0.0f; historyLength = 0.0f; }
            else historyLength = sampleBilinearCustom1(ds.prevHistoryLength, prevUV, W, H, tapsValid);

            // ---- accumulate
            historyLength = historyLength + 1.0f;
            const f3 Vprev = normalize(prevWorldPos - prevCam.pos);
            const float NoVprev = fabsf(dot(n, Vprev));
            float sizeQuality = (NoVprev + 1e-3f) / (NoV + 1e-3f);
            sizeQuality *= sizeQuality; sizeQuality *= sizeQuality;
            footprintQuality *= lerpf(0.1f, 1.0f, saturate(sizeQuality));
            if (footprintQuality < 1.0f) { historyLength *= sqrtf(footprintQuality); historyLength = fmaxr(historyLength, 1.0f); }
            historyLength = fminr(historyLength, p.maxAccumulatedFrameNum);
            const float alpha = (reprojFound > 0) ? fmaxr(1.0f / (p.maxAccumulatedFrameNum + 1.0f), 1.0f / historyLength) : 1.0f;
            const float alphaFast = (reprojFound > 0) ? fmaxr(1.0f / (p.maxFastAccumulatedFrameNum + 1.0f), 1.0f / historyLength) : 1.0f;
            ds.ping[pix] = lerp4(prevIllum, F4(illum, m2), alpha);
            ds.pong[pix] = F4(lerp3(prevFast, illum, alphaFast), 0.0f);
            ds.historyLength[pix] = historyLength;
        }
}

// ------------------------------------------------------------------ HistoryFix
inline void historyFix(Scene &sc, DenoiseState &ds, const Camera &cam)
{
    const int W = sc.width, H = sc.height;
    const GBufferSet &g = sc.gb[sc.cur];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            const size_t pix = (size_t)y * W + x;
            const float z = g.depth[pix];
            const float hl = ds.historyLength[pix];
            if (z > kDenoisingRange || hl > 4.0f) continue;
            const float cMat = matU16(g.material, W, H, x, y);
            const f3 cn = xyz(g.normalRoughness[pix]);
            const f3 cpos = worldPosFromPixel(cam, x, y, z);
            const float depthThr = 0.003f * z;
            f4 sum = ds.ping[pix];
            float wsum = 1.0f;
            const float r = exp2f(4.0f - hl) + 1.0f;
            for (int j = -2; j <= 2; ++j)
                for (int i = -2; i <= 2; ++i)
                {
                    const int dx = (int)(i * r), dy = (int)(j * r);
                    const int sx = x + dx, sy = y + dy;
                    const bool inside = sx >= 0 && sy >= 0 && sx < W && sy < H;
                    if (i == 0 && j == 0) continue;
                    const float sMat = matU16(g.material, W, H, sx, sy);
                    const f3 sn = xyz(ldc(g.normalRoughness, W, H, sx, sy));
                    const float sz = ldc(g.depth, W, H, sx, sy);
                    const f3 spos = worldPosFromPixel(cam, sx, sy, sz);
                    float w = planeDistWeightAtrous(cpos, cn, spos, depthThr);
                    w *= powf(fmaxr(0.01f, dot(cn, sn)), fmaxr(8.0f, 0.01f));
                    w = inside ? w : 0;
                    w *= (sMat == cMat) ? 1.0f : 0.0f;
                    if (w > 1e-4f) { sum += ldc(ds.ping, W, H, sx, sy) * w; wsum += w; }
                }
            ds.pong[pix] = sum / wsum;
        }
}

// ------------------------------------------------------------------ HistoryClamping
inline void historyClamping(Scene &sc, DenoiseState &ds)
{
    const int W = sc.width, H = sc.height;
    const GBufferSet &g = sc.gb[sc.cur];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            const size_t pix = (size_t)y * W + x;
            if (g.depth[pix] > kDenoisingRange) continue;
            const float hl = ds.historyLength[pix];
            f3 rM1 = F3(0.0f), rM2 = F3(0.0f), nM1 = F3(0.0f); float nM2 = 0.0f;
            for (int dx = -2; dx <= 2; ++dx)
                for (int dy = -2; dy <= 2; ++dy)
                {
                    const f3 s = rgbToYCoCg(xyz(ldc(ds.pong, W, H, x + dx, y + dy)));
                    rM1 += s; rM2 += s * s;
                    const f3 nz = xyz(ldc(sc.illumination, W, H, x + dx, y + dy));
                    const float nl = luminance(nz);
                    nM1 += nz; nM2 += nl * nl;
                }
            rM1 /= 25.0f; rM2 /= 25.0f; nM1 /= 25.0f; nM2 /= 25.0f;
            const f3 sigma = sqrt3(max3f(F3(0.0f), rM2 - rM1 * rM1));
            f3 cmin = rM1 - 2.0f * sigma, cmax = rM1 + 2.0f * sigma;
            const f3 centerY = rgbToYCoCg(xyz(ds.pong[pix]));
            cmin = (cmin.x < centerY.x) ? cmin : centerY; // Float3 operator< compares x only
            cmax = (cmax.x > centerY.x) ? cmax : centerY;
            const f4 acc = ds.ping[pix];
            const f3 accY = rgbToYCoCg(xyz(acc));
            const f3 clampedY = clamp3(accY, cmin, cmax);
            const f3 clamped = yCoCgToRgb(clampedY);
            f4 outD = F4(clamped, acc.w);
            const f3 respCenter = yCoCgToRgb(centerY);
            f4 outR = F4(respCenter, 0.0f);
            if (hl <= 4.0f) { outD.x = outR.x; outD.y = outR.y; outD.z = outR.z; }
            float clampFactor = (clampedY.x - accY.x) == 0.0f ? 0.0f : saturate((clampedY.x - accY.x) / (centerY.x - accY.x));
            if (hl <= 4.0f) clampFactor = 1.0f;
            float histDiffL = 10.0f * 0.3f * luminance(abs3(respCenter - xyz(acc)));
            histDiffL *= clampFactor;
            if (hl <= 4.0f) histDiffL = 0.0f;
            const f3 distToNoisy = nM1 - respCenter;
            const float distToNoisyL = luminance(abs3(distToNoisy));
            f3 accel = (distToNoisyL == 0.0f) ? F3(0.0f) : distToNoisy * histDiffL / distToNoisyL;
            const float accelL = luminance(abs3(accel));
            const float ratio = (accelL == 0.0f) ? 0.0f : distToNoisyL / accelL;
            if (ratio < 1.0f) accel *= ratio;
            if (ratio <= 0.0f) accel = F3(0.0f);
            outD.x += accel.x; outD.y += accel.y; outD.z += accel.z;
            outR.x += accel.x; outR.y += accel.y; outR.z += accel.z;
            const float diffL = luminance(xyz(acc));
            const float noisyL = luminance(nM1);
            const float tSigma = 0.5f * sqrtf(fmaxr(0.0f, nM2 - noisyL * noisyL));
            const float sSigma = 4.5f * sigma.x;
            float reset = 0.5f * fmaxr(0.0f, fabsf(diffL - noisyL) - sSigma - tSigma) / (1.0e-6f + fmaxr(diffL, noisyL) + sSigma + tSigma);
            reset = saturate(reset);
            const f3 noisyC = xyz(sc.illumination[pix]);
            f3 d3 = lerp3(xyz(outD), noisyC, reset), r3 = lerp3(xyz(outR), noisyC, reset);
            outD = F4(d3, outD.w); outR = F4(r3, outR.w);
            const float outL = luminance(xyz(outD));
            outD.w += (outL * outL - diffL * diffL);
            outD.w = fmaxr(0.0f, outD.w);
            ds.prevIllum[pix] = outD;
            ds.prevFastIllum[pix] = outR;
            ds.prevHistoryLength[pix] = hl;
        }
}

// ------------------------------------------------------------------ AtrousSmem (first spatial pass)
inline void atrousSmem(Scene &sc, const std::vector<f4> &in, std::vector<f4> &out, const std::vector<float> &histLen,
                       const Camera &cam, const DenoisingParams &p)
{
    const int W = sc.width, H = sc.height;
    const GBufferSet &g = sc.gb[sc.cur];
    auto wposMat = [&](int x, int y) {
        x = clampi(x, 0, W - 1); y = clampi(y, 0, H - 1);
        size_t i = (size_t)y * W + x;
        return F4(worldPosFromPixel(cam, x, y, g.depth[i]), g.material[i]);
    };
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            const size_t pix = (size_t)y * W + x;
            const float z = g.depth[pix];
            if (z > 500000.0f) continue;
            const f3 cn = xyz(g.normalRoughness[pix]);
            const f4 cwm = wposMat(x, y);
            const f3 cpos = xyz(cwm); const float cMat = cwm.w;
            const float hl = histLen[pix];
            if (hl >= 3.0f)
            {
                f4 vsum = F4(0.0f);
                const float kern[4] = {1.0f / 4.0f, 1.0f / 8.0f, 1.0f / 8.0f, 1.0f / 16.0f};
                for (int dx = -1; dx <= 1; dx++)
                    for (int dy = -1; dy <= 1; dy++)
                        vsum += ldc(in, W, H, x + dx, y + dy) * kern[abs(dx) * 2 + abs(dy)];
                const float v1 = luminance(xyz(vsum));
                const float cVar = fmaxr(0.0f, vsum.w - v1 * v1);
                const float cLum = luminance(xyz(in[pix]));
                const float phiInv = 1.0f / fmaxr(1.0e-4f, p.phiLuminance * sqrtf(cVar));
                const float nParam = normalWeightParam2(1.0f, p.lobeAngleFraction);
                float sumW = 0.0f; f4 sum = F4(0.0f);
                const float k3[2] = {0.44198f, 0.27901f};
                const float depthThr = p.depthThreshold * z;
                for (int cx = -1; cx <= 1; cx++)
                    for (int cy = -1; cy <= 1; cy++)
                    {
                        const int sx = x + cx, sy = y + cy;
                        const bool center = (cx == 0 && cy == 0);
                        const bool inside = sx >= 0 && sy >= 0 && sx < W && sy < H;
                        const float kernel = inside ? k3[abs(cx)] * k3[abs(cy)] : 0.0f;
                        const f3 sn = xyz(ldc(g.normalRoughness, W, H, sx, sy));
                        const f4 swm = wposMat(sx, sy);
                        float geomW = planeDistWeightAtrous(cpos, cn, xyz(swm), depthThr);
                        geomW *= kernel;
                        const float normalW = nonExpWeight(acosApprox(dot(cn, sn)), nParam, 0.0f);
                        const f4 sv = ldc(in, W, H, sx, sy);
                        const float sLum = luminance(xyz(sv));
                        const float lumW = fabsf(cLum - sLum) * phiInv;
                        float w = geomW * normalW * expf(-lumW);
                        w = center ? kernel : w;
                        w *= (swm.w == cMat) ? 1.0f : 0.0f;
                        sumW += w;
                        sum += w * sv;
                    }
                sumW = fmaxr(sumW, 1e-6f);
                sum = sum / sumW;
                const float m1 = luminance(xyz(sum));
                out[pix] = F4(xyz(sum), fmaxr(0.0f, sum.w - m1 * m1));
            }
            else
            {
                float sumW = 0.0f; f3 sumI = F3(0.0f); float s1 = 0.0f, s2 = 0.0f;
                const float nParam = normalWeightParam2(1.0f, p.lobeAngleFraction);
                for (int cx = -2; cx <= 2; cx++)
                    for (int cy = -2; cy <= 2; cy++)
                    {
                        const int sx = x + cx, sy = y + cy;
                        const f3 sn = xyz(ldc(g.normalRoughness, W, H, sx, sy));
                        const float sMat = ldc(g.material, W, H, sx, sy);
                        const float normalW = nonExpWeight(acosApprox(dot(cn, sn)), nParam, 0.0f);
                        const f4 sv = ldc(in, W, H, sx, sy);
                        const float l1 = luminance(xyz(sv));
                        float w = normalW * 1.0f;
                        w *= (sMat == cMat) ? 1.0f : 0.0f;
                        sumW += w; sumI += xyz(sv) * w; s1 += l1 * w; s2 += sv.w * w;
                    }
                const float boost = fmaxr(1.0f, 4.0f / (hl + 1.0f));
                sumW = fmaxr(sumW, 1e-6f);
                sumI /= sumW; s1 /= sumW; s2 /= sumW;
                float var = fmaxr(0.0f, s2 - s1 * s1);
                var *= boost;
                out[pix] = F4(sumI, var);
            }
        }
}

// ------------------------------------------------------------------ Atrous (step 2^k passes)
inline void atrous(Scene &sc, const std::vector<f4> &in, std::vector<f4> &out, const std::vector<float> &histLen,
                   const Camera &cam, unsigned frameIndex, unsigned stepSize, const DenoisingParams &p)
{
    const int W = sc.width, H = sc.height;
    const GBufferSet &g = sc.gb[sc.cur];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            const size_t pix = (size_t)y * W + x;
            const float z = g.depth[pix];
            if (z > 500000.0f) continue;
            const float cMat = matU16(g.material, W, H, x, y);
            const f3 cn = xyz(g.normalRoughness[pix]);
            const f3 cpos = worldPosFromPixel(cam, x, y, z);
            const float hl = histLen[pix];
            float lobeFrac = p.lobeAngleFraction / sqrtf((float)stepSize);
            lobeFrac = lerpf(0.99f, lobeFrac, saturate(hl / 5.0f));
            const f4 cv = in[pix];
            const float cLum = luminance(xyz(cv));
            const float phiInv = 1.0f / fmaxr(1.0e-4f, p.phiLuminance * sqrtf(cv.w));
            const float nParam = normalWeightParam2(1.0f, lobeFrac);
            float sumW = 0.44198f * 0.44198f;
            f4 sum = cv * f4{sumW, sumW, sumW, sumW * sumW};
            const float k3[2] = {0.44198f, 0.27901f};
            const float depthThr = p.depthThreshold * z;
            int offx = 0, offy = 0;
            if (stepSize > 4)
            {
                uint32_t zorder = seqExplode((uint32_t)x) | (seqExplode((uint32_t)y) << 1);
                uint32_t seed = seqHash(frameIndex + 0x035F9F29u);
                uint32_t st = seed ^ (seqHash(zorder) + 0x9E3779B9u + (seed << 6) + (seed >> 2));
                st = seqHash(st); const float u0 = st / 4294967295.0f;
                st = seqHash(st); const float u1 = st / 4294967295.0f;
                offx = (int)((float)stepSize * 0.5f * (u0 - 0.5f));
                offy = (int)((float)stepSize * 0.5f * (u1 - 0.5f));
            }
            for (int yy = -1; yy <= 1; yy++)
                for (int xx = -1; xx <= 1; xx++)
                {
                    if (xx == 0 && yy == 0) continue;
                    const int sx = x + offx + xx * (int)stepSize, sy = y + offy + yy * (int)stepSize;
                    const bool inside = sx >= 0 && sy >= 0 && sx < W && sy < H;
                    const float kernel = k3[abs(xx)] * k3[abs(yy)];
                    const float sMat = matU16(g.material, W, H, sx, sy);
                    const f3 sn = xyz(ldc(g.normalRoughness, W, H, sx, sy));
                    const float sz = ldc(g.depth, W, H, sx, sy);
                    const f3 spos = worldPosFromPixel(cam, sx, sy, sz);
                    float geomW = planeDistWeightAtrous(cpos, cn, spos, depthThr);
                    geomW *= kernel;
                    geomW *= (inside && sz < 500000.0f) ? 1.0f : 0.0f;
                    const float normalW = nonExpWeight(acosApprox(dot(cn, sn)), nParam, 0.0f);
                    float w = geomW * normalW;
                    w *= (sMat == cMat) ? 1.0f : 0.0f;
                    if (w > 1e-4f)
                    {
                        const f4 sv = ldc(in, W, H, sx, sy);
                        const float sLum = luminance(xyz(sv));
                        const float lumW = fabsf(cLum - sLum) * phiInv; // min(+inf, .) is the identity (Atrous.h:86-87,124)
                        w *= expf(-lumW);
                        sumW += w;
                        sum += f4{w, w, w, w * w} * sv;
                    }
                }
            out[pix] = sum / f4{sumW, sumW, sumW, sumW * sumW};
        }
}

// ------------------------------------------------------------------ HitDistReconstruction<8,2> (HitDistReconstruction.h:50-161)
// 5x5 weighted fill of the hit distance (illumination.w) into IlluminationPing. Default off in the shipped settings.
inline float expApprox(float x) { return 1.0f / (x * x - x + 1.0f); }                                  // DenoiserCommon.h:62-65
inline float expWeight(float x, float px, float py, float scale) { return expApprox(-scale * fabsf(x * px + py)); } // :67-70
inline float gaussianWeight(float r) { return expf(-0.66f * r * r); }                                  // :48-51
inline float bilateralWeight(float z, float zc)                                                         // :53-60
{
    const float t = fabsf(z - zc) * (1.0f / (fmaxr(fabsf(z), fabsf(zc)) + 1e-6f));
    return linearStep(0.03f, 0.0f, t);
}
inline void hitDistReconstruction(Scene &sc, DenoiseState &ds)
{
    const int W = sc.width, H = sc.height;
    const GBufferSet &g = sc.gb[sc.cur];
    const f2 invRes = {1.0f / (float)W, 1.0f / (float)H};
    // GetNormalWeightParams(1,1,1) (:11-17): the reference evaluates this in double (atan(m*0.75/(1.0-0.75)), 0.5f*M_PI/180)
    const double lobeAngle = std::atan(1.0 * 0.75 / (1.0 - 0.75)) * (double)lerpf(1.0f, 1.0f, 1.0f);
    const float nParam = (float)(1.0 / std::max(lobeAngle, (double)(0.5f * 3.14159265358979323846 / 180.0)));
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            const size_t pix = (size_t)y * W + x;
            const float cz = fabsf(g.depth[pix]);
            if (cz > kDenoisingRange) continue;
            const f3 cn = xyz(g.normalRoughness[pix]);
            const f4 ci = sc.illumination[pix];
            const float chd = ci.w;
            const f2 pixelUv = {((float)x + 0.5f) * invRes.x, ((float)y + 0.5f) * invRes.y};
            float sumW = 1000.0f * (chd != 0.0f ? 1.0f : 0.0f);
            float sumHD = chd * sumW;
            for (int dy = -2; dy <= 2; ++dy)
                for (int dx = -2; dx <= 2; ++dx)
                {
                    if (dx == 0 && dy == 0) continue;
                    const f3 sn = xyz(ldc(g.normalRoughness, W, H, x + dx, y + dy));
                    const float sz = fabsf(ldc(g.depth, W, H, x + dx, y + dy));
                    float shd = ldc(sc.illumination, W, H, x + dx, y + dy).w;
                    const float angle = acosApprox(saturate(dot(cn, sn)));
                    const f2 uv = {pixelUv.x + (float)dx * invRes.x, pixelUv.y + (float)dy * invRes.y};
                    float w = (saturate(uv.x) == uv.x && saturate(uv.y) == uv.y) ? 1.0f : 0.0f;   // IsInScreen (:43-46)
                    w *= gaussianWeight(sqrtf((float)(dx * dx + dy * dy)) * 0.5f);
                    w *= bilateralWeight(sz, cz);
                    float dw = w * expWeight(angle, nParam, 0.0f, 3.0f);
                    shd = dw == 0.0f ? 0.0f : shd;                                                  // Denanify
                    dw *= (shd != 0.0f) ? 1.0f : 0.0f;
                    sumHD += shd * dw;
                    sumW += dw;
                }
            sumHD /= fmaxr(sumW, 1e-6f);
            ds.ping[pix] = {ci.x, ci.y, ci.z, sumHD};
        }
}

// ------------------------------------------------------------------ PrePass (PrePass.h:6-149)
// 8-tap Poisson pre-blur of IlluminationPing into IlluminationBuffer, radius from the hit distance; rotator from
// Weyl1D(0.5, frameIndex). Quirks kept: the G-buffer taps are read at round(floor(p)+0.5) = floor(p)+1 while the radiance
// tap (a linear-filter texture fetch at a texel centre = that texel, clamp addressing) is read at floor(p).
inline float weyl1D(float p, int n)                                                                      // DenoiserCommon.h:336-339
{
    const int32_t m = (int32_t)((uint32_t)n * 10368889u);   // int multiply wraps
    float ip;
    return modff(p + (float)m / 16777216.0f, &ip);
}
inline void prePassRotator(int frameIndex, float rot[4])
{
    const float angle = weyl1D(0.5f, frameIndex) * (90.0f * kPiOver180);
    const float ca = cosf(angle), sa = sinf(angle);
    rot[0] = ca; rot[1] = sa; rot[2] = -sa; rot[3] = ca;                                                 // GetRotator (:328-334)
}
inline float specMagicCurve(float roughness, float power = 0.25f)                                        // :390-395
{
    float f = 1.0f - exp2f(-200.0f * roughness * roughness);
    f *= powf(saturate(roughness), power);
    return f;
}
inline void prePass(Scene &sc, DenoiseState &ds, const Camera &cam, int frameIndex)
{
    const int W = sc.width, H = sc.height;
    const GBufferSet &g = sc.gb[sc.cur];
    const f2 invRes = {1.0f / (float)W, 1.0f / (float)H};
    static const float kPoisson8[8][3] = {
        {-0.4706069f, -0.4427112f, +0.6461146f}, {-0.9057375f, +0.3003471f, +0.9542373f}, {-0.3487388f, +0.4037880f, +0.5335386f},
        {+0.1023042f, +0.6439373f, +0.6520134f}, {+0.5699277f, +0.3513750f, +0.6695386f}, {+0.2939128f, -0.1131226f, +0.3149309f},
        {+0.7836658f, -0.4208784f, +0.8895339f}, {+0.1564120f, -0.8198990f, +0.8346850f}};
    float rot[4];
    prePassRotator(frameIndex, rot);
    const float unproject = cam.tanHalfFov.x / (cam.resolution.x / 2);                                    // Camera.h:128-131
    const float nParam = normalWeightParam2(1.0f, 0.25f * 0.5f);
    std::vector<f4> out(sc.illumination);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            const size_t pix = (size_t)y * W + x;
            const float cz = g.depth[pix];
            if (cz > kDenoisingRange) continue;
            const float cMat = g.material[pix];
            const f3 cn = xyz(g.normalRoughness[pix]);
            const f3 cpos = worldPosFromPixel(cam, x, y, cz);
            const f2 pixelUv = {((float)x + 0.5f) * invRes.x, ((float)y + 0.5f) * invRes.y};
            f4 acc = ds.ping[pix];
            const float frustumSize = (float)std::min(W, H) * unproject * cz;                             // PixelRadiusToWorld
            const float hitDist = acc.w == 0.0f ? 1.0f : acc.w;
            float blurRadius = 30.0f * saturate(hitDist / frustumSize);
            if (acc.w == 0.0f) blurRadius = fmaxr(blurRadius, 1.0f);
            // GetHitDistanceWeightParams(hitDist, 1/9) (:397-405)
            const float norm = lerpf(0.0005f, 1.0f, fminr(1.0f / 9.0f, specMagicCurve(1.0f)));
            const float hdA = 1.0f / norm, hdB = -(acc.w * hdA);
            float weightSum = 1.0f;
            for (int i = 0; i < 8; ++i)
            {
                const f2 rv = {kPoisson8[i][0] * rot[0] + kPoisson8[i][1] * rot[1], kPoisson8[i][0] * rot[2] + kPoisson8[i][1] * rot[3]};
                f2 uv = {pixelUv.x * (float)W + rv.x * blurRadius, pixelUv.y * (float)H + rv.y * blurRadius};
                uv = {floorf(uv.x) + 0.5f, floorf(uv.y) + 0.5f};
                const int gx = (int)roundf(uv.x), gy = (int)roundf(uv.y);   // G-buffer tap (samplePosInt)
                const int tx = (int)floorf(uv.x), ty = (int)floorf(uv.y);   // radiance tap (texel centre)
                uv = {uv.x * invRes.x, uv.y * invRes.y};
                const float sMat = ldc(g.material, W, H, gx, gy);
                const f3 sn = xyz(ldc(g.normalRoughness, W, H, gx, gy));
                const float sz = ldc(g.depth, W, H, gx, gy);
                const f3 spos = worldPosFromPixel(cam, gx, gy, sz);
                float w = (uv.x >= 0.0f && uv.x < 1.0f && uv.y >= 0.0f && uv.y < 1.0f) ? 1.0f : 0.0f; // IsInScreenNearest
                w *= (sz < kDenoisingRange) ? 1.0f : 0.0f;
                w *= (cMat == sMat) ? 1.0f : 0.0f;
                w *= (fabsf(dot(spos - cpos, cn)) / cz > 0.003f) ? 0.0f : 1.0f;                          // GetPlaneDistanceWeight (:300-305)
                w *= nonExpWeight(acosApprox(dot(cn, sn)), nParam, 0.0f);
                f4 sv = ldc(ds.ping, W, H, tx, ty);
                if (w == 0.0f) sv = F4(0.0f);                                                            // Denanify
                w *= lerpf(0.2f, 1.0f, expWeight(sv.w, hdA, hdB, 3.0f));
                w *= gaussianWeight(kPoisson8[i][2]);
                weightSum += w;
                acc += sv * w;
            }
            out[pix] = acc / weightSum;
        }
    sc.illumination.swap(out);
}

// ------------------------------------------------------------------ Denoiser::run (Denoiser.cu:24-408)
// iterationIndex is the value AFTER the render's post-increment (GlobalSettings::iterationIndex).
inline void denoiseRun(Scene &sc, DenoiseState &ds, const Camera &cam, const Camera &prevCam, const DenoisingParams &p,
                       int frameNum, int iterationIndex)
{
    const int W = sc.width, H = sc.height;
    const size_t npix = (size_t)W * H;
    GBufferSet &g = sc.gb[sc.cur];
    const int usedIter = iterationIndex > 0 ? iterationIndex - 1 : 0;
    if (p.enableFireflyFilter) fireflyFilter(sc, cam, usedIter & 1, 80.0f, 5.0f, 0.8f, 0.02f, p.phiLuminance);
    int finalBuf = 0; // 0 illum, 1 ping, 2 pong, 3 prevIllum
    if (p.enableHitDistanceReconstruction) { hitDistReconstruction(sc, ds); finalBuf = 1; }
    for (size_t i = 0; i < npix; ++i)
        if (g.depth[i] > kDenoisingRange) ds.illumOutput[i] = sc.illumination[i]; // BufferCopySky
    if (p.enablePrePass) prePass(sc, ds, cam, iterationIndex); // reads Ping (stale scratch when the reconstruction is off), writes Illumination
    if (frameNum == 0)
    {
        ds.prevIllum = sc.illumination; ds.prevFastIllum = sc.illumination;
        ds.historyLength.assign(npix, 0.0f); ds.prevHistoryLength.assign(npix, 0.0f);
    }
    if (p.enableTemporalAccumulation && frameNum > 0)
    {
        temporalAccumulation(sc, ds, cam, prevCam, p);
        finalBuf = 1;
        if (p.enableHistoryFix) { historyFix(sc, ds, cam); finalBuf = 2; }
        if (p.enableHistoryClamping) { historyClamping(sc, ds); finalBuf = 3; }
    }
    if (p.enableSpatialFiltering)
    {
        atrousSmem(sc, ds.prevIllum, ds.ping, ds.historyLength, cam, p);
        finalBuf = 1;
        if (p.atrousIterationNum > 0)
        {
            int idx = 1, step = 1 << idx;
            const int maxIt = p.atrousIterationNum * 2;
            while (idx < maxIt)
            {
                atrous(sc, ds.ping, ds.pong, ds.historyLength, cam, (unsigned)iterationIndex, (unsigned)step, p);
                ++idx; step = 1 << idx;
                atrous(sc, ds.pong, ds.ping, ds.historyLength, cam, (unsigned)iterationIndex, (unsigned)step, p);
                ++idx; step = 1 << idx;
            }
            atrous(sc, ds.ping, ds.pong, ds.historyLength, cam, (unsigned)iterationIndex, (unsigned)step, p);
            finalBuf = 2;
        }
    }
    const std::vector<f4> &fin = finalBuf == 1 ? ds.ping : finalBuf == 2 ? ds.pong : finalBuf == 3 ? ds.prevIllum : sc.illumination;
    for (size_t i = 0; i < npix; ++i) // BufferCopyNonSky
        if (!(g.depth[i] > kDenoisingRange))
        {
            f3 v = xyz(fin[i]) * xyz(g.albedo[i]);
            ds.illumOutput[i] = F4(v, 0.0f);
        }
    // Prev G-buffer copies (Denoiser.cu:394-407) are pointer ping-pong: the next render flips sc.cur.
}

} // namespace orc
