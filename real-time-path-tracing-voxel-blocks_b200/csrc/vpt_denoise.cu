// Denoiser pass chain for sm_100a — bandwidth-bound stencil kernels over linear fp32 planes.
// Replaces /root/reference/renderer/denoising/*.h (cudaArray surface kernels, 8x8 blocks):
//   FireflyBoilingFilter  FireflyFilter.h:9-251      -> fireflyDetectKernel + fireflyApplyKernel
//   BufferCopySky/NonSky  BufferCopy.h:6-34, 36-116  -> copySkyKernel, compositeKernel
//   TemporalAccumulation  TemporalAccumulation.h     -> temporalKernel
//   HistoryFix            HistoryFix.h:20-120        -> historyFixKernel
//   HistoryClamping       HistoryClamping.h:27-219   -> historyClampKernel
//   AtrousSmem            AtrousSmem.h:66-303        -> atrousFirstKernel
//   Atrous                Atrous.h:6-158             -> atrousKernel
// Layout: every plane is a dense row-major array (float4 or float per pixel); a warp covers 32
// x-consecutive pixels, so each float4 row access is four full 128-byte lines. Stencil taps are served
// from L1/L2 (read-only path); each plane crosses HBM once per pass. Clamp addressing everywhere, as the
// reference's cudaBoundaryModeClamp surface reads. Rows [rowBegin,rowEnd) are processed so the same kernels
// serve the row-band sharded multi-GPU path. Compiled -fmad=false (oracle-exact + - * / sqrt).
#include "vpt_kernels.h"
#include "vpt_math.cuh"

namespace vpt {

constexpr float kDenoisingRange = 500000.0f;
constexpr int kBX = 32, kBY = 8;

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

#define PIXEL_GUARD(W_, rowBegin_, rowEnd_)                       \
    const int x = blockIdx.x * kBX + threadIdx.x;                 \
    const int y = (rowBegin_) + blockIdx.y * kBY + threadIdx.y;   \
    if (x >= (W_) || y >= (rowEnd_)) return;                      \
    const size_t pix = (size_t)y * (W_) + x;

// ------------------------------------------------------------------------------------------------ firefly
VPT_DEV bool reservoirValid(const VptReservoir &r) { return r.lightData != 0 && isfinite(r.weightSum) && r.weightSum > 0.0f; }
VPT_DEV VptReservoir ldRes(const VptReservoir *p)
{
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
    VptReservoir r;
    r.lightData = __ldg(q); r.uvData = __ldg(q + 1);
    r.weightSum = __uint_as_float(__ldg(q + 2)); r.targetPdf = __uint_as_float(__ldg(q + 3)); r.M = __uint_as_float(__ldg(q + 4));
    return r;
}

// Detect + filter, reading only pre-pass values; results go to a patch list that fireflyApplyKernel commits
// (the reference's in-place read-modify-write is a race, FireflyFilter.h:151,236). 256 threads = 8 warps,
// each warp is one 8x4 tile exactly as the reference's 8x4 block, so the tile statistics partition matches.
__global__ void __launch_bounds__(256) fireflyDetectKernel(int W, int H, int rowBegin, int rowEnd, const float4 *__restrict__ illum,
                                                           const float4 *__restrict__ normalRough, const float *__restrict__ depth,
                                                           const float *__restrict__ material, const VptReservoir *__restrict__ res,
                                                           float weightThreshold, float minWeight, float normalThreshold, float depthSigma,
                                                           float phiLuminance, VptCamera camIn, FireflyPatch *patches, int *patchCount, int maxPatches)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tilesX = (W + 7) >> 3;
    const int tile = blockIdx.x * 8 + warp;
    const int tilesY = (rowEnd - rowBegin + 3) >> 2;
    if (tile >= tilesX * tilesY) return;
    const int x = (tile % tilesX) * 8 + (lane & 7);
    const int y = rowBegin + (tile / tilesX) * 4 + (lane >> 3);
    const bool inb = x < W && y < rowEnd;
    const size_t pix = (size_t)y * W + x;
    float centerDepth = 0.0f;
    VptReservoir reservoir; reservoir.lightData = 0; reservoir.uvData = 0; reservoir.weightSum = 0; reservoir.targetPdf = 0; reservoir.M = 0;
    bool participates = false;
    if (inb)
    {
        centerDepth = __ldg(depth + pix);
        if (!(centerDepth > kDenoisingRange)) { reservoir = ldRes(res + pix); participates = true; }
    }
    const bool valid = participates && reservoirValid(reservoir);
    float wsum = valid ? reservoir.weightSum : 0.0f;
    unsigned wcnt = valid ? 1u : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        wsum += __shfl_down_sync(0xffffffffu, wsum, off);
        wcnt += __shfl_down_sync(0xffffffffu, wcnt, off);
    }
    wsum = __shfl_sync(0xffffffffu, wsum, 0);
    wcnt = __shfl_sync(0xffffffffu, wcnt, 0);
    if (!valid) return;

    const float currentWeight = reservoir.weightSum;
    const float neighborWeightSum = wsum - currentWeight;
    const int neighborValidCount = (int)wcnt - 1;
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
    if (!isFirefly) return;

    const Cam cam = loadCam(camIn);
    const f4 centerColor4 = F4(__ldg(illum + pix));
    const float centerLum = luminance(xyz(centerColor4));
    f3 centerNormal = xyz(__ldg(normalRough + pix));
    const float cnLen = length(centerNormal);
    if (cnLen > 0.0f) centerNormal /= cnLen; else centerNormal = F3(0.0f, 1.0f, 0.0f);
    const float centerMaterial = __ldg(material + pix);
    const f3 centerWorldPos = worldPosFromPixel(cam, x, y, centerDepth);
    const float gaussian[3] = {1.0f, 2.0f, 1.0f};
    f4 filteredColor = centerColor4; float filteredWeight = 1.0f;
    f4 fallbackColor = centerColor4 * (gaussian[0] * gaussian[0]); float fallbackWeight = gaussian[0] * gaussian[0];
    const float depthScale = fmaxf(fabsf(centerDepth), 1.0f);
    const float normalWeightParam = normalWeightParam2(1.0f, 0.25f);
    VptReservoir best = reservoir; float bestScore = FLT_MAX; bool hasReplacement = false;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx)
        {
            if (dx == 0 && dy == 0) continue;
            const int sx = x + dx, sy = y + dy;
            if (sx < 0 || sy < 0 || sx >= W || sy >= H) continue;
            const size_t sp = (size_t)sy * W + sx;
            const float gw = gaussian[abs(dx)] * gaussian[abs(dy)];
            const f4 sc4 = F4(__ldg(illum + sp));
            fallbackColor += sc4 * gw; fallbackWeight += gw;
            const float sd = __ldg(depth + sp);
            if (sd > kDenoisingRange) continue;
            f3 sn = xyz(__ldg(normalRough + sp));
            const float snLen = length(sn);
            if (snLen <= 0.0f) continue;
            sn /= snLen;
            const float nd = dot(centerNormal, sn);
            if (nd < normalThreshold) continue;
            if (fabsf(__ldg(material + sp) - centerMaterial) > 0.5f) continue;
            const f3 swp = worldPosFromPixel(cam, sx, sy, sd);
            const float geomW = planeDistWeightAtrous(centerWorldPos, centerNormal, swp, depthSigma * depthScale);
            if (geomW <= 0.0f) continue;
            const float normalW = nonExpWeight(acosApprox(clampf(nd, -1.0f, 1.0f)), normalWeightParam, 0.0f);
            const float depthW = expf(-fabsf(sd - centerDepth) / (depthScale * depthSigma + 1e-6f));
            const float lumW = expf(-fabsf(luminance(xyz(sc4)) - centerLum) * phiLuminance);
            const float total = gw * geomW * normalW * depthW * lumW;
            if (total > 1e-5f) { filteredColor += sc4 * total; filteredWeight += total; }
            const VptReservoir nr = ldRes(res + sp);
            const bool nValid = nr.lightData != 0 && isfinite(nr.weightSum) && nr.weightSum > 0.0f && nr.weightSum < currentWeight;
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
    VptReservoir outRes;
    if (hasReplacement) outRes = best;
    else
    {
        outRes = reservoir;
        float avg = (neighborValidCount > 0) ? (neighborWeightSum / float(neighborValidCount)) : minWeight;
        float target = (neighborValidCount > 0) ? (avg * weightThreshold) : minWeight;
        target = fmaxf(target, minWeight);
        outRes.weightSum = fminf(outRes.weightSum, target);
    }
    const int slot = atomicAdd(patchCount, 1);
    if (slot < maxPatches)
    {
        patches[slot].pixel = (int)pix;
        patches[slot].color = toFloat4(outColor);
        patches[slot].reservoir = outRes;
    }
}
__global__ void fireflyApplyKernel(const FireflyPatch *__restrict__ patches, const int *__restrict__ patchCount, int maxPatches,
                                   float4 *illum, VptReservoir *res)
{
    const int n = min(*patchCount, maxPatches);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const FireflyPatch p = patches[i];
        illum[p.pixel] = p.color;
        res[p.pixel] = p.reservoir;
    }
}

// ------------------------------------------------------------------------------------------------ copies
__global__ void __launch_bounds__(kBX *kBY) copySkyKernel(int W, int rowBegin, int rowEnd, const float4 *__restrict__ illum,
                                                           const float *__restrict__ depth, float4 *__restrict__ out)
{
    PIXEL_GUARD(W, rowBegin, rowEnd)
    if (__ldg(depth + pix) <= kDenoisingRange) return;
    out[pix] = __ldg(illum + pix);
}
__global__ void __launch_bounds__(kBX *kBY) compositeKernel(int W, int rowBegin, int rowEnd, const float4 *__restrict__ fin,
                                                             const float *__restrict__ depth, const float4 *__restrict__ albedo,
                                                             float4 *__restrict__ out)
{
    PIXEL_GUARD(W, rowBegin, rowEnd)
    if (__ldg(depth + pix) > kDenoisingRange) return;
    const float4 v = __ldg(fin + pix), al = __ldg(albedo + pix);
    out[pix] = make_float4(v.x * al.x, v.y * al.y, v.z * al.z, 0.0f);
}
// frame 0 (Denoiser.cu:121-142): Illum -> PrevIllum, PrevFastIllum; HistoryLength = PrevHistoryLength = 0
__global__ void __launch_bounds__(kBX *kBY) frame0Kernel(int W, int rowBegin, int rowEnd, const float4 *__restrict__ illum,
                                                          float4 *__restrict__ prevIllum, float4 *__restrict__ prevFast,
                                                          float *__restrict__ histLen, float *__restrict__ prevHistLen)
{
    PIXEL_GUARD(W, rowBegin, rowEnd)
    const float4 v = __ldg(illum + pix);
    prevIllum[pix] = v; prevFast[pix] = v; histLen[pix] = 0.0f; prevHistLen[pix] = 0.0f;
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
VPT_DEV f4 sampleBilinearCustom4(const float4 *tex, f2 uv, int W, int H, f4 cw)
{
    f2 f; int x0, y0; bilinearSetup(uv, W, H, f, x0, y0);
    f2 w1 = f, w0 = {1.0f - f.x, 1.0f - f.y};
    const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
    const float ws[4] = {w0.x * w0.y * cw.x, w1.x * w0.y * cw.y, w0.x * w1.y * cw.z, w1.x * w1.y * cw.w};
    f4 out = F4(0.0f); float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        f4 v = ld4(tex, W, H, xs[i], ys[i]);
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
VPT_DEV f4 sampleBicubic12(const float4 *tex, f2 uv, int W, int H)
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
    for (int i = 0; i < 12; ++i) { sum += ws[i]; out += ld4(tex, W, H, xs[i], ys[i]) * ws[i]; }
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
};
__global__ void __launch_bounds__(kBX *kBY) temporalKernel(const __grid_constant__ TemporalArgs a)
{
    const int W = a.W, H = a.H;
    PIXEL_GUARD(W, a.rowBegin, a.rowEnd)
    const float z = __ldg(a.depth + pix);
    if (z > a.denoisingRange) return;
    const Cam cam = loadCam(a.cam), prevCam = loadCam(a.prevCam);
    const quat prevToCur = rotationBetween(prevCam.dir, cam.dir);
    const f2 pixelUv = {(float(x) + 0.5f) * (1.0f / (float)W), (float(y) + 0.5f) * (1.0f / (float)H)};
    const f3 n = xyz(__ldg(a.normalRough + pix));
    const f2 curUV = {(float(x) + 0.5f) * cam.invResX, (float(y) + 0.5f) * cam.invResY};
    const f3 viewVec = uvToWorldDirection(cam, curUV);
    const f3 worldPos = worldPosFromPixel(cam, x, y, z);
    const f3 V = -normalize(viewVec);
    const float NoV = fabsf(dot(n, V));
    const f3 prevWorldPos = worldPos;
    const f2 prevUV = worldDirectionToUV(prevCam, normalize(prevWorldPos - prevCam.pos));
    const f3 illum = xyz(__ldg(a.illum + pix));
    f3 nAvg = n;
#pragma unroll
    for (int i = -1; i <= 1; ++i)
#pragma unroll
        for (int j = -1; j <= 1; ++j)
        {
            if (i == 0 && j == 0) continue;
            nAvg += xyz(ld4(a.normalRough, W, H, x + i, y + j));
        }
    nAvg /= 9.0f;
    const float m1 = luminance(illum), m2 = m1 * m1;
    const f3 camDelta = prevCam.pos - cam.pos;
    const f2 rect = {(float)W, (float)H};
    const float par1 = parallaxInPixels(prevWorldPos + camDelta, pixelUv, prevCam, rect);
    const float par2 = parallaxInPixels(prevWorldPos - camDelta, prevUV, cam, rect);
    const float parMax = fmaxr(par1, par2);
    const float thrBonus = a.disocclusionThreshold + (1.5f / H);
    const float thrAltBonus = a.disocclusionThresholdAlternate + (1.5f / H);
    const float disThr = lerpf(thrBonus, thrAltBonus, 0.0f);

    const f3 curNormalAvg = normalize(nAvg);
    const float estPrevDepth = length(prevWorldPos - prevCam.pos);
    const f2 prevPixF = {prevUV.x * W, prevUV.y * H};
    const int bx = (int)floorf(prevPixF.x - 0.5f), by = (int)floorf(prevPixF.y - 0.5f);
    const float pixelSize = (cam.tanHalfFovX / (cam.resX / 2)) * z;
    const float frustumSize = pixelSize * (float)min(W, H);
    const float slopeScale = 1.0f / lerpf(lerpf(0.05f, 1.0f, NoV), 1.0f, saturate(parMax / 30.0f));
    float thr[4];
    {
        const float base = saturate(disThr * slopeScale) * frustumSize;
        const int px0 = bx, py0 = by, px1 = bx + 1, py1 = by + 1;
        float rx0 = (px0 >= 0) ? 1.0f : 0.0f, ry0 = (py0 >= 0) ? 1.0f : 0.0f, rx1 = (px1 >= 0) ? 1.0f : 0.0f, ry1 = (py1 >= 0) ? 1.0f : 0.0f;
        rx0 *= (px0 < W) ? 1.0f : 0.0f; ry0 *= (py0 < H) ? 1.0f : 0.0f; rx1 *= (px1 < W) ? 1.0f : 0.0f; ry1 *= (py1 < H) ? 1.0f : 0.0f;
        const float inScreen[4] = {rx0 * ry0, rx1 * ry0, rx0 * ry1, rx1 * ry1};
#pragma unroll
        for (int i = 0; i < 4; ++i) { thr[i] = base * inScreen[i]; thr[i] -= 1e-6f; }
    }
    const int bic[4][2][2] = {{{0, -1}, {-1, 0}}, {{1, -1}, {2, 0}}, {{-1, 1}, {0, 2}}, {{2, 1}, {1, 2}}};
    const int bil[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
    float bicubicValid = 1.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
        {
            float pz = ld1(a.prevDepth, W, H, bx + bic[i][j][0], by + bic[i][j][1]);
            bicubicValid *= fabsf(pz - estPrevDepth) > thr[i] ? 0.0f : 1.0f;
        }
    float tv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        float pz = ld1(a.prevDepth, W, H, bx + bil[i][0], by + bil[i][1]);
        float v = fabsf(pz - estPrevDepth) > thr[i] ? 0.0f : 1.0f;
        bicubicValid *= v; tv[i] = v;
    }
    f4 tapsValid = {tv[0], tv[1], tv[2], tv[3]};
    const f3 prevNFlat = normalize(sampleSmoothStep3(a.prevNormalRough, prevUV, W, H));
    const f3 prevNRot = normalize(qrotate(prevToCur, prevNFlat));
    if (dot(curNormalAvg, prevNRot) < 0.0f) { tapsValid = F4(0.0f); bicubicValid = 0.0f; }
    const bool useBicubic = bicubicValid > 0;
    f4 prevIllum; f3 prevFast;
    if (useBicubic)
    {
        prevIllum = sampleBicubic12(a.prevIllum, prevUV, W, H);
        prevFast = xyz(sampleBicubic12(a.prevFast, prevUV, W, H));
    }
    else
    {
        prevIllum = sampleBilinearCustom4(a.prevIllum, prevUV, W, H, tapsValid);
        prevFast = xyz(sampleBilinearCustom4(a.prevFast, prevUV, W, H, tapsValid));
    }
    prevIllum = max4f(prevIllum, F4(0.0f));
    prevFast = max3f(prevFast, F3(0.0f));
    float reprojFound = (bicubicValid > 0.0f) ? 2.0f : 1.0f;
    const f4 bw = bilinearWeight(prevUV, W, H);
    float footprintQuality = (bicubicValid > 0) ? 1.0f : dot4(bw, F4(1.0f));
    float historyLength;
    if (dot4(tapsValid, F4(1.0f)) == 0.0f) { reprojFound = 0.0f; footprintQuality = 0.0f; historyLength = 0.0f; }
    else historyLength = sampleBilinearCustom1(a.prevHistLen, prevUV, W, H, tapsValid);

    historyLength = historyLength + 1.0f;
    const f3 Vprev = normalize(prevWorldPos - prevCam.pos);
    const float NoVprev = fabsf(dot(n, Vprev));
    float sizeQuality = (NoVprev + 1e-3f) / (NoV + 1e-3f);
    sizeQuality *= sizeQuality; sizeQuality *= sizeQuality;
    footprintQuality *= lerpf(0.1f, 1.0f, saturate(sizeQuality));
    if (footprintQuality < 1.0f) { historyLength *= sqrtf(footprintQuality); historyLength = fmaxr(historyLength, 1.0f); }
    historyLength = fminr(historyLength, a.maxAccum);
    const float alpha = (reprojFound > 0) ? fmaxr(1.0f / (a.maxAccum + 1.0f), 1.0f / historyLength) : 1.0f;
    const float alphaFast = (reprojFound > 0) ? fmaxr(1.0f / (a.maxFastAccum + 1.0f), 1.0f / historyLength) : 1.0f;
    a.ping[pix] = toFloat4(lerp4(prevIllum, F4(illum, m2), alpha));
    a.pong[pix] = toFloat4(F4(lerp3(prevFast, illum, alphaFast), 0.0f));
    a.histLen[pix] = historyLength;
}

// ------------------------------------------------------------------------------------------------ history fix
__global__ void __launch_bounds__(kBX *kBY) historyFixKernel(int W, int H, int rowBegin, int rowEnd, VptCamera camIn,
                                                              const float *__restrict__ depth, const float *__restrict__ material,
                                                              const float4 *__restrict__ normalRough, const float *__restrict__ histLen,
                                                              const float4 *__restrict__ ping, float4 *__restrict__ pong)
{
    PIXEL_GUARD(W, rowBegin, rowEnd)
    const float z = __ldg(depth + pix);
    const float hl = __ldg(histLen + pix);
    if (z > kDenoisingRange || hl > 4.0f) return;
    const Cam cam = loadCam(camIn);
    const float cMat = matU16(material, W, H, x, y);
    const f3 cn = xyz(__ldg(normalRough + pix));
    const f3 cpos = worldPosFromPixel(cam, x, y, z);
    const float depthThr = 0.003f * z;
    f4 sum = F4(__ldg(ping + pix));
    float wsum = 1.0f;
    const float r = exp2f(4.0f - hl) + 1.0f;
    for (int j = -2; j <= 2; ++j)
        for (int i = -2; i <= 2; ++i)
        {
            const int dx = (int)(i * r), dy = (int)(j * r);
            const int sx = x + dx, sy = y + dy;
            const bool inside = sx >= 0 && sy >= 0 && sx < W && sy < H;
            if (i == 0 && j == 0) continue;
            const float sMat = matU16(material, W, H, sx, sy);
            const f3 sn = xyz(ld4(normalRough, W, H, sx, sy));
            const float sz = ld1(depth, W, H, sx, sy);
            const f3 spos = worldPosFromPixel(cam, sx, sy, sz);
            float w = planeDistWeightAtrous(cpos, cn, spos, depthThr);
            w *= powf(fmaxr(0.01f, dot(cn, sn)), fmaxr(8.0f, 0.01f));
            w = inside ? w : 0;
            w *= (sMat == cMat) ? 1.0f : 0.0f;
            if (w > 1e-4f) { sum += ld4(ping, W, H, sx, sy) * w; wsum += w; }
        }
    pong[pix] = toFloat4(sum / wsum);
}

// ------------------------------------------------------------------------------------------------ history clamping
__global__ void __launch_bounds__(kBX *kBY) historyClampKernel(int W, int H, int rowBegin, int rowEnd, const float *__restrict__ depth,
                                                                const float4 *__restrict__ illum, const float4 *__restrict__ ping,
                                                                const float4 *__restrict__ pong, const float *__restrict__ histLen,
                                                                float4 *__restrict__ prevIllum, float4 *__restrict__ prevFast,
                                                                float *__restrict__ prevHistLen)
{
    PIXEL_GUARD(W, rowBegin, rowEnd)
    if (__ldg(depth + pix) > kDenoisingRange) return;
    const float hl = __ldg(histLen + pix);
    f3 rM1 = F3(0.0f), rM2 = F3(0.0f), nM1 = F3(0.0f); float nM2 = 0.0f;
    for (int dx = -2; dx <= 2; ++dx)
        for (int dy = -2; dy <= 2; ++dy)
        {
            const f3 s = rgbToYCoCg(xyz(ld4(pong, W, H, x + dx, y + dy)));
            rM1 += s; rM2 += s * s;
            const f3 nz = xyz(ld4(illum, W, H, x + dx, y + dy));
            const float nl = luminance(nz);
            nM1 += nz; nM2 += nl * nl;
        }
    rM1 /= 25.0f; rM2 /= 25.0f; nM1 /= 25.0f; nM2 /= 25.0f;
    const f3 sigma = sqrt3(max3f(F3(0.0f), rM2 - rM1 * rM1));
    f3 cmin = rM1 - 2.0f * sigma, cmax = rM1 + 2.0f * sigma;
    const f3 centerY = rgbToYCoCg(xyz(__ldg(pong + pix)));
    cmin = (cmin.x < centerY.x) ? cmin : centerY;
    cmax = (cmax.x > centerY.x) ? cmax : centerY;
    const f4 acc = F4(__ldg(ping + pix));
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
    const f3 noisyC = xyz(__ldg(illum + pix));
    f3 d3 = lerp3(xyz(outD), noisyC, reset), r3 = lerp3(xyz(outR), noisyC, reset);
    outD = F4(d3, outD.w); outR = F4(r3, outR.w);
    const float outL = luminance(xyz(outD));
    outD.w += (outL * outL - diffL * diffL);
    outD.w = fmaxr(0.0f, outD.w);
    prevIllum[pix] = toFloat4(outD);
    prevFast[pix] = toFloat4(outR);
    prevHistLen[pix] = hl;
}

// ------------------------------------------------------------------------------------------------ à-trous
struct AtrousArgs
{
    int W, H, rowBegin, rowEnd;
    VptCamera cam;
    float phiLuminance, depthThreshold, lobeAngleFraction;
    unsigned frameIndex, step;
    const float4 *in, *normalRough;
    const float *material, *depth, *histLen;
    float4 *out;
};
__global__ void __launch_bounds__(kBX *kBY) atrousFirstKernel(const __grid_constant__ AtrousArgs a)
{
    const int W = a.W, H = a.H;
    PIXEL_GUARD(W, a.rowBegin, a.rowEnd)
    const float z = __ldg(a.depth + pix);
    if (z > 500000.0f) return;
    const Cam cam = loadCam(a.cam);
    const f3 cn = xyz(__ldg(a.normalRough + pix));
    const f3 cpos = worldPosFromPixel(cam, x, y, z);
    const float cMat = __ldg(a.material + pix);
    const float hl = __ldg(a.histLen + pix);
    if (hl >= 3.0f)
    {
        f4 vsum = F4(0.0f);
        const float kern[4] = {1.0f / 4.0f, 1.0f / 8.0f, 1.0f / 8.0f, 1.0f / 16.0f};
#pragma unroll
        for (int dx = -1; dx <= 1; dx++)
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
                vsum += ld4(a.in, W, H, x + dx, y + dy) * kern[abs(dx) * 2 + abs(dy)];
        const float v1 = luminance(xyz(vsum));
        const float cVar = fmaxr(0.0f, vsum.w - v1 * v1);
        const float cLum = luminance(xyz(__ldg(a.in + pix)));
        const float phiInv = 1.0f / fmaxr(1.0e-4f, a.phiLuminance * sqrtf(cVar));
        const float nParam = normalWeightParam2(1.0f, a.lobeAngleFraction);
        float sumW = 0.0f; f4 sum = F4(0.0f);
        const float k3[2] = {0.44198f, 0.27901f};
        const float depthThr = a.depthThreshold * z;
        for (int cx = -1; cx <= 1; cx++)
            for (int cy = -1; cy <= 1; cy++)
            {
                const int sx = x + cx, sy = y + cy;
                const bool center = (cx == 0 && cy == 0);
                const bool inside = sx >= 0 && sy >= 0 && sx < W && sy < H;
                const float kernel = inside ? k3[abs(cx)] * k3[abs(cy)] : 0.0f;
                const int qx = clampi(sx, 0, W - 1), qy = clampi(sy, 0, H - 1);
                const size_t sp = (size_t)qy * W + qx;
                const f3 sn = xyz(__ldg(a.normalRough + sp));
                const f3 spos = worldPosFromPixel(cam, qx, qy, __ldg(a.depth + sp));
                const float sMat = __ldg(a.material + sp);
                float geomW = planeDistWeightAtrous(cpos, cn, spos, depthThr);
                geomW *= kernel;
                const float normalW = nonExpWeight(acosApprox(dot(cn, sn)), nParam, 0.0f);
                const f4 sv = F4(__ldg(a.in + sp));
                const float sLum = luminance(xyz(sv));
                const float lumW = fabsf(cLum - sLum) * phiInv;
                float w = geomW * normalW * expf(-lumW);
                w = center ? kernel : w;
                w *= (sMat == cMat) ? 1.0f : 0.0f;
                sumW += w;
                sum += w * sv;
            }
        sumW = fmaxr(sumW, 1e-6f);
        sum = sum / sumW;
        const float m1 = luminance(xyz(sum));
        a.out[pix] = make_float4(sum.x, sum.y, sum.z, fmaxr(0.0f, sum.w - m1 * m1));
    }
    else
    {
        float sumW = 0.0f; f3 sumI = F3(0.0f); float s1 = 0.0f, s2 = 0.0f;
        const float nParam = normalWeightParam2(1.0f, a.lobeAngleFraction);
        for (int cx = -2; cx <= 2; cx++)
            for (int cy = -2; cy <= 2; cy++)
            {
                const int qx = clampi(x + cx, 0, W - 1), qy = clampi(y + cy, 0, H - 1);
                const size_t sp = (size_t)qy * W + qx;
                const f3 sn = xyz(__ldg(a.normalRough + sp));
                const float sMat = __ldg(a.material + sp);
                const float normalW = nonExpWeight(acosApprox(dot(cn, sn)), nParam, 0.0f);
                const f4 sv = F4(__ldg(a.in + sp));
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
        a.out[pix] = make_float4(sumI.x, sumI.y, sumI.z, var);
    }
}

__global__ void __launch_bounds__(kBX *kBY) atrousKernel(const __grid_constant__ AtrousArgs a)
{
    const int W = a.W, H = a.H;
    PIXEL_GUARD(W, a.rowBegin, a.rowEnd)
    const float z = __ldg(a.depth + pix);
    if (z > 500000.0f) return;
    const Cam cam = loadCam(a.cam);
    const float cMat = matU16(a.material, W, H, x, y);
    const f3 cn = xyz(__ldg(a.normalRough + pix));
    const f3 cpos = worldPosFromPixel(cam, x, y, z);
    const float hl = __ldg(a.histLen + pix);
    const unsigned stepSize = a.step;
    float lobeFrac = a.lobeAngleFraction / sqrtf((float)stepSize);
    lobeFrac = lerpf(0.99f, lobeFrac, saturate(hl / 5.0f));
    const f4 cv = F4(__ldg(a.in + pix));
    const float cLum = luminance(xyz(cv));
    const float phiInv = 1.0f / fmaxr(1.0e-4f, a.phiLuminance * sqrtf(cv.w));
    const float nParam = normalWeightParam2(1.0f, lobeFrac);
    float sumW = 0.44198f * 0.44198f;
    f4 sum = cv * f4{sumW, sumW, sumW, sumW * sumW};
    const float k3[2] = {0.44198f, 0.27901f};
    const float depthThr = a.depthThreshold * z;
    int offx = 0, offy = 0;
    if (stepSize > 4)
    {
        uint32_t zorder = seqExplode((uint32_t)x) | (seqExplode((uint32_t)y) << 1);
        uint32_t seed = seqHash(a.frameIndex + 0x035F9F29u);
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
            const float sMat = matU16(a.material, W, H, sx, sy);
            const f3 sn = xyz(ld4(a.normalRough, W, H, sx, sy));
            const float sz = ld1(a.depth, W, H, sx, sy);
            const f3 spos = worldPosFromPixel(cam, sx, sy, sz);
            float geomW = planeDistWeightAtrous(cpos, cn, spos, depthThr);
            geomW *= kernel;
            geomW *= (inside && sz < 500000.0f) ? 1.0f : 0.0f;
            const float normalW = nonExpWeight(acosApprox(dot(cn, sn)), nParam, 0.0f);
            float w = geomW * normalW;
            w *= (sMat == cMat) ? 1.0f : 0.0f;
            if (w > 1e-4f)
            {
                const f4 sv = ld4(a.in, W, H, sx, sy);
                const float sLum = luminance(xyz(sv));
                const float lumW = fabsf(cLum - sLum) * phiInv;
                w *= expf(-lumW);
                sumW += w;
                sum += f4{w, w, w, w * w} * sv;
            }
        }
    a.out[pix] = toFloat4(sum / f4{sumW, sumW, sumW, sumW * sumW});
}

// ------------------------------------------------------------------------------------------------ launchers
static dim3 gridFor(const DenoiseLaunch &d)
{
    return dim3((d.width + kBX - 1) / kBX, (d.rowEnd - d.rowBegin + kBY - 1) / kBY);
}
static const dim3 kBlock(kBX, kBY);

cudaError_t launchFirefly(const DenoiseLaunch &d, FireflyPatch *patches, int *patchCount, int maxPatches)
{
    cudaError_t e = cudaMemsetAsync(patchCount, 0, sizeof(int), d.stream);
    if (e != cudaSuccess) return e;
    const int tiles = ((d.width + 7) / 8) * ((d.rowEnd - d.rowBegin + 3) / 4);
    fireflyDetectKernel<<<(tiles + 7) / 8, 256, 0, d.stream>>>(d.width, d.height, d.rowBegin, d.rowEnd, d.b.illumination, d.b.cur.normalRoughness,
                                                               d.b.cur.depth, d.b.cur.material, d.b.reservoirs, 80.0f, 5.0f, 0.8f, 0.02f,
                                                               d.p.phiLuminance, d.cam, patches, patchCount, maxPatches);
    fireflyApplyKernel<<<64, 256, 0, d.stream>>>(patches, patchCount, maxPatches, d.b.illumination, d.b.reservoirs);
    return cudaGetLastError();
}
cudaError_t launchCopySky(const DenoiseLaunch &d)
{
    copySkyKernel<<<gridFor(d), kBlock, 0, d.stream>>>(d.width, d.rowBegin, d.rowEnd, d.b.illumination, d.b.cur.depth, d.b.illumOutput);
    return cudaGetLastError();
}
cudaError_t launchFrame0Init(const DenoiseLaunch &d)
{
    frame0Kernel<<<gridFor(d), kBlock, 0, d.stream>>>(d.width, d.rowBegin, d.rowEnd, d.b.illumination, d.b.prevIllum, d.b.prevFastIllum,
                                                      d.b.historyLength, d.b.prevHistoryLength);
    return cudaGetLastError();
}
cudaError_t launchTemporal(const DenoiseLaunch &d)
{
    TemporalArgs a;
    a.W = d.width; a.H = d.height; a.rowBegin = d.rowBegin; a.rowEnd = d.rowEnd;
    a.cam = d.cam; a.prevCam = d.prevCam;
    a.denoisingRange = d.p.denoisingRange; a.disocclusionThreshold = d.p.disocclusionThreshold;
    a.disocclusionThresholdAlternate = d.p.disocclusionThresholdAlternate;
    a.maxAccum = d.p.maxAccumulatedFrameNum; a.maxFastAccum = d.p.maxFastAccumulatedFrameNum;
    a.depth = d.b.cur.depth; a.prevDepth = d.b.prev.depth; a.prevHistLen = d.b.prevHistoryLength;
    a.normalRough = d.b.cur.normalRoughness; a.prevNormalRough = d.b.prev.normalRoughness;
    a.illum = d.b.illumination; a.prevIllum = d.b.prevIllum; a.prevFast = d.b.prevFastIllum;
    a.ping = d.b.ping; a.pong = d.b.pong; a.histLen = d.b.historyLength;
    temporalKernel<<<gridFor(d), kBlock, 0, d.stream>>>(a);
    return cudaGetLastError();
}
cudaError_t launchHistoryFix(const DenoiseLaunch &d)
{
    historyFixKernel<<<gridFor(d), kBlock, 0, d.stream>>>(d.width, d.height, d.rowBegin, d.rowEnd, d.cam, d.b.cur.depth, d.b.cur.material,
                                                          d.b.cur.normalRoughness, d.b.historyLength, d.b.ping, d.b.pong);
    return cudaGetLastError();
}
cudaError_t launchHistoryClamping(const DenoiseLaunch &d)
{
    historyClampKernel<<<gridFor(d), kBlock, 0, d.stream>>>(d.width, d.height, d.rowBegin, d.rowEnd, d.b.cur.depth, d.b.illumination, d.b.ping,
                                                            d.b.pong, d.b.historyLength, d.b.prevIllum, d.b.prevFastIllum, d.b.prevHistoryLength);
    return cudaGetLastError();
}
static AtrousArgs atrousArgs(const DenoiseLaunch &d, const float4 *in, float4 *out, unsigned frameIndex, unsigned step)
{
    AtrousArgs a;
    a.W = d.width; a.H = d.height; a.rowBegin = d.rowBegin; a.rowEnd = d.rowEnd;
    a.cam = d.cam;
    a.phiLuminance = d.p.phiLuminance; a.depthThreshold = d.p.depthThreshold; a.lobeAngleFraction = d.p.lobeAngleFraction;
    a.frameIndex = frameIndex; a.step = step;
    a.in = in; a.normalRough = d.b.cur.normalRoughness; a.material = d.b.cur.material; a.depth = d.b.cur.depth;
    a.histLen = d.b.historyLength; a.out = out;
    return a;
}
cudaError_t launchAtrousSmem(const DenoiseLaunch &d, const float4 *in, float4 *out)
{
    atrousFirstKernel<<<gridFor(d), kBlock, 0, d.stream>>>(atrousArgs(d, in, out, 0, 1));
    return cudaGetLastError();
}
cudaError_t launchAtrous(const DenoiseLaunch &d, const float4 *in, float4 *out, unsigned frameIndex, unsigned step)
{
    atrousKernel<<<gridFor(d), kBlock, 0, d.stream>>>(atrousArgs(d, in, out, frameIndex, step));
    return cudaGetLastError();
}
cudaError_t launchCompositeNonSky(const DenoiseLaunch &d, const float4 *finalBuf)
{
    compositeKernel<<<gridFor(d), kBlock, 0, d.stream>>>(d.width, d.rowBegin, d.rowEnd, finalBuf, d.b.cur.depth, d.b.cur.albedo, d.b.illumOutput);
    return cudaGetLastError();
}

} // namespace vpt
