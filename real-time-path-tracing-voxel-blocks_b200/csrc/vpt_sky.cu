// Sky / sun radiance maps on the device: replaces SkyModel::update
// (/root/reference/renderer/sky/Sky.cu:355-396: Sky -> thrust::reduce -> SkyLowerHemisphere -> SkySun kernels, the
// Hosek-Wilkie evaluation GetSkyRadiance :133-173 and GetSunRadiance :175-257, updateSkyState :52-79 on the host).
// The coefficient tables (SkyData.h) are data uploaded by the caller. Not on the per-frame path (runs when the sky
// parameters change); compiled without fast-math options and without FMA contraction so the maps agree with the CPU
// restatement to libm precision.
#include "vpt_kernels.h"
#include "vpt_math.cuh"

namespace vpt {

struct SkyConsts
{
    float configs[90], radiances[10];
    float sunDir[3];
    float brightness;
};

__device__ __forceinline__ f3 spectrumToXyz(int c)
{
    const float X[] = {2.372527e-02f, 1.955480e+00f, 1.074553e+01f, 5.056697e+00f, 4.698190e+00f, 2.391135e+01f, 3.798705e+01f, 1.929414e+01f, 2.970610e+00f, 2.092986e-01f};
    const float Y[] = {6.813859e-04f, 6.771017e-02f, 1.171193e+00f, 6.997765e+00f, 2.666710e+01f, 3.758372e+01f, 2.503930e+01f, 8.150395e+00f, 1.098635e+00f, 7.563256e-02f};
    const float Z[] = {1.119121e-01f, 9.441195e+00f, 5.597921e+01f, 3.589996e+01f, 5.070894e+00f, 3.523189e-01f, 3.422707e-02f, 2.539118e-03f, 7.836666e-06f, 0.000000e+00f};
    const float integral = 106.856895f;
    return F3(X[c], Y[c], Z[c]) / integral;
}
__device__ __forceinline__ f3 xyzToRgbSrgb(f3 v)
{
    const mat3 m = {3.2404542f, -0.9692660f, 0.0556434f, -1.5371385f, 1.8760108f, -0.2040259f, -0.4985314f, 0.0415560f, 1.0572252f};
    return mul(m, v);
}
__device__ f3 skyRadiance(const SkyConsts &k, f3 raydir, f3 sunDir)
{
    const float theta = acosf(raydir.y);
    const float gamma = acosf(clampf(dot(raydir, sunDir), -1.0f, 1.0f));
    f3 xyzc = F3(0.0f);
#pragma unroll 1
    for (int c = 0; c < 10; ++c)
    {
        const float *cf = k.configs + c * 9;
        const float expM = expf(cf[4] * gamma);
        const float rayM = cosf(gamma) * cosf(gamma);
        const float mieM = (1.0f + cosf(gamma) * cosf(gamma)) / powf((1.0f + cf[8] * cf[8] - 2.0f * cf[8] * cosf(gamma)), 1.5f);
        const float zenith = sqrtf(cosf(theta));
        const float radianceInternal = (1.0f + cf[0] * expf(cf[1] / (cosf(theta) + 0.01f))) *
                                       (cf[2] + cf[3] * expM + cf[5] * rayM + cf[6] * mieM + cf[7] * zenith);
        const float radiance = radianceInternal * k.radiances[c];
        xyzc += radiance * spectrumToXyz(c);
    }
    return xyzToRgbSrgb(xyzc);
}
__device__ f3 sunRadiance(const float *__restrict__ solar, const float *__restrict__ limb, f3 raydir, f3 sunDir)
{
    const float gamma = acosf(clampf(dot(raydir, sunDir), -1.0f, 1.0f));
    const float elevation = (kPi / 2.0f) - acosf(sunDir.y);
    const float sunAngle = 0.51f;
    const float solarRadius = sunAngle * kPi / 180.0f / 2.0f;
    const float scale = 1.0f / ((sunAngle / 0.51f) * (sunAngle / 0.51f));
    f3 xyzc = F3(0.0f);
    const float solRadSin = sinf(solarRadius);
    const float ar2 = 1.0f / (solRadSin * solRadSin);
    const float singamma = sinf(gamma);
    float sc2 = 1.0f - ar2 * singamma * singamma;
    if (sc2 < 0.0f) sc2 = 0.0f;
    const float sampleCosine = sqrtf(sc2);
    if (sampleCosine == 0.0f) return F3(0.0f);
#pragma unroll 1
    for (int c = 0; c < 10; ++c)
    {
        const int pieces = 45, order = 4;
        int pos = (int)(powf(2.0f * elevation / kPi, 1.0f / 3.0f) * pieces);
        if (pos > 44) pos = 44;
        const float breakX = powf(((float)pos / (float)pieces), 3.0f) * (kPi * 0.5f);
        const float *coefs = solar + c * 180 + (order * (pos + 1) - 1);
        float res = 0.0f;
        const float x = elevation - breakX;
        float xExp = 1.0f;
        for (int i = 0; i < order; ++i) { res += xExp * *coefs--; xExp *= x; }
        float direct = res;
        const float *ld = limb + c * 6;
        const float dark = ld[0] + ld[1] * sampleCosine + ld[2] * powf(sampleCosine, 2.0f) + ld[3] * powf(sampleCosine, 3.0f) +
                           ld[4] * powf(sampleCosine, 4.0f) + ld[5] * powf(sampleCosine, 5.0f);
        direct *= dark * scale;
        xyzc += direct * spectrumToXyz(c);
    }
    return xyzToRgbSrgb(xyzc);
}

// Sky (Sky.cu:259-283): upper hemisphere of the equal-area sphere map, rows [H/2, H)
__global__ void skyUpperKernel(const __grid_constant__ SkyConsts k, float4 *__restrict__ sky, float *__restrict__ pdf, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, half = H / 2;
    if (x >= W || y >= half) return;
    const float u = ((float)x + 0.5f) / W, v = ((float)y + 0.5f) / half;
    const float r = sqrtf(1.0f - v * v), phi = kTwoPi * u;
    const f3 dir = {r * cosf(phi), v, r * sinf(phi)};
    const f3 sunDir = {k.sunDir[0], k.sunDir[1], k.sunDir[2]};
    f3 color = skyRadiance(k, dir, sunDir) * k.brightness;
    color = max3f(color, F3(0.0f));
    const size_t i = (size_t)W * (y + half) + x;
    sky[i] = make_float4(color.x, color.y, color.z, 0.0f);
    pdf[i] = luminance(color);
}
// SkyLowerHemisphere (Sky.cu:285-307): mist colour blended towards the horizon row
__global__ void skyLowerKernel(float4 *__restrict__ sky, float *__restrict__ pdf, int W, int H, float sumSkyPdf)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, half = H / 2;
    if (x >= W || y >= half) return;
    const float v = ((float)y + 0.5f) / half - 1.0f;
    const f3 mist = F3(sumSkyPdf / (W * H));
    const float w = clampf((v + 0.4f) * (1.0f / 0.5f), 0.0f, 1.0f);
    const f3 horizon = xyz(sky[(size_t)W * half + x]);
    const f3 color = mist + (w * w * (3.0f - 2.0f * w)) * (horizon - mist);
    const size_t i = (size_t)W * y + x;
    sky[i] = make_float4(color.x, color.y, color.z, 0.0f);
    pdf[i] = luminance(color);
}
// SkySun (Sky.cu:309-327): the solar disc over the equal-area cone map
__global__ void skySunKernel(const __grid_constant__ SkyConsts k, const float *__restrict__ solar, const float *__restrict__ limb,
                             float4 *__restrict__ sun, float *__restrict__ pdf, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float u = ((float)x + 0.5f) / W, v = ((float)y + 0.5f) / H;
    const f3 sunDir = {k.sunDir[0], k.sunDir[1], k.sunDir[2]};
    const f3 dir = equalAreaMapCone(sunDir, u, v, cosf(0.51f * kPi / 180.0f / 2.0f));
    f3 color = sunRadiance(solar, limb, dir, sunDir) * k.brightness;
    color = max3f(color, F3(0.0f));
    const size_t i = (size_t)W * y + x;
    sun[i] = make_float4(color.x, color.y, color.z, 0.0f);
    pdf[i] = luminance(color);
}

cudaError_t launchSkyUpper(const float *configs, const float *radiances, const float *sunDir, float brightness, float4 *sky, float *pdf, int W, int H,
                           cudaStream_t s)
{
    SkyConsts k;
    for (int i = 0; i < 90; ++i) k.configs[i] = configs[i];
    for (int i = 0; i < 10; ++i) k.radiances[i] = radiances[i];
    for (int i = 0; i < 3; ++i) k.sunDir[i] = sunDir[i];
    k.brightness = brightness;
    const dim3 b(32, 8), g((W + 31) / 32, (H / 2 + 7) / 8);
    skyUpperKernel<<<g, b, 0, s>>>(k, sky, pdf, W, H);
    return cudaGetLastError();
}
cudaError_t launchSkyLower(float4 *sky, float *pdf, int W, int H, float sumSkyPdf, cudaStream_t s)
{
    const dim3 b(32, 8), g((W + 31) / 32, (H / 2 + 7) / 8);
    skyLowerKernel<<<g, b, 0, s>>>(sky, pdf, W, H, sumSkyPdf);
    return cudaGetLastError();
}
cudaError_t launchSkySun(const float *sunDir, float brightness, const float *solar, const float *limb, float4 *sun, float *pdf, int W, int H, cudaStream_t s)
{
    SkyConsts k = {};
    for (int i = 0; i < 3; ++i) k.sunDir[i] = sunDir[i];
    k.brightness = brightness;
    const dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    skySunKernel<<<g, b, 0, s>>>(k, solar, limb, sun, pdf, W, H);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- output stage
// FilmicToneMapping (renderer/postprocessing/FilmicToneMapping.h:12-117) with the manual exposure, then
// OfflineBackend::writeFrameBufferToPNG's conversion (renderer/core/OfflineBackend.cpp:191-221).
__device__ __forceinline__ f3 clamp01(f3 v) { return {clampf(v.x, 0.0f, 1.0f), clampf(v.y, 0.0f, 1.0f), clampf(v.z, 0.0f, 1.0f)}; }
__device__ __forceinline__ f3 acesFilm(f3 x)
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return clamp01(x * (a * x + F3(b)) / (x * (c * x + F3(d)) + F3(e)));
}
__device__ __forceinline__ f3 uncharted2(f3 x)
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, Fc = 0.30f;
    return ((x * (A * x + F3(C * B)) + F3(D * E)) / (x * (A * x + F3(B)) + F3(D * Fc))) - F3(E / Fc);
}
__device__ __forceinline__ float linearToSrgb(float c) { return (c <= 0.0031308f) ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f; }
__global__ void tonemapKernel(const float4 *__restrict__ hdr, int W, int H, VptToneMappingParams p, uint8_t *__restrict__ rgb8, float4 *__restrict__ ldr)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float4 in = __ldg(hdr + (size_t)y * W + x);
    f3 color = F3(in.x, in.y, in.z) * p.manualExposure;
    f3 tm;
    if (p.curve == 1)
    {
        const f3 whiteScale = F3(1.0f) / uncharted2(F3(p.whitePoint));
        tm = uncharted2(color * 2.0f) * whiteScale;
    }
    else if (p.curve == 2)
    {
        const f3 numerator = color * (F3(1.0f) + (color / (p.whitePoint * p.whitePoint)));
        tm = numerator / (F3(1.0f) + color);
    }
    else tm = acesFilm(color);
    tm = clamp01(tm);
    tm = {powf(tm.x, p.contrast), powf(tm.y, p.contrast), powf(tm.z, p.contrast)};
    const float lum = dot(tm, F3(0.2126f, 0.7152f, 0.0722f));
    tm = lerp3(F3(lum), tm, p.saturation);
    tm = clamp01(tm * p.gain + F3(p.lift));
    tm = {linearToSrgb(tm.x), linearToSrgb(tm.y), linearToSrgb(tm.z)};
    if (ldr) ldr[(size_t)y * W + x] = make_float4(tm.x, tm.y, tm.z, 1.0f);
    if (rgb8)
    {
        uint8_t *q = rgb8 + ((size_t)(H - 1 - y) * W + x) * 3;
        q[0] = (uint8_t)(fminf(1.0f, fmaxf(0.0f, tm.x)) * 255.0f);
        q[1] = (uint8_t)(fminf(1.0f, fmaxf(0.0f, tm.y)) * 255.0f);
        q[2] = (uint8_t)(fminf(1.0f, fmaxf(0.0f, tm.z)) * 255.0f);
    }
}
cudaError_t launchTonemap(const float4 *hdr, int W, int H, const VptToneMappingParams &p, uint8_t *rgb8, float4 *ldr, cudaStream_t s)
{
    const dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    tonemapKernel<<<g, b, 0, s>>>(hdr, W, H, p, rgb8, ldr);
    return cudaGetLastError();
}

} // namespace vpt
