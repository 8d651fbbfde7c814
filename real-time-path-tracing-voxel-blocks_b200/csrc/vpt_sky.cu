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

} // namespace vpt
