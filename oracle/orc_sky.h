// ORACLE — test infrastructure only (see orc_math.h header). CPU restatement of the reference's sky generator:
//
//   SkyModel::update           /root/reference/renderer/sky/Sky.cu:355-396 (sun direction from timeOfDay / axis angles,
//                              Sky -> reduce -> SkyLowerHemisphere -> SkySun, pdf planes for the alias tables)
//   updateSkyState             Sky.cu:52-79, getFittingData/2 :18-50 (quintic Bezier in solarElevation^(1/3))
//   GetSkyRadiance             Sky.cu:133-173 (Hosek-Wilkie, 10 spectral channels -> XYZ -> sRGB)
//   GetSunRadiance             Sky.cu:175-257 (solar disc radiance + limb darkening)
//   Sky / SkyLowerHemisphere / SkySun kernels  Sky.cu:259-327
//   SpectrumToXyz              Sky.cu:81-131, XyzToRgbSrgb renderer/util/ColorSpace.h:18-29
//   EqualAreaHemisphereMap     renderer/shaders/LinearMath.h:1841-1848, smoothstep3f :1033, rotate3f :1368
// The coefficient tables (SkyData.h) are DATA passed in by the caller (data/sky_tables.bin).
// M_PI is a float literal in the reference (LinearMath.h:17); double-typed literals (1.5, 2.0, 1.0/3.0) promote as in C.
// One documented departure: thrust::reduce's summation order is unspecified; the upper-hemisphere luminance sum is
// taken in double and rounded once (the CUDA path does the same on the host).
#pragma once
#include "orc_math.h"
#include <vector>

namespace orc {

struct SkyTables { const float *skyDataSets, *skyDataSetsRad, *solar, *limb; };
inline SkyTables skyTablesFrom(const float *blob) { return {blob, blob + 540, blob + 600, blob + 2400}; }

inline float skyFit(const float *m, float s, int i)
{
    return (powf(1.0f - s, 5.0f) * m[i] + 5.0f * powf(1.0f - s, 4.0f) * s * m[i + 9] + 10.0f * powf(1.0f - s, 3.0f) * powf(s, 2.0f) * m[i + 18] +
            10.0f * powf(1.0f - s, 2.0f) * powf(s, 3.0f) * m[i + 27] + 5.0f * (1.0f - s) * powf(s, 4.0f) * m[i + 36] + powf(s, 5.0f) * m[i + 45]);
}
inline float skyFit2(const float *m, float s)
{
    return (powf(1.0f - s, 5.0f) * m[0] + 5.0f * powf(1.0f - s, 4.0f) * s * m[1] + 10.0f * powf(1.0f - s, 3.0f) * powf(s, 2.0f) * m[2] +
            10.0f * powf(1.0f - s, 2.0f) * powf(s, 3.0f) * m[3] + 5.0f * (1.0f - s) * powf(s, 4.0f) * m[4] + powf(s, 5.0f) * m[5]);
}
struct SkyState { float configs[90], radiances[10]; };
inline SkyState skyUpdateState(const SkyTables &t, f3 sunDir)
{
    SkyState st;
    const float elevation = (kPi / 2.0f) - (float)acos((double)sunDir.y); // acos(float) -> double overload in host code
    const float solarElevation = powf(elevation / (kPi / 2.0f), (1.0f / 3.0f));
    for (int c = 0; c < 10; ++c)
    {
        for (int i = 0; i < 9; ++i) st.configs[c * 9 + i] = skyFit(t.skyDataSets + c * 54, solarElevation, i);
        st.radiances[c] = skyFit2(t.skyDataSetsRad + c * 6, solarElevation);
    }
    return st;
}
inline f3 spectrumToXyz(int c)
{
    static const float X[] = {2.372527e-02f, 1.955480e+00f, 1.074553e+01f, 5.056697e+00f, 4.698190e+00f, 2.391135e+01f, 3.798705e+01f, 1.929414e+01f, 2.970610e+00f, 2.092986e-01f};
    static const float Y[] = {6.813859e-04f, 6.771017e-02f, 1.171193e+00f, 6.997765e+00f, 2.666710e+01f, 3.758372e+01f, 2.503930e+01f, 8.150395e+00f, 1.098635e+00f, 7.563256e-02f};
    static const float Z[] = {1.119121e-01f, 9.441195e+00f, 5.597921e+01f, 3.589996e+01f, 5.070894e+00f, 3.523189e-01f, 3.422707e-02f, 2.539118e-03f, 7.836666e-06f, 0.000000e+00f};
    const float integral = 106.856895f;
    return F3(X[c], Y[c], Z[c]) / integral;
}
inline f3 xyzToRgbSrgb(f3 v)
{
    const mat3 m = {3.2404542f, -0.9692660f, 0.0556434f, -1.5371385f, 1.8760108f, -0.2040259f, -0.4985314f, 0.0415560f, 1.0572252f};
    return mul(m, v);
}
inline f3 skyRadiance(const SkyState &st, f3 raydir, f3 sunDir)
{
    const float theta = acosf(raydir.y);
    const float gamma = acosf(clampf(dot(raydir, sunDir), -1.0f, 1.0f));
    f3 xyz = F3(0.0f);
    for (int c = 0; c < 10; ++c)
    {
        const float *cf = st.configs + c * 9;
        const float expM = expf(cf[4] * gamma);
        const float rayM = cosf(gamma) * cosf(gamma);
        const float mieM = (1.0f + cosf(gamma) * cosf(gamma)) / powf((1.0f + cf[8] * cf[8] - 2.0f * cf[8] * cosf(gamma)), 1.5f);
        const float zenith = sqrtf(cosf(theta));
        const float radianceInternal = (1.0f + cf[0] * expf(cf[1] / (cosf(theta) + 0.01f))) *
                                       (cf[2] + cf[3] * expM + cf[5] * rayM + cf[6] * mieM + cf[7] * zenith);
        const float radiance = radianceInternal * st.radiances[c];
        xyz += radiance * spectrumToXyz(c);
    }
    return xyzToRgbSrgb(xyz);
}
inline f3 sunRadiance(const SkyTables &t, f3 raydir, f3 sunDir)
{
    const float gamma = acosf(clampf(dot(raydir, sunDir), -1.0f, 1.0f));
    const float elevation = (kPi / 2.0f) - acosf(sunDir.y);
    const float sunAngle = 0.51f;
    const float solarRadius = sunAngle * kPi / 180.0f / 2.0f;
    const float scale = 1.0f / ((sunAngle / 0.51f) * (sunAngle / 0.51f));
    f3 xyz = F3(0.0f);
    const float solRadSin = sinf(solarRadius);
    const float ar2 = 1.0f / (solRadSin * solRadSin);
    const float singamma = sinf(gamma);
    float sc2 = 1.0f - ar2 * singamma * singamma;
    if (sc2 < 0.0f) sc2 = 0.0f;
    const float sampleCosine = sqrtf(sc2);
    if (sampleCosine == 0.0f) return F3(0.0f);
    for (int c = 0; c < 10; ++c)
    {
        const int pieces = 45, order = 4;
        int pos = (int)(powf(2.0f * elevation / kPi, 1.0f / 3.0f) * pieces);
        if (pos > 44) pos = 44;
        const float breakX = powf(((float)pos / (float)pieces), 3.0f) * (kPi * 0.5f);
        const float *coefs = t.solar + c * 180 + (order * (pos + 1) - 1);
        float res = 0.0f;
        const float x = elevation - breakX;
        float xExp = 1.0f;
        for (int i = 0; i < order; ++i) { res += xExp * *coefs--; xExp *= x; }
        float direct = res;
        const float *ld = t.limb + c * 6;
        const float dark = ld[0] + ld[1] * sampleCosine + ld[2] * powf(sampleCosine, 2.0f) + ld[3] * powf(sampleCosine, 3.0f) +
                           ld[4] * powf(sampleCosine, 4.0f) + ld[5] * powf(sampleCosine, 5.0f);
        direct *= dark * scale;
        xyz += direct * spectrumToXyz(c);
    }
    return xyzToRgbSrgb(xyz);
}
inline f3 equalAreaHemisphereMap(float u, float v)
{
    const float r = sqrtf(1.0f - v * v);
    const float phi = kTwoPi * u;
    return F3(r * cosf(phi), v, r * sinf(phi));
}
// SkyModel::update: sun direction (Sky.cu:362-367)
inline f3 skySunDir(float timeOfDay, float sunAxisAngle, float sunAxisRotate)
{
    const float d2r = kPi / 180.0f;
    f3 axis = F3(1.0f, cosf(sunAxisAngle * d2r), sinf(sunAxisAngle * d2r));
    axis = axis * F3(sinf(sunAxisRotate * d2r), 1.0f, cosf(sunAxisRotate * d2r));
    axis = normalize(axis);
    const float angle = fmodf(timeOfDay * kPi, kTwoPi);
    // rotate3f(axis, angle, v) = rotate(Quat::axisAngle(axis, angle), v).v
    const f3 v = cross(F3(0.0f, 1.0f, 0.0f), axis);
    const quat q = {normalize(axis) * sinf(angle / 2), cosf(angle / 2)};
    return normalize(qrotate(q, v));
}
// params: timeOfDay, sunAxisAngle, sunAxisRotate, skyBrightness. sky: skyW x skyH float4 (w = 0), pdf planes = luminance.
inline void generateSky(const float *params, const float *tableBlob, int skyW, int skyH, int sunW, int sunH, f4 *sky, f4 *sun, float *skyPdf,
                        float *sunPdf, f3 *sunDirOut)
{
    const SkyTables t = skyTablesFrom(tableBlob);
    const f3 sunDir = skySunDir(params[0], params[1], params[2]);
    const float brightness = params[3];
    *sunDirOut = sunDir;
    const SkyState st = skyUpdateState(t, sunDir);
    const int half = skyH / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < half; ++y)
        for (int x = 0; x < skyW; ++x)
        {
            const float u = ((float)x + 0.5f) / skyW, v = ((float)y + 0.5f) / half;
            const f3 dir = equalAreaHemisphereMap(u, v);
            f3 color = skyRadiance(st, dir, sunDir) * brightness;
            color = max3f(color, F3(0.0f));
            const size_t i = (size_t)skyW * (y + half) + x;
            sky[i] = F4(color, 0.0f);
            skyPdf[i] = luminance(color);
        }
    double sum = 0.0;
    for (size_t i = (size_t)skyW * half; i < (size_t)skyW * skyH; ++i) sum += (double)skyPdf[i];
    const float sumSkyPdf = (float)sum;
    for (int y = 0; y < half; ++y)
        for (int x = 0; x < skyW; ++x)
        {
            const float v = ((float)y + 0.5f) / half - 1.0f;
            const f3 mist = F3(sumSkyPdf / (skyW * skyH));
            const float w = clampf((v + 0.4f) * (1.0f / 0.5f), 0.0f, 1.0f);
            const f3 horizon = xyz(sky[(size_t)skyW * half + x]);
            const f3 color = mist + (w * w * (3.0f - 2.0f * w)) * (horizon - mist);
            const size_t i = (size_t)skyW * y + x;
            sky[i] = F4(color, 0.0f);
            skyPdf[i] = luminance(color);
        }
    const float cosMax = cosf(0.51f * kPi / 180.0f / 2.0f);
    for (int y = 0; y < sunH; ++y)
        for (int x = 0; x < sunW; ++x)
        {
            const float u = ((float)x + 0.5f) / sunW, v = ((float)y + 0.5f) / sunH;
            const f3 dir = equalAreaMapCone(sunDir, u, v, cosMax);
            f3 color = sunRadiance(t, dir, sunDir) * brightness;
            color = max3f(color, F3(0.0f));
            const size_t i = (size_t)sunW * y + x;
            sun[i] = F4(color, 0.0f);
            sunPdf[i] = luminance(color);
        }
}

// ---------------------------------------------------------------------------------------------- output stage
// FilmicToneMapping (/root/reference/renderer/postprocessing/FilmicToneMapping.h:12-117) with the manual exposure, then the
// PNG conversion of OfflineBackend::writeFrameBufferToPNG (/root/reference/renderer/core/OfflineBackend.cpp:191-221).
struct ToneMappingParams { float manualExposure; int curve; float highlightDesaturation, whitePoint, contrast, saturation, lift, gain; };
inline f3 clamp01(f3 v) { return {clampf(v.x, 0.0f, 1.0f), clampf(v.y, 0.0f, 1.0f), clampf(v.z, 0.0f, 1.0f)}; }
inline f3 acesFilm(f3 x)
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return clamp01(x * (a * x + F3(b)) / (x * (c * x + F3(d)) + F3(e)));
}
inline f3 uncharted2(f3 x)
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, Fc = 0.30f;
    return ((x * (A * x + F3(C * B)) + F3(D * E)) / (x * (A * x + F3(B)) + F3(D * Fc))) - F3(E / Fc);
}
inline float linearToSrgb(float c) { return (c <= 0.0031308f) ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f; }
inline void tonemap(const f4 *hdr, int W, int H, const ToneMappingParams &p, uint8_t *rgb8, f4 *ldr)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            const f4 in = hdr[(size_t)y * W + x];
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
            if (ldr) ldr[(size_t)y * W + x] = F4(tm, 1.0f);
            if (rgb8)
            {
                uint8_t *q = rgb8 + ((size_t)(H - 1 - y) * W + x) * 3;
                q[0] = (uint8_t)(fminf(1.0f, fmaxf(0.0f, tm.x)) * 255.0f);
                q[1] = (uint8_t)(fminf(1.0f, fmaxf(0.0f, tm.y)) * 255.0f);
                q[2] = (uint8_t)(fminf(1.0f, fmaxf(0.0f, tm.z)) * 255.0f);
            }
        }
}

} // namespace orc
