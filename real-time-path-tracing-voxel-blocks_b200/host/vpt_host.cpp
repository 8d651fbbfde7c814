// Host-side mirror of the reference's host code on the hot path (no device work), exported through the
// C ABI (include/vpt.h). Compiled by g++ with -ffp-contract=off so the camera matrices and the noise maps
// are the same bits the CPU oracle derives.
//
//   Camera::init / updateMatrices / update      /root/reference/renderer/shaders/Camera.h:29-99
//   YawPitchToDir / DirToYawPitch               /root/reference/renderer/shaders/LinearMath.h:1692-1740
//   offline camera setup                        /root/reference/mainOffline.cpp:227-247
//   PerlinNoiseGenerator (siv::PerlinNoise)     /root/reference/voxelengine/Noise.cpp:4-18, ext/PerlinNoise.hpp:228-243, 449-497
//   chunk noise maps                            /root/reference/voxelengine/VoxelSceneGen.cu:358-376
//   AliasTable::update (CPU build)              /root/reference/renderer/shaders/AliasTable.cu:66-153
//   GlobalSettings::LoadFromYAML                /root/reference/renderer/core/GlobalSettings.cpp:69-138, 456-492
//   SceneConfigParser::LoadFromFile             /root/reference/renderer/core/SceneConfig.cpp:6-114, 184-247
#include "../../include/vpt.h"
#include "../csrc/vpt_fastdiv.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct V3 { float x, y, z; };
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dopf(float a, float b, float c, float d)
{
    float cd = c * d;
    float err = fmaf(-c, d, cd);
    float r = fmaf(a, b, -cd);
    return r + err;
}
inline V3 cross3(V3 a, V3 b) { return {dopf(a.y, b.z, a.z, b.y), dopf(a.z, b.x, a.x, b.z), dopf(a.x, b.y, a.y, b.x)}; }
inline V3 normalize3(V3 v)
{
    float n = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (n < 1e-8f || std::isnan(n)) return {0.0f, 0.0f, 1.0f};
    return {v.x / n, v.y / n, v.z / n};
}
// compensated 3-term inner product (LinearMath.h:112-146)
inline float inner3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    float p0 = a0 * b0, e0 = fmaf(a0, b0, -p0);
    float p1 = a1 * b1, e1 = fmaf(a1, b1, -p1);
    float p2 = a2 * b2, e2 = fmaf(a2, b2, -p2);
    float s12 = p1 + p2, d12 = s12 - p1, se12 = (p1 - (s12 - d12)) + (p2 - d12);
    float tpv = s12, tpe = e1 + (e2 + se12);
    float s = p0 + tpv, d = s - p0, se = (p0 - (s - d)) + (tpv - d);
    return s + (e0 + (tpe + se));
}
struct M3 { float m00, m10, m20, m01, m11, m21, m02, m12, m22; };
inline M3 cols(V3 a, V3 b, V3 c) { return {a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z}; }
inline M3 zero3() { M3 m; std::memset(&m, 0, sizeof m); return m; }
inline M3 transpose3(M3 m)
{
    M3 r = m;
    r.m01 = m.m10; r.m10 = m.m01; r.m02 = m.m20; r.m20 = m.m02; r.m12 = m.m21; r.m21 = m.m12;
    return r;
}
inline M3 mul3(const M3 &A, const M3 &B)
{
    M3 C;
    C.m00 = A.m00 * B.m00 + A.m01 * B.m10 + A.m02 * B.m20;
    C.m01 = A.m00 * B.m01 + A.m01 * B.m11 + A.m02 * B.m21;
    C.m02 = A.m00 * B.m02 + A.m01 * B.m12 + A.m02 * B.m22;
    C.m10 = A.m10 * B.m00 + A.m11 * B.m10 + A.m12 * B.m20;
    C.m11 = A.m10 * B.m01 + A.m11 * B.m11 + A.m12 * B.m21;
    C.m12 = A.m10 * B.m02 + A.m11 * B.m12 + A.m12 * B.m22;
    C.m20 = A.m20 * B.m00 + A.m21 * B.m10 + A.m22 * B.m20;
    C.m21 = A.m20 * B.m01 + A.m21 * B.m11 + A.m22 * B.m21;
    C.m22 = A.m20 * B.m02 + A.m21 * B.m12 + A.m22 * B.m22;
    return C;
}
inline void store(float *dst, const M3 &m) { std::memcpy(dst, &m, sizeof m); }

constexpr float kPiOver2 = 1.5707963267948966192313216916397514420985f;
constexpr float kPiOver180 = 0.01745329251f;
inline float clampf(float a, float lo, float hi) { return a < lo ? lo : a > hi ? hi : a; }

V3 yawPitchToDir(float yaw, float pitch)
{
    if (std::isnan(yaw) || std::isnan(pitch)) return {0, 0, 1};
    pitch = clampf(pitch, -kPiOver2 + 0.01f, kPiOver2 - 0.01f);
    float sy = sinf(yaw), cy = cosf(yaw), sp = sinf(pitch), cp = cosf(pitch);
    return normalize3({sy * cp, sp, cy * cp});
}

// ---- siv::BasicPerlinNoise<float>
struct PerlinNoise
{
    std::array<uint8_t, 256> perm;
    explicit PerlinNoise(uint32_t seed)
    {
        for (int i = 0; i < 256; ++i) perm[i] = (uint8_t)i;
        std::mt19937 urbg(seed);
        for (int i = 1; i < 256; ++i)
        {
            const uint64_t j = (uint64_t)urbg() % ((uint64_t)i + 1);
            std::swap(perm[i], perm[(size_t)j]);
        }
    }
    static float fade(float t) { return t * t * t * (t * (t * 6 - 15) + 10); }
    static float mix(float a, float b, float t) { return a + (b - a) * t; }
    static float grad(uint8_t hash, float x, float y, float z)
    {
        const uint8_t h = hash & 15;
        const float u = h < 8 ? x : y;
        const float v = h < 4 ? y : (h == 12 || h == 14 ? x : z);
        return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
    }
    float noise3D(float x, float y, float z) const
    {
        const float fx0 = std::floor(x), fy0 = std::floor(y), fz0 = std::floor(z);
        const int ix = (int)fx0 & 255, iy = (int)fy0 & 255, iz = (int)fz0 & 255;
        const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
        const float u = fade(fx), v = fade(fy), w = fade(fz);
        const uint8_t A = (perm[ix & 255] + iy) & 255, B = (perm[(ix + 1) & 255] + iy) & 255;
        const uint8_t AA = (perm[A] + iz) & 255, AB = (perm[(A + 1) & 255] + iz) & 255;
        const uint8_t BA = (perm[B] + iz) & 255, BB = (perm[(B + 1) & 255] + iz) & 255;
        const float p0 = grad(perm[AA], fx, fy, fz), p1 = grad(perm[BA], fx - 1, fy, fz);
        const float p2 = grad(perm[AB], fx, fy - 1, fz), p3 = grad(perm[BB], fx - 1, fy - 1, fz);
        const float p4 = grad(perm[(AA + 1) & 255], fx, fy, fz - 1), p5 = grad(perm[(BA + 1) & 255], fx - 1, fy, fz - 1);
        const float p6 = grad(perm[(AB + 1) & 255], fx, fy - 1, fz - 1), p7 = grad(perm[(BB + 1) & 255], fx - 1, fy - 1, fz - 1);
        const float q0 = mix(p0, p1, u), q1 = mix(p2, p3, u), q2 = mix(p4, p5, u), q3 = mix(p6, p7, u);
        return mix(mix(q0, q1, v), mix(q2, q3, v), w);
    }
    float octave2D_01(float x, float y, int octaves) const
    {
        float result = 0, amplitude = 1;
        for (int i = 0; i < octaves; ++i)
        {
            result += noise3D(x, y, (float)0.34567) * amplitude;
            x *= 2; y *= 2; amplitude *= 0.5f;
        }
        if (result <= -1.0f) return 0.0f;
        if (1.0f <= result) return 1.0f;
        return result * 0.5f + 0.5f;
    }
};

std::string trim(const std::string &s)
{
    size_t b = s.find_first_not_of(" \t\r\n");
    if (b == std::string::npos) return "";
    size_t e = s.find_last_not_of(" \t\r\n");
    return s.substr(b, e - b + 1);
}
bool parseFloat(const std::string &v, float &out)
{
    try { out = std::stof(trim(v)); return true; } catch (const std::exception &) { return false; }
}
bool parseInt(const std::string &v, int &out)
{
    try { out = std::stoi(trim(v)); return true; } catch (const std::exception &) { return false; }
}
bool parseBool(const std::string &v, int32_t &out)
{
    std::string t = trim(v);
    std::transform(t.begin(), t.end(), t.begin(), ::tolower);
    if (t == "true" || t == "1" || t == "yes" || t == "on") { out = 1; return true; }
    if (t == "false" || t == "0" || t == "no" || t == "off") { out = 0; return true; }
    return false;
}
bool parseFloat3(const std::string &value, float *out)
{
    std::string c = value;
    if (!c.empty() && c.front() == '[') c = c.substr(1);
    if (!c.empty() && c.back() == ']') c = c.substr(0, c.size() - 1);
    std::istringstream iss(c);
    std::string tok;
    float vals[3];
    int count = 0;
    while (std::getline(iss, tok, ',') && count < 3)
    {
        try { vals[count++] = std::stof(trim(tok)); } catch (const std::exception &) { return false; }
    }
    if (count != 3) return false;
    out[0] = vals[0]; out[1] = vals[1]; out[2] = vals[2];
    return true;
}

} // namespace

extern "C" {

void vpt_camera_init(VptCamera *c, int width, int height)
{
    std::memset(c, 0, sizeof *c);
    c->pos[0] = 16.0f; c->pos[1] = 25.0f; c->pos[2] = 16.0f;
    V3 d = normalize3({1.0f, -1.0f, 1.0f});
    c->dir[0] = d.x; c->dir[1] = d.y; c->dir[2] = d.z;
    c->resolution[0] = (float)width; c->resolution[1] = (float)height;
    c->inversedResolution[0] = 1.0f / c->resolution[0]; c->inversedResolution[1] = 1.0f / c->resolution[1];
    float fovX = 90.0f * kPiOver180;
    float fovY = fovX * (c->resolution[1] / c->resolution[0]);
    c->tanHalfFov[0] = tanf(fovX * 0.5f); c->tanHalfFov[1] = tanf(fovY * 0.5f);
}

void vpt_camera_update(VptCamera *c)
{
    float *pd = c->posDelta;
    if (std::isnan(pd[0]) || std::isnan(pd[1]) || std::isnan(pd[2]) || fabsf(pd[0]) > 1000.0f || fabsf(pd[1]) > 1000.0f || fabsf(pd[2]) > 1000.0f)
        pd[0] = pd[1] = pd[2] = 0.0f;
    for (int i = 0; i < 3; ++i) { c->pos[i] = c->pos[i] + pd[i]; pd[i] = 0.0f; }
    V3 dir = yawPitchToDir(c->yaw, c->pitch);
    c->dir[0] = dir.x; c->dir[1] = dir.y; c->dir[2] = dir.z;
    V3 worldUp = {0.0f, 1.0f, 0.0f};
    V3 left = normalize3(cross3(worldUp, dir));
    V3 up = normalize3(cross3(dir, left));
    M3 uvToNdc = cols({2, 0, 0}, {0, 2, 0}, {-1, -1, 1});
    M3 ndcToView = zero3();
    ndcToView.m00 = c->tanHalfFov[0]; ndcToView.m11 = c->tanHalfFov[1]; ndcToView.m22 = 1.0f;
    M3 viewToWorld = cols(-left, up, dir);
    store(c->uvToWorld, mul3(mul3(viewToWorld, ndcToView), uvToNdc));
    store(c->uvToView, mul3(viewToWorld, ndcToView));
    M3 ndcToUv = cols({0.5f, 0, 0}, {0, 0.5f, 0}, {0.5f, 0.5f, 1.0f});
    M3 worldToView = transpose3(viewToWorld);
    M3 viewToNdc = zero3();
    viewToNdc.m00 = 1.0f / c->tanHalfFov[0]; viewToNdc.m11 = 1.0f / c->tanHalfFov[1]; viewToNdc.m22 = 1.0f;
    store(c->worldToUv, mul3(mul3(ndcToUv, viewToNdc), worldToView));
    store(c->viewToUv, mul3(viewToNdc, worldToView));
}

void vpt_camera_from_scene(VptCamera *c, int width, int height, const float *position, const float *direction, float fovDegrees)
{
    vpt_camera_init(c, width, height);
    c->pos[0] = position[0]; c->pos[1] = position[1]; c->pos[2] = position[2];
    // SceneConfigParser normalises the direction (SceneConfig.cpp:108), mainOffline normalises again (:235)
    V3 dir = normalize3(normalize3({direction[0], direction[1], direction[2]}));
    // DirToYawPitch: Float3::normalize() member uses the compensated length (LinearMath.h:536-548)
    float n = sqrtf(inner3(dir.x, dir.x, dir.y, dir.y, dir.z, dir.z));
    dir = {dir.x / n, dir.y / n, dir.z / n};
    c->yaw = atan2f(dir.x, dir.z);
    c->pitch = asinf(dir.y);
    float fovX = fovDegrees * kPiOver180;
    float fovY = fovX * (c->resolution[1] / c->resolution[0]);
    c->tanHalfFov[0] = tanf(fovX * 0.5f); c->tanHalfFov[1] = tanf(fovY * 0.5f);
    vpt_camera_update(c);
}

void vpt_perlin_noise_chunks(int chunksX, int chunksY, int chunksZ, unsigned seed, float *out)
{
    PerlinNoise gen(seed); // same seed for all chunks, 4 octaves (VoxelSceneGen.cu:358)
    const float freq = 1.0f / (float)(chunksX * 32);
    for (int c = 0; c < chunksX * chunksY * chunksZ; ++c)
    {
        const int chunkX = c % chunksX, chunkZ = (c / chunksX) % chunksZ;
        float *noise = out + (size_t)c * 1024;
        for (int x = 0; x < 32; ++x)
            for (int z = 0; z < 32; ++z)
            {
                float gx = (float)(chunkX * 32 + x), gz = (float)(chunkZ * 32 + z);
                noise[z * 32 + x] = gen.octave2D_01(gx * freq, gz * freq, 4);
            }
    }
}

void vpt_build_alias_table(const float *weights, unsigned n, VptAliasBin *bins)
{
    // thrust::reduce's order is unspecified in the reference; sum in double, round once
    double acc = 0.0;
    for (unsigned i = 0; i < n; ++i) acc += weights[i];
    const float sum = (float)acc;
    std::vector<float> prob(n), scaled(n);
    std::vector<int> alias(n, -1);
    for (unsigned i = 0; i < n; ++i) { float p = weights[i] / sum; prob[i] = p; scaled[i] = p * n; }
    std::vector<int> smallQ, largeQ;
    smallQ.reserve(n); largeQ.reserve(n);
    size_t sh = 0, lh = 0;
    for (unsigned i = 0; i < n; ++i) (scaled[i] < 1.0f ? smallQ : largeQ).push_back((int)i);
    while (sh < smallQ.size() && lh < largeQ.size())
    {
        const int s = smallQ[sh++], l = largeQ[lh++];
        alias[s] = l;
        scaled[l] -= (1.0f - scaled[s]);
        (scaled[l] < 1.0f ? smallQ : largeQ).push_back(l);
    }
    while (sh < smallQ.size()) scaled[smallQ[sh++]] = 1.0f;
    while (lh < largeQ.size()) scaled[largeQ[lh++]] = 1.0f;
    for (unsigned i = 0; i < n; ++i) { bins[i].p = prob[i]; bins[i].q = scaled[i]; bins[i].alias = alias[i]; }
}

void vpt_default_denoising_params(VptDenoisingParams *p)
{
    p->enableHitDistanceReconstruction = 0; p->enablePrePass = 0; p->enableTemporalAccumulation = 1; p->enableHistoryFix = 1;
    p->enableHistoryClamping = 1; p->enableSpatialFiltering = 1; p->enableFireflyFilter = 1;
    p->maxAccumulatedFrameNum = 30.0f; p->maxFastAccumulatedFrameNum = 6.0f;
    p->phiLuminance = 2.0f; p->lobeAngleFraction = 0.5f; p->roughnessFraction = 0.15f; p->depthThreshold = 0.003f;
    p->atrousIterationNum = 5;
    p->disocclusionThreshold = 0.01f; p->disocclusionThresholdAlternate = 0.05f; p->denoisingRange = 500000.0f;
}

int vpt_load_denoising_settings(const char *yamlPath, VptDenoisingParams *p)
{
    std::ifstream file(yamlPath);
    if (!file.is_open()) return VPT_ERR_IO;
    std::string line, section;
    while (std::getline(file, line))
    {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        if (line.back() == ':') { section = line.substr(0, line.size() - 1); continue; }
        size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        const std::string key = trim(line.substr(0, colon)), value = trim(line.substr(colon + 1));
        if (section != "denoising") continue;
        if (key == "enableHitDistanceReconstruction") parseBool(value, p->enableHitDistanceReconstruction);
        else if (key == "enablePrePass") parseBool(value, p->enablePrePass);
        else if (key == "enableTemporalAccumulation") parseBool(value, p->enableTemporalAccumulation);
        else if (key == "enableHistoryFix") parseBool(value, p->enableHistoryFix);
        else if (key == "enableHistoryClamping") parseBool(value, p->enableHistoryClamping);
        else if (key == "enableSpatialFiltering") parseBool(value, p->enableSpatialFiltering);
        else if (key == "enableFireflyFilter") parseBool(value, p->enableFireflyFilter);
        else if (key == "maxAccumulatedFrameNum") parseFloat(value, p->maxAccumulatedFrameNum);
        else if (key == "maxFastAccumulatedFrameNum") parseFloat(value, p->maxFastAccumulatedFrameNum);
        else if (key == "phiLuminance") parseFloat(value, p->phiLuminance);
        else if (key == "lobeAngleFraction") parseFloat(value, p->lobeAngleFraction);
        else if (key == "roughnessFraction") parseFloat(value, p->roughnessFraction);
        else if (key == "depthThreshold") parseFloat(value, p->depthThreshold);
        else if (key == "atrousIterationNum") { int v; if (parseInt(value, v)) p->atrousIterationNum = v; }
        else if (key == "disocclusionThreshold") parseFloat(value, p->disocclusionThreshold);
        else if (key == "disocclusionThresholdAlternate") parseFloat(value, p->disocclusionThresholdAlternate);
        else if (key == "denoisingRange") parseFloat(value, p->denoisingRange);
    }
    return VPT_OK;
}

void vpt_default_tonemapping_params(VptToneMappingParams *p)
{
    p->manualExposure = 10.0f; p->curve = 0; p->highlightDesaturation = 0.8f; p->whitePoint = 10.0f;
    p->contrast = 1.0f; p->saturation = 1.0f; p->lift = 0.0f; p->gain = 1.0f;
}
// one pass of the reference's line parser over a section: calls f(key, value) for its "key: value" lines
extern "C++" template <typename F> int parseSection(const char *yamlPath, const char *wanted, F f)
{
    std::ifstream file(yamlPath);
    if (!file.is_open()) return VPT_ERR_IO;
    std::string line, section;
    while (std::getline(file, line))
    {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        if (line.back() == ':') { section = line.substr(0, line.size() - 1); continue; }
        size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        if (section != wanted) continue;
        f(trim(line.substr(0, colon)), trim(line.substr(colon + 1)));
    }
    return VPT_OK;
}
int vpt_load_tonemapping_settings(const char *yamlPath, VptToneMappingParams *p)
{
    return parseSection(yamlPath, "postprocess", [&](const std::string &key, const std::string &value) {
        if (key == "manualExposure") parseFloat(value, p->manualExposure);
        else if (key == "toneMappingCurve") { int v; if (parseInt(value, v) && v >= 0 && v <= 2) p->curve = v; }
        else if (key == "highlightDesaturation") parseFloat(value, p->highlightDesaturation);
        else if (key == "whitePoint") parseFloat(value, p->whitePoint);
        else if (key == "contrast") parseFloat(value, p->contrast);
        else if (key == "saturation") parseFloat(value, p->saturation);
        else if (key == "gain") parseFloat(value, p->gain);
        else if (key == "lift") parseFloat(value, p->lift);
    });
}
int vpt_load_sky_settings(const char *yamlPath, VptSkyParams *p)
{
    return parseSection(yamlPath, "sky", [&](const std::string &key, const std::string &value) {
        if (key == "timeOfDay") parseFloat(value, p->timeOfDay);
        else if (key == "sunAxisAngle") parseFloat(value, p->sunAxisAngle);
        else if (key == "sunAxisRotate") parseFloat(value, p->sunAxisRotate);
        else if (key == "skyBrightness") parseFloat(value, p->skyBrightness);
    });
}

int vpt_load_scene_config(const char *yamlPath, float *out9, float *fov, unsigned *chunks3)
{
    // CameraConfig defaults (SceneConfig.h:10-16)
    float pos[3] = {20.0f, 15.0f, 20.0f}, dir[3] = {-1.0f, -0.3f, -1.0f}, up[3] = {0.0f, 1.0f, 0.0f};
    *fov = 90.0f;
    chunks3[0] = chunks3[1] = chunks3[2] = 0;
    std::ifstream file(yamlPath);
    const bool ok = file.is_open();
    if (ok)
    {
        std::string line, section;
        while (std::getline(file, line))
        {
            line = trim(line);
            if (line.empty() || line[0] == '#') continue;
            if (line.back() == ':') { section = line.substr(0, line.size() - 1); continue; }
            size_t colon = line.find(':');
            if (colon == std::string::npos) continue;
            const std::string key = trim(line.substr(0, colon)), value = trim(line.substr(colon + 1));
            if (section == "camera")
            {
                if (key == "position") parseFloat3(value, pos);
                else if (key == "direction") parseFloat3(value, dir);
                else if (key == "up") parseFloat3(value, up);
                else if (key == "fov") parseFloat(value, *fov);
            }
            else if (section == "chunk_config")
            {
                try
                {
                    if (key == "chunksX") chunks3[0] = (unsigned)std::stoul(value);
                    else if (key == "chunksY") chunks3[1] = (unsigned)std::stoul(value);
                    else if (key == "chunksZ") chunks3[2] = (unsigned)std::stoul(value);
                }
                catch (const std::exception &) {}
            }
        }
    }
    V3 nd = normalize3({dir[0], dir[1], dir[2]});
    out9[0] = pos[0]; out9[1] = pos[1]; out9[2] = pos[2];
    out9[3] = nd.x; out9[4] = nd.y; out9[5] = nd.z;
    out9[6] = up[0]; out9[7] = up[1]; out9[8] = up[2];
    return ok ? VPT_OK : VPT_ERR_IO;
}

} // extern "C"

// ---- sky state on the host: SkyModel::update's sun direction (renderer/sky/Sky.cu:362-367) and updateSkyState
// (Sky.cu:52-79; getFittingData/2 :18-50). tables = data/sky_tables.bin (skyDataSets[540], skyDataSetsRad[60], ...).
namespace {
inline float dot3c(V3 a, V3 b) { return inner3(a.x, b.x, a.y, b.y, a.z, b.z); }
struct Q4 { V3 v; float w; };
inline V3 scale3(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 add3(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Q4 qmul(Q4 p, Q4 q) { return {add3(add3(scale3(q.v, p.w), scale3(p.v, q.w)), cross3(p.v, q.v)), p.w * q.w - dot3c(p.v, q.v)}; }
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
} // namespace
extern "C" void vpt_sky_state(const VptSkyParams *p, const float *tables, float *configs90, float *radiances10, float *sunDir3)
{
    const float kPi = 3.1415926535897932384626422832795028841971f, kTwoPi = 6.2831853071795864769252867665590057683943f;
    const float d2r = kPi / 180.0f;
    V3 axis = {1.0f, cosf(p->sunAxisAngle * d2r), sinf(p->sunAxisAngle * d2r)};
    axis = {axis.x * sinf(p->sunAxisRotate * d2r), axis.y * 1.0f, axis.z * cosf(p->sunAxisRotate * d2r)};
    axis = normalize3(axis);
    const float angle = fmodf(p->timeOfDay * kPi, kTwoPi);
    const V3 v = cross3({0.0f, 1.0f, 0.0f}, axis);
    const Q4 q = {scale3(normalize3(axis), sinf(angle / 2)), cosf(angle / 2)};
    const Q4 r = qmul(qmul(q, Q4{v, 0.0f}), Q4{-q.v, q.w});
    const V3 sd = normalize3(r.v);
    sunDir3[0] = sd.x; sunDir3[1] = sd.y; sunDir3[2] = sd.z;
    const float elevation = (kPi / 2.0f) - (float)acos((double)sd.y);
    const float solarElevation = powf(elevation / (kPi / 2.0f), (1.0f / 3.0f));
    for (int c = 0; c < 10; ++c)
    {
        for (int i = 0; i < 9; ++i) configs90[c * 9 + i] = skyFit(tables + c * 54, solarElevation, i);
        radiances10[c] = skyFit2(tables + 540 + c * 6, solarElevation);
    }
}

// ---- world chunk files (renderer/core/WorldSceneManager.cpp:240-308, 310-458; SceneConfig.cpp:95-148)
// ---- material / block tables (AssetRegistry.cpp:60-150, 260-300; MaterialManager.cpp:58-120, 151-190; MaterialDefinition.h:18-28)
namespace {
struct YamlLine { int indent; std::string key, value; bool item; };
// "  - key: value  # comment" -> indent of the key, item flag for a leading "- "; quotes stripped
bool splitYamlLine(const std::string &raw, YamlLine &o)
{
    size_t i = 0;
    while (i < raw.size() && raw[i] == ' ') ++i;
    if (i >= raw.size() || raw[i] == '#') return false;
    o.item = false;
    if (raw[i] == '-' && i + 1 < raw.size() && raw[i + 1] == ' ') { o.item = true; i += 2; while (i < raw.size() && raw[i] == ' ') ++i; }
    o.indent = (int)i;
    const size_t colon = raw.find(':', i);
    if (colon == std::string::npos) return false;
    o.key = trim(raw.substr(i, colon - i));
    std::string v = raw.substr(colon + 1);
    bool inQuote = false;
    for (size_t k = 0; k < v.size(); ++k)
    {
        if (v[k] == '"') inQuote = !inQuote;
        else if (v[k] == '#' && !inQuote) { v = v.substr(0, k); break; }
    }
    v = trim(v);
    if (v.size() >= 2 && v.front() == '"' && v.back() == '"') v = v.substr(1, v.size() - 2);
    o.value = v;
    return true;
}
void copyPath(char *dst, const std::string &v) { std::string p = "data/" + v; std::strncpy(dst, p.c_str(), 255); dst[255] = 0; }
} // namespace

extern "C" int vpt_load_materials(const char *materialsYamlPath, const char *blocksYamlPath, VptMaterial *materials, VptMaterialTexturePaths *paths,
                                  int maxMaterials, int *count, uint16_t *blockToMaterial256)
{
    if (!materialsYamlPath || !materials || !count || maxMaterials <= 0) return VPT_ERR_ARG;
    std::ifstream mf(materialsYamlPath);
    if (!mf.is_open()) return VPT_ERR_IO;
    std::map<std::string, int> idToIndex;
    std::vector<std::array<float, 3>> emissive;
    std::string line, sub;
    int n = 0, itemIndent = -1;
    bool inList = false;
    while (std::getline(mf, line))
    {
        YamlLine y;
        if (!splitYamlLine(line, y)) continue;
        if (y.indent == 0) { inList = (y.key == "materials"); continue; }
        if (!inList) continue;
        if (y.item)
        {
            if (n >= maxMaterials) return VPT_ERR_ARG;
            VptMaterial &m = materials[n];
            std::memset(&m, 0, sizeof m);
            m.albedo[0] = m.albedo[1] = m.albedo[2] = 1.0f; m.roughness = 0.5f; m.uvScale = 1.0f; // MaterialProperties defaults
            m.materialId = n;
            if (paths) std::memset(&paths[n], 0, sizeof paths[n]);
            emissive.push_back({0.0f, 0.0f, 0.0f});
            itemIndent = y.indent; sub.clear();
            ++n;
        }
        if (n == 0) continue;
        VptMaterial &m = materials[n - 1];
        if (y.indent == itemIndent)
        {
            sub = (y.value.empty() && (y.key == "textures" || y.key == "properties")) ? y.key : std::string();
            if (y.key == "id") idToIndex[y.value] = n - 1;
            continue;
        }
        if (sub == "textures" && paths)
        {
            if (y.key == "albedo") copyPath(paths[n - 1].albedo, y.value);
            else if (y.key == "normal") copyPath(paths[n - 1].normal, y.value);
            else if (y.key == "roughness") copyPath(paths[n - 1].roughness, y.value);
            else if (y.key == "metallic") copyPath(paths[n - 1].metallic, y.value);
        }
        else if (sub == "properties")
        {
            float f; int32_t b;
            if (y.key == "albedo") parseFloat3(y.value, m.albedo);
            else if (y.key == "roughness") parseFloat(y.value, m.roughness);
            else if (y.key == "metallic") { if (parseFloat(y.value, f)) m.metallic = f != 0.0f ? 1 : 0; } // float -> bool (0.8 -> true)
            else if (y.key == "uv_scale") parseFloat(y.value, m.uvScale);
            else if (y.key == "translucency") parseFloat(y.value, m.translucency);
            else if (y.key == "is_emissive") { if (parseBool(y.value, b)) m.isEmissive = b; }
            else if (y.key == "is_thinfilm") { if (parseBool(y.value, b)) m.isThinfilm = b; }
            else if (y.key == "use_world_grid_uv") { if (parseBool(y.value, b)) m.useWorldGridUV = b; }
            else if (y.key == "emissive_radiance") parseFloat3(y.value, emissive[(size_t)n - 1].data());
        }
    }
    for (int i = 0; i < n; ++i)
        if (materials[i].isEmissive) { materials[i].albedo[0] = emissive[(size_t)i][0]; materials[i].albedo[1] = emissive[(size_t)i][1]; materials[i].albedo[2] = emissive[(size_t)i][2]; }
    *count = n;
    if (n == 0) return VPT_ERR_IO; // "No materials found in registry" (MaterialManager.cpp:64-68)
    if (blockToMaterial256)
    {
        std::memset(blockToMaterial256, 0, 256 * sizeof(uint16_t)); // unmapped blocks -> material 0 (getMaterialIndexForBlock default)
        if (!blocksYamlPath) return VPT_OK;
        std::ifstream bf(blocksYamlPath);
        if (!bf.is_open()) return VPT_ERR_IO;
        int blockId = -1; inList = false;
        while (std::getline(bf, line))
        {
            YamlLine y;
            if (!splitYamlLine(line, y)) continue;
            if (y.indent == 0) { inList = (y.key == "blocks"); continue; }
            if (!inList) continue;
            if (y.item) { blockId = -1; itemIndent = y.indent; }
            if (y.indent != itemIndent) continue;
            if (y.key == "id") { int v; if (parseInt(y.value, v)) blockId = v; }
            else if (y.key == "material" && blockId >= 0 && blockId < 256 && y.value != "null" && !y.value.empty())
            {
                auto it = idToIndex.find(y.value);
                if (it != idToIndex.end()) blockToMaterial256[blockId] = (uint16_t)it->second;
            }
        }
    }
    return VPT_OK;
}

// ---- PNG reader (the reference decodes data/textures/*.png with stb_image, TextureManager.cu:178): 8/16-bit, grey / grey+alpha /
// RGB / palette / RGBA, non-interlaced; zlib stream inflated here (stored, fixed and dynamic Huffman blocks). No dependency.
namespace {
struct BitReader
{
    const uint8_t *p, *end;
    uint32_t acc = 0; int n = 0;
    bool need(int k) { while (n < k) { if (p >= end) return false; acc |= (uint32_t)*p++ << n; n += 8; } return true; }
    bool bits(int k, uint32_t &v) { if (k == 0) { v = 0; return true; } if (!need(k)) return false; v = acc & ((1u << k) - 1u); acc >>= k; n -= k; return true; }
};
struct Huffman
{
    uint16_t count[16], symbol[288];
    void build(const uint8_t *len, int nsym)
    {
        std::memset(count, 0, sizeof count);
        for (int i = 0; i < nsym; ++i) count[len[i]]++;
        count[0] = 0;
        uint16_t offs[16]; offs[1] = 0;
        for (int i = 1; i < 15; ++i) offs[i + 1] = offs[i] + count[i];
        for (int i = 0; i < nsym; ++i) if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    }
    int decode(BitReader &br) const
    {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; ++l)
        {
            uint32_t b;
            if (!br.bits(1, b)) return -1;
            code |= (int)b;
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};
bool inflateZlib(const uint8_t *src, size_t n, std::vector<uint8_t> &out, size_t expected)
{
    if (n < 6) return false;
    BitReader br{src + 2, src + n};
    out.clear(); out.reserve(expected);
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint32_t final = 0;
    while (!final)
    {
        uint32_t type;
        if (!br.bits(1, final) || !br.bits(2, type)) return false;
        if (type == 0)
        {
            br.acc = 0; br.n = 0; // byte align (the accumulator only ever holds bits of bytes already consumed)
            if (br.end - br.p < 4) return false;
            const unsigned len = br.p[0] | (br.p[1] << 8);
            br.p += 4;
            if ((size_t)(br.end - br.p) < len || out.size() + len > expected + 65536) return false;
            out.insert(out.end(), br.p, br.p + len);
            br.p += len;
            continue;
        }
        if (type == 3) return false;
        Huffman lit, dist;
        uint8_t lens[320];
        if (type == 1)
        {
            for (int i = 0; i < 144; ++i) lens[i] = 8;
            for (int i = 144; i < 256; ++i) lens[i] = 9;
            for (int i = 256; i < 280; ++i) lens[i] = 7;
            for (int i = 280; i < 288; ++i) lens[i] = 8;
            lit.build(lens, 288);
            for (int i = 0; i < 30; ++i) lens[i] = 5;
            dist.build(lens, 30);
        }
        else
        {
            uint32_t hlit, hdist, hclen;
            if (!br.bits(5, hlit) || !br.bits(5, hdist) || !br.bits(4, hclen)) return false;
            hlit += 257; hdist += 1; hclen += 4;
            uint8_t cl[19] = {0};
            for (uint32_t i = 0; i < hclen; ++i) { uint32_t v; if (!br.bits(3, v)) return false; cl[order[i]] = (uint8_t)v; }
            Huffman clh; clh.build(cl, 19);
            uint32_t i = 0;
            while (i < hlit + hdist)
            {
                const int sym = clh.decode(br);
                if (sym < 0) return false;
                if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                uint32_t rep, prev = 0;
                if (sym == 16) { if (i == 0) return false; prev = lens[i - 1]; if (!br.bits(2, rep)) return false; rep += 3; }
                else if (sym == 17) { if (!br.bits(3, rep)) return false; rep += 3; }
                else { if (!br.bits(7, rep)) return false; rep += 11; }
                if (i + rep > hlit + hdist) return false;
                while (rep--) lens[i++] = (uint8_t)prev;
            }
            lit.build(lens, (int)hlit);
            dist.build(lens + hlit, (int)hdist);
        }
        for (;;)
        {
            const int sym = lit.decode(br);
            if (sym < 0) return false;
            if (out.size() > expected) return false; // more data than the image holds: corrupt stream
            if (sym < 256) { out.push_back((uint8_t)sym); continue; }
            if (sym == 256) break;
            if (sym > 285) return false;
            uint32_t eb;
            if (!br.bits(lext[sym - 257], eb)) return false;
            const size_t len = lbase[sym - 257] + eb;
            const int ds = dist.decode(br);
            if (ds < 0 || ds > 29) return false;
            if (!br.bits(dext[ds], eb)) return false;
            const size_t d = dbase[ds] + eb;
            if (d > out.size()) return false;
            const size_t from = out.size() - d;
            for (size_t k = 0; k < len; ++k) out.push_back(out[from + k]);
        }
    }
    return true;
}
inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
} // namespace

extern "C" int vpt_load_png_rgba8(const char *path, uint32_t *out, size_t maxTexels, int *width, int *height, int *channels)
{
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return VPT_ERR_IO;
    std::vector<uint8_t> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 33 || std::memcmp(file.data(), sig, 8) != 0) return VPT_ERR_IO;
    uint32_t W = 0, H = 0; int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    size_t pos = 8;
    while (pos + 12 <= file.size())
    {
        const uint32_t len = be32(&file[pos]);
        const uint8_t *type = &file[pos + 4], *data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) return VPT_ERR_IO;
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) { W = be32(data); H = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12]; }
        else if (!std::memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (W == 0 || H == 0 || W > 32768u || H > 32768u || interlace != 0 || (depth != 8 && depth != 16) || (ctype == 3 && depth != 8)) return VPT_ERR_IO;
    const int nch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (nch == 0) return VPT_ERR_IO;
    if (width) *width = (int)W;
    if (height) *height = (int)H;
    if (channels) *channels = ctype == 3 ? (trns.empty() ? 3 : 4) : nch; // what stbi_load(..., 0) reports
    if (!out) return VPT_OK; // size query
    if ((size_t)W * H > maxTexels) return VPT_ERR_ARG;
    const size_t bpp = (size_t)nch * (depth / 8), stride = bpp * W;
    std::vector<uint8_t> raw;
    if (!inflateZlib(idat.data(), idat.size(), raw, (stride + 1) * H) || raw.size() < (stride + 1) * H) return VPT_ERR_IO;
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    for (uint32_t y = 0; y < H; ++y)
    {
        const uint8_t *row = &raw[(stride + 1) * y];
        const int filter = row[0];
        for (size_t i = 0; i < stride; ++i)
        {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int pred = 0;
            if (filter == 1) pred = a;
            else if (filter == 2) pred = b;
            else if (filter == 3) pred = (a + b) >> 1;
            else if (filter == 4) { const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
            else if (filter != 0) return VPT_ERR_IO;
            cur[i] = (uint8_t)(row[1 + i] + pred);
        }
        for (uint32_t x = 0; x < W; ++x)
        {
            const uint8_t *px = &cur[x * bpp];
            const int s = depth / 8; // 16-bit samples: the high byte (stb's 8-bit conversion)
            uint32_t r, g, b, al = 255;
            if (ctype == 0) { r = g = b = px[0]; }
            else if (ctype == 4) { r = g = b = px[0]; al = px[s]; }
            else if (ctype == 2) { r = px[0]; g = px[s]; b = px[2 * s]; }
            else if (ctype == 6) { r = px[0]; g = px[s]; b = px[2 * s]; al = px[3 * s]; }
            else
            {
                const size_t k = px[0];
                if (k * 3 + 2 >= palette.size()) return VPT_ERR_IO;
                r = palette[k * 3]; g = palette[k * 3 + 1]; b = palette[k * 3 + 2];
                if (k < trns.size()) al = trns[k];
            }
            out[(size_t)y * W + x] = r | (g << 8) | (b << 16) | (al << 24);
        }
        prev.swap(cur);
    }
    return VPT_OK;
}

// ---- image diff (ImageDiff.cpp:95-372): different-pixel count, RMSE, global SSIM on the 3x3-gaussian filtered luma
extern "C" int vpt_image_diff(const uint32_t *A, const uint32_t *B, int w, int h, int channels, VptImageDiffResult *r)
{
    if (!A || !B || !r || w <= 0 || h <= 0 || channels < 1 || channels > 4) return VPT_ERR_ARG;
    const size_t n = (size_t)w * h;
    auto ch = [](uint32_t v, int c) -> int { return (int)((v >> (8 * c)) & 0xffu); };
    size_t different = 0;
    double sq = 0.0;
    std::vector<float> ga(n), gb(n);
    for (size_t i = 0; i < n; ++i)
    {
        bool diff = false;
        for (int c = 0; c < channels; ++c)
        {
            const int d = ch(A[i], c) - ch(B[i], c);
            if ((float)std::abs(d) / 255.0f > 0.01f) diff = true;
            sq += (double)d * d;
        }
        different += diff ? 1 : 0;
        // luma of the first three channels (grey images: the single channel)
        if (channels >= 3)
        {
            ga[i] = 0.299f * ch(A[i], 0) + 0.587f * ch(A[i], 1) + 0.114f * ch(A[i], 2);
            gb[i] = 0.299f * ch(B[i], 0) + 0.587f * ch(B[i], 1) + 0.114f * ch(B[i], 2);
        }
        else { ga[i] = (float)ch(A[i], 0); gb[i] = (float)ch(B[i], 0); }
    }
    auto blur = [&](const std::vector<float> &src) {
        std::vector<float> out(n);
        static const float k[3] = {1.0f, 2.0f, 1.0f};
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
            {
                float acc = 0.0f;
                for (int ky = -1; ky <= 1; ++ky)
                    for (int kx = -1; kx <= 1; ++kx)
                    {
                        const int yy = std::min(std::max(y + ky, 0), h - 1), xx = std::min(std::max(x + kx, 0), w - 1);
                        acc += src[(size_t)yy * w + xx] * (k[ky + 1] * k[kx + 1] / 16.0f);
                    }
                out[(size_t)y * w + x] = acc;
            }
        return out;
    };
    const std::vector<float> fa = blur(ga), fb = blur(gb);
    double ma = 0.0, mb = 0.0;
    for (size_t i = 0; i < n; ++i) { ma += fa[i]; mb += fb[i]; }
    ma /= (double)n; mb /= (double)n;
    double va = 0.0, vb = 0.0, cov = 0.0;
    for (size_t i = 0; i < n; ++i) { const double da = fa[i] - ma, db = fb[i] - mb; va += da * da; vb += db * db; cov += da * db; }
    const double dn = n > 1 ? (double)(n - 1) : 1.0;
    va /= dn; vb /= dn; cov /= dn;
    const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
    const double ssim = ((2.0 * ma * mb + C1) * (2.0 * cov + C2)) / ((ma * ma + mb * mb + C1) * (va + vb + C2));
    r->differentPixels = (int32_t)different; r->totalPixels = (int32_t)n;
    r->pixelDifferenceRatio = (float)((double)different / (double)n);
    r->rmse = (float)std::sqrt(sq / ((double)n * channels));
    r->ssim = (float)ssim;
    r->isIdentical = different == 0 ? 1 : 0;
    r->isVeryClose = (r->ssim > 0.99f && r->rmse < 1.0f) ? 1 : 0;
    r->isClose = (r->ssim > 0.95f && r->rmse < 5.0f) ? 1 : 0;
    return VPT_OK;
}

// minimal RGB8 PNG writer for the difference picture (zlib "stored" blocks)
namespace {
uint32_t crcOf(const uint8_t *d, size_t n, uint32_t crc)
{
    static uint32_t table[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ d[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}
void putBe32(std::vector<uint8_t> &v, uint32_t x) { v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x); }
void pngChunk(std::vector<uint8_t> &out, const char *type, const std::vector<uint8_t> &data)
{
    putBe32(out, (uint32_t)data.size());
    std::vector<uint8_t> td(type, type + 4);
    td.insert(td.end(), data.begin(), data.end());
    out.insert(out.end(), td.begin(), td.end());
    putBe32(out, crcOf(td.data(), td.size(), 0xFFFFFFFFu) ^ 0xFFFFFFFFu);
}
bool writeRgbPng(const char *path, int w, int h, const std::vector<uint8_t> &rgb)
{
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * ((size_t)w * 3 + 1));
    for (int y = 0; y < h; ++y) { raw.push_back(0); raw.insert(raw.end(), rgb.begin() + (size_t)y * w * 3, rgb.begin() + (size_t)(y + 1) * w * 3); }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (uint8_t v : raw) { a = (a + v) % 65521u; b = (b + a) % 65521u; }
    for (size_t pos = 0; pos < raw.size();)
    {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xff)); z.push_back((uint8_t)(n >> 8)); z.push_back((uint8_t)(~n & 0xff)); z.push_back((uint8_t)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        pos += n;
    }
    putBe32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr;
    putBe32(ihdr, (uint32_t)w); putBe32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    pngChunk(out, "IHDR", ihdr); pngChunk(out, "IDAT", z); pngChunk(out, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    if (!f.is_open()) return false;
    f.write((const char *)out.data(), (std::streamsize)out.size());
    return (bool)f;
}
} // namespace

extern "C" int vpt_image_diff_files(const char *pngA, const char *pngB, VptImageDiffResult *r, const char *diffPng)
{
    if (!pngA || !pngB || !r) return VPT_ERR_ARG;
    int wa = 0, ha = 0, ca = 0, wb = 0, hb = 0, cb = 0;
    if (vpt_load_png_rgba8(pngA, nullptr, 0, &wa, &ha, &ca) != VPT_OK || vpt_load_png_rgba8(pngB, nullptr, 0, &wb, &hb, &cb) != VPT_OK) return VPT_ERR_IO;
    if (wa != wb || ha != hb) return VPT_ERR_ARG; // "Cannot compare images of different sizes"
    std::vector<uint32_t> A((size_t)wa * ha), B((size_t)wb * hb);
    if (vpt_load_png_rgba8(pngA, A.data(), A.size(), &wa, &ha, &ca) != VPT_OK || vpt_load_png_rgba8(pngB, B.data(), B.size(), &wb, &hb, &cb) != VPT_OK) return VPT_ERR_IO;
    const int channels = std::min(ca, cb);
    const int rc = vpt_image_diff(A.data(), B.data(), wa, ha, channels, r);
    if (rc != VPT_OK || !diffPng) return rc;
    std::vector<uint8_t> rgb((size_t)wa * ha * 3);
    for (size_t i = 0; i < A.size(); ++i)
    {
        int d[3];
        for (int c = 0; c < 3; ++c) d[c] = std::abs((int)((A[i] >> (8 * c)) & 0xffu) - (int)((B[i] >> (8 * c)) & 0xffu));
        if (channels < 3) { d[1] = d[0]; d[2] = d[0]; }
        for (int c = 0; c < 3; ++c) rgb[i * 3 + c] = (uint8_t)std::min(255.0f, (float)d[c] * 3.0f);
    }
    return writeRgbPng(diffPng, wa, ha, rgb) ? VPT_OK : VPT_ERR_IO;
}

// ---- texture mip chains (TextureManager.cu:82-115, 216-217, 395-411): 2x2 box average per channel, truncated, down to 4x4
static int mipLevels(int width)
{
    if (width <= 0 || (width & (width - 1))) return -1;
    int lg = 0;
    while ((1 << lg) < width) ++lg;
    return lg >= 2 ? lg - 1 : 1;
}
extern "C" int vpt_mip_chain_texels(int width)
{
    const int levels = mipLevels(width);
    if (levels < 0) return 0;
    long long n = 0;
    for (int l = 0; l < levels; ++l) n += (long long)(width >> l) * (width >> l);
    return n > 0x7fffffffLL ? 0 : (int)n;
}
extern "C" int vpt_build_mip_chain(const uint32_t *level0, int width, uint32_t *out)
{
    const int levels = mipLevels(width);
    if (levels < 0 || !level0 || !out) return 0;
    std::memcpy(out, level0, (size_t)width * width * 4);
    const uint32_t *src = out;
    uint32_t *dst = out + (size_t)width * width;
    for (int l = 1; l < levels; ++l)
    {
        const int n = width >> l, sn = n * 2;
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x)
            {
                const uint32_t a = src[(size_t)(2 * y) * sn + 2 * x], b = src[(size_t)(2 * y) * sn + 2 * x + 1];
                const uint32_t c = src[(size_t)(2 * y + 1) * sn + 2 * x], d = src[(size_t)(2 * y + 1) * sn + 2 * x + 1];
                uint32_t v = 0;
                for (int ch = 0; ch < 32; ch += 8) // (float sum) * 0.25 truncated == integer sum >> 2 (sum <= 1020 is exact in fp32)
                    v |= ((((a >> ch) & 0xffu) + ((b >> ch) & 0xffu) + ((c >> ch) & 0xffu) + ((d >> ch) & 0xffu)) >> 2) << ch;
                dst[(size_t)y * n + x] = v;
            }
        src = dst;
        dst += (size_t)n * n;
    }
    return levels;
}

extern "C" void vpt_chunk_hash(const uint8_t *chunk, char *hex17)
{
    unsigned long long hash = 1469598103934665603ull;               // FNV-1a 64 offset basis
    for (size_t i = 0; i < 32768; ++i) { hash ^= (unsigned long long)chunk[i]; hash *= 1099511628211ull; }
    std::snprintf(hex17, 17, "%016llx", hash);
}
static std::string float3ToStringLikeReference(const float *v)
{
    std::ostringstream oss;                                         // SceneConfigParser::float3ToString: "[x, y, z]"
    oss << "[" << v[0] << ", " << v[1] << ", " << v[2] << "]";
    return oss.str();
}
extern "C" int vpt_save_world(const char *sceneYamlPath, const char *chunkDir, int cx, int cy, int cz, const uint8_t *ids, const float *cam9, float fov)
{
    if (!sceneYamlPath || !chunkDir || !ids || !cam9 || cx <= 0 || cy <= 0 || cz <= 0) return VPT_ERR_ARG;
    const int total = cx * cy * cz;
    bool ok = true;
    std::vector<std::string> hashes((size_t)total);
    for (int i = 0; i < total; ++i)
    {
        char hex[17];
        vpt_chunk_hash(ids + (size_t)i * 32768, hex);
        hashes[(size_t)i] = hex;
        std::ofstream out(std::string(chunkDir) + "/" + hex + ".bin", std::ios::binary | std::ios::out | std::ios::trunc);
        if (!out.is_open()) { ok = false; continue; }
        out.write((const char *)(ids + (size_t)i * 32768), 32768);
        if (!out.good()) ok = false;
    }
    std::ofstream file(sceneYamlPath);
    if (!file.is_open()) return VPT_ERR_IO;
    const float zero3[3] = {0.0f, 0.0f, 0.0f}, one3[3] = {1.0f, 1.0f, 1.0f};
    file << "# Scene Configuration File\n# Generated automatically\n\n";
    file << "camera:\n  position: " << float3ToStringLikeReference(cam9) << "\n  direction: " << float3ToStringLikeReference(cam9 + 3)
         << "\n  up: " << float3ToStringLikeReference(cam9 + 6) << "\n  fov: " << fov << "\n\n";
    file << "character:\n  position: " << float3ToStringLikeReference(zero3) << "\n  rotation: " << float3ToStringLikeReference(zero3)
         << "\n  scale: " << float3ToStringLikeReference(one3) << "\n";
    file << "\nchunk_config:\n  chunksX: " << cx << "\n  chunksY: " << cy << "\n  chunksZ: " << cz << "\n";
    file << "\nchunks:\n";
    for (int i = 0; i < total; ++i) file << "  " << i << ": " << hashes[(size_t)i] << "\n";
    return (ok && file.good()) ? VPT_OK : VPT_ERR_IO;
}
extern "C" int vpt_load_world(const char *sceneYamlPath, const char *chunkDir, int cx, int cy, int cz, uint8_t *ids, int *loaded, int *failed)
{
    if (!sceneYamlPath || !chunkDir || !ids || cx <= 0 || cy <= 0 || cz <= 0) return VPT_ERR_ARG;
    const unsigned total = (unsigned)(cx * cy * cz);
    int nOk = 0, nBad = 0;
    const int rc = parseSection(sceneYamlPath, "chunks", [&](const std::string &key, const std::string &value) {
        int index = -1;
        if (!parseInt(key, index) || index < 0) return;
        if ((unsigned)index >= total) { ++nBad; return; }               // "Chunk index out of range in scene file"
        std::ifstream in(std::string(chunkDir) + "/" + value + ".bin", std::ios::binary | std::ios::in | std::ios::ate);
        if (!in.is_open() || (long long)in.tellg() != 32768) { ++nBad; return; }   // missing file / size mismatch: skipped
        in.seekg(0);
        std::vector<char> buf(32768);
        in.read(buf.data(), 32768);
        if (in.gcount() != 32768) { ++nBad; return; }
        std::memcpy(ids + (size_t)index * 32768, buf.data(), 32768);
        ++nOk;
    });
    if (loaded) *loaded = nOk;
    if (failed) *failed = nBad;
    return rc;
}

// Test hook: the magic-number division used by the kernels (csrc/vpt_fastdiv.h), evaluated on the host.
extern "C" void vpt_debug_fastdiv(uint32_t n, uint32_t d, uint32_t *q, uint32_t *r)
{
    const vpt::FastDiv f = vpt::makeFastDiv(d);
    f.div(n, *q, *r);
}
