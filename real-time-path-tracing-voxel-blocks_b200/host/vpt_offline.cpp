// vpt_offline — mainOffline-compatible offline render entry over the C ABI (include/vpt.h).
//
// Mirrors /root/reference/mainOffline.cpp:29-512 for the hot path: same CLI flags (:57-133), the same settings and
// scene files (data/settings/global_settings.yaml via GlobalSettings::LoadFromYAML, data/scene/scene_export.yaml via
// SceneConfigParser::LoadFromFile), the same frame loop (historyCamera = camera; camera.update(); renderFrame;
// frames 1,4,16,64 saved as <prefix>_<frame:04>.png) and the PNG writer of OfflineBackend::writeFrameBufferToPNG
// (renderer/core/OfflineBackend.cpp:191-221: y-flip, clamp to [0,1], *255). Out of scope here, as in SURVEY §2:
// the wall-clock / history dependent post effects (auto-exposure, bloom, lens flare, vignette); the deterministic part —
// FilmicToneMapping with the manual exposure of the settings file, sRGB, the PNG conversion — runs on the device (vpt_tonemap),
// (the canonical comparison --test-canonical / --update-canonical / --canonical-image runs the reference's ImageDiff criteria
// through vpt_image_diff_files; the golden PNG itself is absent from the reference tree). The scripted edit tests (--test-sequence /
// --test-remove20 / --test-remove-circle, mainOffline.cpp:168-188, 279-393) run through the picker: vpt_pick_voxel + vpt_set_voxel.
// New flags of this build: --spp N, --bounces T D, --chunks X Y Z, --exposure E, --tables PATH, --sky-tables PATH.
#include "../../include/vpt.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <string>
#include <vector>

namespace {

// ---- minimal PNG writer (zlib "stored" blocks; no external dependency)
uint32_t crcTable[256];
void initCrc()
{
    for (uint32_t n = 0; n < 256; ++n)
    {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crcTable[n] = c;
    }
}
uint32_t crc32(const uint8_t *p, size_t n, uint32_t c = 0xFFFFFFFFu)
{
    for (size_t i = 0; i < n; ++i) c = crcTable[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c;
}
void put32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::vector<uint8_t> &out, const char *type, const std::vector<uint8_t> &data)
{
    put32(out, (uint32_t)data.size());
    std::vector<uint8_t> td(type, type + 4);
    td.insert(td.end(), data.begin(), data.end());
    out.insert(out.end(), td.begin(), td.end());
    put32(out, crc32(td.data(), td.size()) ^ 0xFFFFFFFFu);
}
bool writePng(const std::string &path, int w, int h, const std::vector<uint8_t> &rgb)
{
    initCrc();
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * (w * 3 + 1));
    for (int y = 0; y < h; ++y)
    {
        raw.push_back(0);
        raw.insert(raw.end(), rgb.begin() + (size_t)y * w * 3, rgb.begin() + (size_t)(y + 1) * w * 3);
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    size_t pos = 0;
    uint32_t a = 1, b = 0;
    for (uint8_t v : raw) { a = (a + v) % 65521; b = (b + a) % 65521; }
    while (pos < raw.size())
    {
        size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back(n & 0xFF); z.push_back(n >> 8); z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        pos += n;
    }
    put32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put32(ihdr, w); put32(ihdr, h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    f.write((const char *)out.data(), out.size());
    return (bool)f;
}

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != VPT_OK) { std::fprintf(stderr, "Error: %s -> %s\n", #call, vpt_last_error()); return 1; } \
    } while (0)

struct MaterialRow { float r, g, b, rough; };

} // namespace

int main(int argc, char *argv[])
{
    int width = 3840, height = 2160;
    std::string outputPrefix = "offline_render", sceneFile = "data/scene/scene_export.yaml";
    std::string settingsFile = "data/settings/global_settings.yaml", tablesFile = "data/bluenoise_tables.bin", skyTablesFile = "data/sky_tables.bin";
    std::string assetsDir, dataRoot = "."; // --assets: materials.yaml/blocks.yaml directory (reference: data/assets); texture paths resolve against --data-root
    bool useTextures = true;
    bool testCanonical = false, updateCanonical = false; // mainOffline.cpp:41-42, 422-497
    // PerformanceTracker::saveReport (renderer/util/PerformanceTracker.h:98-175): one summary line per run, appended. The reference
    // always writes ../../data/perf/performance_report.txt; here only when --perf-report names the file.
    std::string runComment = "default run", perfReportPath;
    std::string canonicalImagePath = "../../data/canonical/canonical_render.png"; // the reference's default (mainOffline.cpp:41)
    // scripted block edits (mainOffline.cpp:43-51, 168-188, 279-393): clicks are consumed by the next frame's VoxelEngine::update
    bool enableTestSequence = false, enableRemovalStressTest = false, enableCircularRemovalTest = false;
    const int removalTestClickCount = 20, circularTestViewDirections = 8, circularTestRemovalsPerDirection = 5;
    const int circularTestTotalRemovals = circularTestViewDirections * circularTestRemovalsPerDirection;
    const float circularYawAmplitude = 12.0f * 3.14159265358979323846f / 180.0f, circularPitchAmplitude = 6.0f * 3.14159265358979323846f / 180.0f;
    std::string worldChunkDir, saveWorldDir; // WorldSceneManager chunk storage: load the scene's "chunks:" records / save the world
    int totalFrames = 64;
    std::vector<int> savedFrames = {1, 4, 16, 64};
    int spp = 1, totalBounce = 3, diffuseBounce = 1;
    int chunks[3] = {2, 1, 2}; // ChunkConfiguration default (voxelengine/VoxelSceneGen.h:10-20)
    bool chunksFromCli = false;
    float exposure = -1.0f; // --exposure overrides postprocess.manualExposure of the settings file
    for (int i = 1; i < argc; i++)
    {
        std::string arg = argv[i];
        if (arg == "--width" && i + 1 < argc) width = std::atoi(argv[++i]);
        else if (arg == "--height" && i + 1 < argc) height = std::atoi(argv[++i]);
        else if (arg == "--output" && i + 1 < argc) outputPrefix = argv[++i];
        else if (arg == "--scene" && i + 1 < argc) sceneFile = argv[++i];
        else if (arg == "--settings" && i + 1 < argc) settingsFile = argv[++i];
        else if (arg == "--tables" && i + 1 < argc) tablesFile = argv[++i];
        else if (arg == "--sky-tables" && i + 1 < argc) skyTablesFile = argv[++i];
        else if (arg == "--assets" && i + 1 < argc) assetsDir = argv[++i];
        else if (arg == "--data-root" && i + 1 < argc) dataRoot = argv[++i];
        else if (arg == "--no-textures") useTextures = false;
        else if (arg == "--world-chunks" && i + 1 < argc) worldChunkDir = argv[++i];
        else if (arg == "--save-world" && i + 1 < argc) saveWorldDir = argv[++i];
        else if (arg == "--test-canonical" || arg == "--test") testCanonical = true;
        else if (arg == "--update-canonical") updateCanonical = true;
        else if (arg == "--canonical-image" && i + 1 < argc) canonicalImagePath = argv[++i];
        else if (arg == "--comment" && i + 1 < argc) runComment = argv[++i];
        else if (arg == "--perf-report" && i + 1 < argc) perfReportPath = argv[++i];
        else if (arg == "--test-sequence") enableTestSequence = true;
        else if (arg == "--test-remove20") enableRemovalStressTest = true;
        else if (arg == "--test-remove-circle") enableCircularRemovalTest = true;
        else if (arg == "--frames" && i + 1 < argc)
        {
            totalFrames = std::atoi(argv[++i]);
            if (totalFrames == 1) savedFrames = {1};
        }
        else if (arg == "--spp" && i + 1 < argc) spp = std::atoi(argv[++i]);
        else if (arg == "--bounces" && i + 2 < argc) { totalBounce = std::atoi(argv[++i]); diffuseBounce = std::atoi(argv[++i]); }
        else if (arg == "--chunks" && i + 3 < argc) { chunks[0] = std::atoi(argv[++i]); chunks[1] = std::atoi(argv[++i]); chunks[2] = std::atoi(argv[++i]); chunksFromCli = true; }
        else if (arg == "--exposure" && i + 1 < argc) exposure = (float)std::atof(argv[++i]);
        else if (arg == "--help" || arg == "-h")
        {
            std::printf("Offline Voxel Path Tracer (B200-native hot path)\nUsage: %s [options]\n"
                        "  --width <int> --height <int> --output <prefix> --scene <file> --frames <int>\n"
                        "  --test-canonical --update-canonical --canonical-image <path> --comment <text>\n"
                        "  --test-sequence --test-remove20 --test-remove-circle   scripted block edits through the picker (vpt_pick_voxel + vpt_set_voxel)\n"
                        "  --spp <int> --bounces <total> <diffuse> --chunks <x> <y> <z> --exposure <f> --settings <file> --tables <file> --sky-tables <file>\n"
                        "  --perf-report <file>  append the run summary line of PerformanceTracker::saveReport (with --comment)\n"
                        "  --assets <dir>  materials.yaml + blocks.yaml (reference: data/assets) with their textures under --data-root <dir> (default .); --no-textures\n"
                        "  --world-chunks <dir>  load the chunk files the scene lists (WorldSceneManager::LoadScene)   --save-world <dir>  write them\n", argv[0]);
            return 0;
        }
    }
    std::printf("=== Offline Voxel Path Tracer ===\nResolution: %dx%d, frames %d, spp %d, bounces %d/%d\n", width, height, totalFrames, spp, totalBounce, diffuseBounce);

    VptDenoisingParams dn;
    vpt_default_denoising_params(&dn);
    if (vpt_load_denoising_settings(settingsFile.c_str(), &dn) != VPT_OK)
        std::fprintf(stderr, "Failed to open global settings file: %s (defaults kept)\n", settingsFile.c_str());
    VptToneMappingParams tone;
    vpt_default_tonemapping_params(&tone);
    vpt_load_tonemapping_settings(settingsFile.c_str(), &tone);
    if (exposure > 0.0f) tone.manualExposure = exposure;
    float cam9[9], fov; unsigned sceneChunks[3];
    if (vpt_load_scene_config(sceneFile.c_str(), cam9, &fov, sceneChunks) != VPT_OK)
        std::printf("Scene file not found: %s, using defaults\n", sceneFile.c_str());
    if (!chunksFromCli && sceneChunks[0] && sceneChunks[1] && sceneChunks[2]) { chunks[0] = sceneChunks[0]; chunks[1] = sceneChunks[1]; chunks[2] = sceneChunks[2]; }

    std::vector<uint8_t> tables(327680);
    {
        std::ifstream f(tablesFile, std::ios::binary);
        if (!f.read((char *)tables.data(), tables.size())) { std::fprintf(stderr, "Error: cannot read blue-noise tables %s\n", tablesFile.c_str()); return 1; }
    }
    vpt_ctx *ctx = nullptr;
    CHECK(vpt_create(0, width, height, &ctx));
    CHECK(vpt_set_tables(ctx, tables.data(), tables.data() + 65536, tables.data() + 196608));
    // VoxelEngine::init -> initVoxelsMultiChunk per chunk (voxelengine/VoxelEngine.cu:754-820)
    std::vector<float> noise((size_t)chunks[0] * chunks[1] * chunks[2] * 1024);
    vpt_perlin_noise_chunks(chunks[0], chunks[1], chunks[2], 124, noise.data());
    CHECK(vpt_generate_terrain(ctx, chunks[0], chunks[1], chunks[2], noise.data()));
    if (!worldChunkDir.empty() || !saveWorldDir.empty())
    {
        // WorldSceneManager::LoadScene / SaveScene, chunk part (renderer/core/WorldSceneManager.cpp:310-458)
        std::vector<uint8_t> ids((size_t)chunks[0] * chunks[1] * chunks[2] * 32768);
        CHECK(vpt_get_grid(ctx, ids.data(), ids.size()));
        if (!worldChunkDir.empty())
        {
            int loaded = 0, failed = 0;
            if (vpt_load_world(sceneFile.c_str(), worldChunkDir.c_str(), chunks[0], chunks[1], chunks[2], ids.data(), &loaded, &failed) != VPT_OK)
                std::printf("Scene file not found: %s (no chunks loaded)\n", sceneFile.c_str());
            std::printf("World chunks: %d loaded, %d failed from %s\n", loaded, failed, worldChunkDir.c_str());
            if (loaded > 0) CHECK(vpt_set_grid(ctx, chunks[0], chunks[1], chunks[2], ids.data()));
        }
        if (!saveWorldDir.empty())
        {
            const std::string sceneOut = saveWorldDir + "/scene.yaml";
            if (vpt_save_world(sceneOut.c_str(), saveWorldDir.c_str(), chunks[0], chunks[1], chunks[2], ids.data(), cam9, fov) != VPT_OK)
                std::fprintf(stderr, "Failed to save the world to %s\n", saveWorldDir.c_str());
            else std::printf("Saved scene config to: %s\n", sceneOut.c_str());
        }
    }
    // terrain materials (data/assets/materials.yaml order; textures out of scope -> flat albedo stand-ins)
    const MaterialRow rows[12] = {{0.76f, 0.70f, 0.50f, 0.8f}, {0.45f, 0.33f, 0.22f, 0.9f}, {0.55f, 0.52f, 0.48f, 0.85f}, {0.40f, 0.30f, 0.20f, 0.9f},
                                  {0.76f, 0.70f, 0.50f, 0.8f}, {0.80f, 0.75f, 0.60f, 0.7f}, {0.50f, 0.50f, 0.52f, 0.85f}, {0.60f, 0.60f, 0.58f, 0.6f},
                                  {0.62f, 0.58f, 0.52f, 0.7f}, {0.78f, 0.74f, 0.66f, 0.65f}, {0.55f, 0.40f, 0.25f, 0.75f}, {0.55f, 0.40f, 0.25f, 0.75f}};
    VptMaterial mats[12];
    std::memset(mats, 0, sizeof mats);
    for (int i = 0; i < 12; ++i)
    {
        mats[i].albedo[0] = rows[i].r; mats[i].albedo[1] = rows[i].g; mats[i].albedo[2] = rows[i].b;
        mats[i].roughness = rows[i].rough; mats[i].uvScale = 2.5f; mats[i].useWorldGridUV = 1; mats[i].materialId = i;
    }
    uint16_t b2m[256] = {0};
    for (int b = 1; b <= 12; ++b) b2m[b] = (uint16_t)(b - 1);
    if (assetsDir.empty())
    {
        // the reference always loads data/assets relative to the working directory (AssetRegistry::loadFromYAML): do the same when it is there
        std::ifstream probe("data/assets/materials.yaml");
        if (probe.is_open()) { assetsDir = "data/assets"; if (dataRoot == ".") dataRoot = "data"; }
    }
    if (assetsDir.empty()) CHECK(vpt_set_materials(ctx, mats, 12, b2m));
    else
    {
        // AssetRegistry + MaterialManager::init + TextureManager::init (renderer/assets/): the reference's own tables and textures
        std::vector<VptMaterial> am(256);
        std::vector<VptMaterialTexturePaths> ap(256);
        int nm = 0;
        if (vpt_load_materials((assetsDir + "/materials.yaml").c_str(), (assetsDir + "/blocks.yaml").c_str(), am.data(), ap.data(), 256, &nm, b2m) != VPT_OK)
        { std::fprintf(stderr, "Error: cannot load %s/materials.yaml + blocks.yaml\n", assetsDir.c_str()); return 1; }
        CHECK(vpt_set_materials(ctx, am.data(), nm, b2m));
        std::printf("Materials: %d from %s\n", nm, assetsDir.c_str());
        if (useTextures)
        {
            std::vector<std::string> names;            // unique texture files, first use first
            std::vector<int32_t> slots((size_t)nm * 4, -1), widths, levels;
            std::vector<float> texSize((size_t)nm * 2, 1024.0f); // MaterialParameter::texSize default (SystemParameter.h:29)
            std::vector<uint32_t> texels;
            for (int m = 0; m < nm; ++m)
            {
                const char *pp[4] = {ap[(size_t)m].albedo, ap[(size_t)m].normal, ap[(size_t)m].roughness, ap[(size_t)m].metallic};
                for (int k = 0; k < 4; ++k)
                {
                    if (!pp[k][0]) continue;
                    const std::string file = dataRoot + "/" + (std::strncmp(pp[k], "data/", 5) == 0 ? pp[k] + 5 : pp[k]);
                    size_t t = 0;
                    while (t < names.size() && names[t] != file) ++t;
                    if (t == names.size())
                    {
                        int w = 0, h = 0, ch = 0;
                        if (vpt_load_png_rgba8(file.c_str(), nullptr, 0, &w, &h, &ch) != VPT_OK || w != h || vpt_mip_chain_texels(w) == 0)
                        { std::fprintf(stderr, "Warning: texture %s missing or not a square power of two, slot left empty\n", file.c_str()); continue; }
                        std::vector<uint32_t> img((size_t)w * h);
                        if (vpt_load_png_rgba8(file.c_str(), img.data(), img.size(), &w, &h, &ch) != VPT_OK) continue;
                        const size_t base = texels.size();
                        texels.resize(base + (size_t)vpt_mip_chain_texels(w));
                        levels.push_back(vpt_build_mip_chain(img.data(), w, texels.data() + base));
                        widths.push_back(w);
                        names.push_back(file);
                    }
                    slots[(size_t)m * 4 + k] = (int32_t)t;
                }
            }
            if (!names.empty())
            {
                CHECK(vpt_set_textures(ctx, (int)names.size(), widths.data(), levels.data(), texels.data(), nm, slots.data(), texSize.data()));
                std::printf("Textures: %zu files, %.1f MB of RGBA8 mip chains\n", names.size(), texels.size() * 4.0 / 1048576.0);
            }
        }
    }
    // SkyModel::init/update (mainOffline.cpp:191 -> OfflineBackend::init -> SkyModel; Sky.cu:355-396): Hosek-Wilkie sky + solar
    // disc evaluated on the device from the "sky" section of the settings file
    {
        std::vector<float> skyTables(2460);
        std::ifstream f(skyTablesFile, std::ios::binary);
        if (!f.read((char *)skyTables.data(), skyTables.size() * sizeof(float))) { std::fprintf(stderr, "Error: cannot read sky tables %s\n", skyTablesFile.c_str()); return 1; }
        VptSkyParams sky = {0.25f, 45.0f, 0.0f, 1.0f};
        vpt_load_sky_settings(settingsFile.c_str(), &sky);
        CHECK(vpt_generate_sky(ctx, &sky, skyTables.data()));
    }
    CHECK(vpt_set_trace_params(ctx, spp, totalBounce, diffuseBounce, 1));

    // camera from the scene (mainOffline.cpp:227-247)
    VptCamera camera, historyCamera;
    vpt_camera_from_scene(&camera, width, height, cam9, cam9 + 3, fov);
    historyCamera = camera;
    std::printf("Camera setup - Position: (%g, %g, %g)\nCamera setup - Direction: (%g, %g, %g)\nCamera setup - FOV: %g degrees\n",
                camera.pos[0], camera.pos[1], camera.pos[2], camera.dir[0], camera.dir[1], camera.dir[2], fov);

    // scripted click sequences (VoxelEngine::configureOfflineClickSequence): removal tests click block id 0, the placement test
    // cycles 16, 0, 16 (VoxelEngine.cu:906-945). Block 16 is an instanced lantern mesh in the reference: here it is a plain voxel.
    std::vector<int> clickSequence;
    if (enableCircularRemovalTest) { clickSequence.assign((size_t)circularTestTotalRemovals, 0); std::printf("Offline circular removal test enabled: %d view directions, %d deletions each.\n", circularTestViewDirections, circularTestRemovalsPerDirection); }
    else if (enableRemovalStressTest) { clickSequence.assign((size_t)removalTestClickCount, 0); std::printf("Offline removal stress test enabled: %d scripted deletions.\n", removalTestClickCount); }
    size_t clickIndex = 0, defaultClickIndex = 0;
    bool clickPending = false;
    int removalClickCounter = 0, circularRemovalsPerformed = 0, lastCircularDirectionIndex = -1, blocksRemoved = 0, blocksPlaced = 0;
    bool circularOrientationReset = false;
    const float baseCameraYaw = camera.yaw, baseCameraPitch = camera.pitch;

    int iterationIndex = 0; // GlobalSettings::iterationIndex, reset for a fresh offline run (mainOffline.cpp:252)
    double traceMs = 0, denoiseMs = 0;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<double> frameWallMs; double lastWallMs = 0.0;
    for (int f = 0; f < totalFrames; ++f)
    {
        const int frameNumber = f + 1;
        historyCamera = camera;
        if (enableCircularRemovalTest)
        {
            if (circularRemovalsPerformed < circularTestTotalRemovals)
            {
                const int directionIndex = circularRemovalsPerformed / circularTestRemovalsPerDirection;
                if (directionIndex != lastCircularDirectionIndex)
                {
                    const float angle = (float)directionIndex * (6.28318530717958647692f / (float)circularTestViewDirections);
                    const float yawOffset = (float)(circularYawAmplitude * std::cos(angle)), pitchOffset = (float)(circularPitchAmplitude * std::sin(angle));
                    camera.yaw = baseCameraYaw + yawOffset; camera.pitch = baseCameraPitch + pitchOffset;
                    std::printf("CIRCULAR TEST: Switching to view direction #%d (yaw offset %g, pitch offset %g)\n", directionIndex + 1, yawOffset, pitchOffset);
                    lastCircularDirectionIndex = directionIndex;
                }
            }
            else if (!circularOrientationReset)
            {
                camera.yaw = baseCameraYaw; camera.pitch = baseCameraPitch; circularOrientationReset = true;
                std::printf("CIRCULAR TEST: Restored base camera orientation after scripted removals.\n");
            }
        }
        vpt_camera_update(&camera);
        // VoxelEngine::update (OfflineBackend::renderFrame -> VoxelEngine.cu:876-985): pick along the camera ray, apply a pending click
        if (clickPending)
        {
            clickPending = false;
            int blockId;
            if (!clickSequence.empty()) { const size_t i = std::min(clickIndex, clickSequence.size() - 1); blockId = clickSequence[i]; if (i + 1 < clickSequence.size()) clickIndex = i + 1; else clickIndex = i; }
            else { static const int defaultSequence[3] = {16, 0, 16}; blockId = defaultSequence[defaultClickIndex % 3]; ++defaultClickIndex; }
            VptPickResult pick;
            CHECK(vpt_pick_voxel(ctx, camera.pos, camera.dir, &pick));
            if (pick.hitSurface) std::printf("CAMERA RAY DEBUG: Hit block at (%d,%d,%d)\n", pick.deletePos[0], pick.deletePos[1], pick.deletePos[2]);
            if (blockId == 0) { if (pick.hitSurface) { CHECK(vpt_set_voxel(ctx, pick.deletePos[0], pick.deletePos[1], pick.deletePos[2], 0)); ++blocksRemoved; } }
            else if (pick.hasSpaceToCreate && pick.hitSurface) { CHECK(vpt_set_voxel(ctx, pick.createPos[0], pick.createPos[1], pick.createPos[2], blockId)); ++blocksPlaced; }
        }
        CHECK(vpt_render(ctx, &camera, &historyCamera, iterationIndex)); // render() post-increments the index
        ++iterationIndex;
        CHECK(vpt_denoise(ctx, &dn, &camera, &historyCamera, f, iterationIndex));
        VptTimings tm;
        CHECK(vpt_get_timings(ctx, &tm));
        traceMs += tm.trace_ms + tm.resolve_ms; denoiseMs += tm.denoise_total_ms;
        {
            const double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            frameWallMs.push_back(now - lastWallMs); lastWallMs = now;
        }
        if (std::find(savedFrames.begin(), savedFrames.end(), frameNumber) != savedFrames.end())
        {
            // PostProcessor::run's deterministic part (FilmicToneMapping with the manual exposure) + the PNG conversion of
            // OfflineBackend::writeFrameBufferToPNG, on the device
            std::vector<uint8_t> rgb((size_t)width * height * 3);
            CHECK(vpt_tonemap(ctx, &tone, rgb.data(), nullptr));
            char name[512];
            std::snprintf(name, sizeof name, "%s_%04d.png", outputPrefix.c_str(), f);
            if (!writePng(name, width, height, rgb)) std::fprintf(stderr, "Failed to save image to: %s\n", name);
            else std::printf("Saved frame %d/%d -> %s\n", frameNumber, totalFrames, name);
        }
        if (enableCircularRemovalTest)
        {
            if (circularRemovalsPerformed < circularTestTotalRemovals)
            {
                ++circularRemovalsPerformed;
                std::printf("CIRCULAR TEST: Frame %d deleting block #%d (direction %d, ID=0)...\n", frameNumber, circularRemovalsPerformed,
                            std::min(circularTestViewDirections - 1, (circularRemovalsPerformed - 1) / circularTestRemovalsPerDirection) + 1);
                clickPending = true;
            }
        }
        else if (enableRemovalStressTest)
        {
            if (removalClickCounter < removalTestClickCount)
            {
                ++removalClickCounter;
                std::printf("REMOVAL TEST: Frame %d deleting block #%d (ID=0)...\n", frameNumber, removalClickCounter);
                clickPending = true;
            }
        }
        else if (enableTestSequence && (frameNumber == 2 || frameNumber == 5 || frameNumber == 8))
        {
            std::printf("TEST FRAME %d: %s light block...\n", frameNumber, frameNumber == 5 ? "Removing" : "Placing");
            clickPending = true;
        }
    }
    if (enableCircularRemovalTest || enableRemovalStressTest || enableTestSequence)
        std::printf("Scripted edits: %d blocks removed, %d placed\n", blocksRemoved, blocksPlaced);
    const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::printf("Rendering completed successfully!\nAverage per frame: path trace %.3f ms, denoiser %.3f ms (device), whole %.3f ms (wall)\n",
                traceMs / totalFrames, denoiseMs / totalFrames, wall / totalFrames);
    if (!perfReportPath.empty() && !frameWallMs.empty())
    {
        std::ifstream probe(perfReportPath);
        const bool fresh = !probe.is_open() || probe.peek() == std::ifstream::traits_type::eof();
        probe.close();
        std::ofstream rep(perfReportPath, std::ios::app);
        if (rep.is_open())
        {
            if (fresh)
                rep << "# Performance Report - Real-time Path Tracing Voxel Renderer (Run Summary)\n"
                    << "# Format: Timestamp         | Frames | Resolution | WholeFrame | StdDev | ScenePrep | RendererUpd | PathTrace | Denoiser | PostProc | Comment\n"
                    << "# ===============================================================================================================================\n";
            double avg = 0.0, var = 0.0;
            for (double v : frameWallMs) avg += v;
            avg /= (double)frameWallMs.size();
            for (double v : frameWallMs) var += (v - avg) * (v - avg);
            const double sd = std::sqrt(var / (double)frameWallMs.size());
            char stamp[32]; const std::time_t tt = std::time(nullptr); std::strftime(stamp, sizeof stamp, "%Y-%m-%d %H:%M:%S", std::localtime(&tt));
            char line[512];
            const std::string res = std::to_string(width) + "x" + std::to_string(height);
            std::snprintf(line, sizeof line, "%-19s | %6zu | %-10s | %10.2f | %6.2f | %9.2f | %11.2f | %9.2f | %8.2f | %8.2f | %s\n", stamp, frameWallMs.size(), res.c_str(), avg, sd,
                          0.0, 0.0, traceMs / totalFrames, denoiseMs / totalFrames, 0.0, runComment.c_str());
            rep << line;
            std::printf("Performance data saved to: %s\n", perfReportPath.c_str());
        }
    }
    vpt_destroy(ctx);
    // canonical image testing / updating on the last frame (mainOffline.cpp:422-497; ImageDiff::compare + generateDiffImage)
    if (testCanonical || updateCanonical)
    {
        char last[512];
        std::snprintf(last, sizeof last, "%s_%04d.png", outputPrefix.c_str(), totalFrames - 1);
        if (updateCanonical)
        {
            std::printf("\n=== Updating Canonical Image ===\n");
            // the destination is opened (= truncated) only once the source is known to be readable
            std::ifstream src(last, std::ios::binary);
            bool ok = src.is_open();
            if (ok)
            {
                std::ofstream dst(canonicalImagePath, std::ios::binary);
                ok = dst.is_open() && (dst << src.rdbuf());
            }
            if (ok) std::printf("Canonical image updated: %s\n", canonicalImagePath.c_str());
            else std::fprintf(stderr, "Failed to update canonical image\n");
        }
        if (testCanonical)
        {
            std::printf("\n=== Canonical Image Testing ===\n");
            std::ifstream canon(canonicalImagePath, std::ios::binary), test(last, std::ios::binary);
            if (!canon.is_open()) std::printf("Warning: Canonical image not found at %s\nUse --update-canonical to create it from current render\n", canonicalImagePath.c_str());
            else if (!test.is_open()) std::fprintf(stderr, "Error: Test image not found at %s\n", last);
            else
            {
                std::printf("Comparing: %s vs %s\n", last, canonicalImagePath.c_str());
                VptImageDiffResult r;
                const std::string diffPath = outputPrefix + "_diff.png";
                const int rc = vpt_image_diff_files(last, canonicalImagePath.c_str(), &r, diffPath.c_str());
                if (rc != VPT_OK) std::fprintf(stderr, "Failed to compare the images (different sizes or unreadable)\n");
                else
                {
                    std::printf("=== Image Comparison Results ===\nDifferent pixels: %d / %d (%.2f%%)\nRMSE: %.4f\nSSIM: %.6f\n", r.differentPixels, r.totalPixels,
                                r.pixelDifferenceRatio * 100.0f, r.rmse, r.ssim);
                    std::printf("Assessment: %s\n", r.isIdentical ? "IDENTICAL" : r.isVeryClose ? "VERY CLOSE (excellent match)" : r.isClose ? "CLOSE (good match)"
                                                                                                                               : "DIFFERENT (significant differences detected)");
                    std::printf("Difference visualization saved to: %s\n", diffPath.c_str());
                    if (!r.isIdentical && !r.isVeryClose)
                    {
                        std::printf("\nWarning: Significant differences detected from canonical image!\nThis may indicate a regression or intentional change.\n");
                        if (!r.isClose) std::printf("Consider investigating the differences.\n");
                    }
                    else std::printf("\nImage matches canonical reference within acceptable tolerance.\n");
                }
            }
        }
    }
    return 0;
}
