// Shared-memory tile versions of the spatial denoiser passes for sm_100a (north_star: "bandwidth-bound stencil kernels with
// TMA/shared-memory halo tiles"). Same arithmetic as the gather kernels of vpt_denoise.cu — those stay as the general path
// (any a-trous step, any image width) and as the A/B reference (VPT_DN_GATHER=1) — but the data path is different:
//
//   atrousTileKernel<STEP,TY>   Atrous (Atrous.h:6-158), steps 2 / 4 / 8: ONE elected thread issues three
//                               cp.async.bulk.tensor.2d copies (radiance, packed G-buffer, material word) of the (32 + 2h) x
//                               (TY + 2h) box around the CTA's 32 x TY pixels, h = step (+ step/4 of hashed jitter for step 8),
//                               and arms an mbarrier with the byte count; the 256 threads wait on the barrier and then take
//                               every tap from shared memory at compile-time offsets. Out-of-image taps are zero-filled by the
//                               TMA unit and carry zero weight (the `inside` test), so interior tiles have no bounds logic at all.
//   atrousFirstTileKernel       AtrousSmem (AtrousSmem.h:66-303): 36 x 20 box (halo 2 for the 5x5 young-history branch); the
//                               reference's clamp-to-edge addressing is restored on border tiles by patching the zero-filled
//                               entries in shared memory from their clamped neighbours.
//   historyClampColKernel       HistoryClamping (HistoryClamping.h:27-219): a warp owns 32 columns and WALKS DOWN the rows; every
//                               pixel's colour transforms are evaluated once and kept in a 5-row register ring, the 5x5 moments
//                               are a vertical sum over the ring and a horizontal sum over lane shuffles — no shared-memory round
//                               trips (the tile version spent 55 % of its time in shared-memory wavefronts, ncu r1l).
//
// Why tiles: the gather kernels issue 27 dependent-address LDGs per pixel through L1 (long-scoreboard 7.7-11.6 stalled cycles per
// issue, ncu r1l) and spend ~25 % of their instructions on clamping / index arithmetic. With the box staged by the copy engine the
// SM issues only LDS at immediate offsets, and 2-4 resident CTAs per SM overlap one tile's transfer with another's arithmetic.
#include "vpt_denoise_common.cuh"
#include "vpt_tma.cuh"
#include <cstdlib>
#include <mutex>
#include <vector>

namespace vpt {

// ------------------------------------------------------------------------------------------------ host: tensor maps
namespace tma {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encodeFn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}
cudaError_t encode2D(CUtensorMap *map, bool asUint32, const void *base, uint64_t dimX, uint64_t dimY, uint64_t rowPitchBytes, uint32_t boxX, uint32_t boxY)
{
    EncodeTiledFn fn = encodeFn();
    if (!fn) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {dimX, dimY};
    const cuuint64_t strides[1] = {rowPitchBytes};
    const cuuint32_t box[2] = {boxX, boxY};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, asUint32 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
} // namespace tma

namespace {
// Encoding a map is a few hundred host nanoseconds of bit packing, but the planes of a context are few and fixed: cache by value.
struct MapKey { const void *base; int W, H, boxW, boxH, elemFloats; };
struct MapEntry { MapKey k; CUtensorMap m; };
std::mutex g_mapMutex;
std::vector<MapEntry> g_maps;
// plane of W x H pixels, elemFloats 4-byte words per pixel; box of boxW x boxH pixels
cudaError_t planeMap(CUtensorMap *out, const void *base, int W, int H, int boxW, int boxH, int elemFloats, bool asUint32)
{
    std::lock_guard<std::mutex> lock(g_mapMutex);
    for (const MapEntry &e : g_maps)
        if (e.k.base == base && e.k.W == W && e.k.H == H && e.k.boxW == boxW && e.k.boxH == boxH && e.k.elemFloats == elemFloats) { *out = e.m; return cudaSuccess; }
    MapEntry e;
    e.k = {base, W, H, boxW, boxH, elemFloats};
    const cudaError_t r = tma::encode2D(&e.m, asUint32, base, (uint64_t)W * elemFloats, (uint64_t)H, (uint64_t)W * elemFloats * 4, (uint32_t)(boxW * elemFloats), (uint32_t)boxH);
    if (r != cudaSuccess) return r;
    if (g_maps.size() > 256) g_maps.clear(); // contexts come and go in the test-suite; the cache never grows without bound
    g_maps.push_back(e);
    *out = e.m;
    return cudaSuccess;
}
} // namespace

__device__ unsigned g_tmaTimeouts = 0; // tile loads whose mbarrier never completed (always 0 in a healthy build)
unsigned debugTmaTimeouts()
{
    unsigned v = 0;
    cudaMemcpyFromSymbol(&v, g_tmaTimeouts, sizeof v);
    return v;
}

// ------------------------------------------------------------------------------------------------ a-trous over a TMA tile
template <int STEP> struct AtrousTile
{
    static constexpr int kHalo = STEP + (STEP > 4 ? STEP / 4 : 0); // hashed jitter: |offset| <= step / 4 (Atrous.h:60-75)
    // The first element of a TMA box must sit on a 16-byte boundary of the plane (a misaligned origin faults as "illegal
    // instruction", tools/tma_probe.cu): any pixel of a float4 plane does, the 4-byte material plane only every 4th pixel, so its
    // box starts at x0 - kHaloQ with the halo rounded up to a multiple of 4 (x0 is a multiple of 32).
    static constexpr int kHaloQ = (kHalo + 3) & ~3;
};
#ifndef VPT_ATILE_MINB
#define VPT_ATILE_MINB 4 // measured on B200 (three passes, 1080p / 4K): 3 -> 142 / 423 us, 4 (64 registers) -> 128 / 380 us
#endif

// One pixel of the pass; ci / cq = its entry in the float4 tiles / in the material tile. The arithmetic is atrousBody's
// (vpt_denoise.cu) through the shared atrousTap, so both data paths give the same image bit for bit.
template <int STEP, int BW, int BWQ, bool kComposite, bool kInterior>
VPT_DEV void atrousTilePixel(const AtrousArgs &a, const float4 *__restrict__ sIn, const float4 *__restrict__ sG, const uint32_t *__restrict__ sMQ, int x, int y,
                             int ci, int cq, float hl, float4 al)
{
    const int W = a.W, H = a.H;
    const int pix = y * W + x;
    const float4 g = sG[ci];
    if (g.w > kSkyZs) return;
    const uint32_t cMat = sMQ[cq] & 0xffffu;
    const f4 cv = F4(sIn[ci]);
    const f3 cn = {g.x, g.y, g.z};
    constexpr int stepSize = STEP;
    const float cLum = luminance(xyz(cv));
    const float phiInv = 1.0f / fmaxr(1.0e-4f, a.phiLuminance * sqrtf(cv.w));
    float nParam = a.nParamFull;
    if (hl < 5.0f)
    {
        float lobeFrac = a.lobeAngleFraction / sqrtf((float)stepSize);
        lobeFrac = lerpf(0.99f, lobeFrac, saturate(hl / 5.0f));
        nParam = normalWeightParam2(1.0f, lobeFrac);
    }
    const PlaneTest pt = planeTest(a.view, x, y, cn, g.w, a.depthThreshold);
    float sumW = 0.44198f * 0.44198f;
    f4 sum = cv * f4{sumW, sumW, sumW, sumW * sumW};
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
    const int bx = x + offx, by = y + offy;
    const int bi = ci + offy * BW + offx; // the jittered centre inside the tile: every tap is a compile-time offset from it
    const int bq = cq + offy * BWQ + offx;
    const float fbx = (float)bx, fby = (float)by, fstep = (float)stepSize;
    constexpr int tx[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, ty[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
#pragma unroll
    for (int half = 0; half < 2; ++half)
    {
        float4 sg[4], sv[4];
        uint32_t sm[4];
        float fx[4], fy[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const int t = half * 4 + k;
            const int ti = bi + (ty[t] * BW + tx[t]) * stepSize;
            if (kInterior) { ok[k] = true; fx[k] = fbx + (float)tx[t] * fstep; fy[k] = fby + (float)ty[t] * fstep; }
            else
            {
                int sx = bx + tx[t] * stepSize, sy = by + ty[t] * stepSize;
                ok[k] = sx >= 0 && sy >= 0 && sx < W && sy < H;
                sx = clampi(sx, 0, W - 1); sy = clampi(sy, 0, H - 1);
                fx[k] = (float)sx; fy[k] = (float)sy;
            }
            sg[k] = sG[ti];
            sm[k] = sMQ[bq + (ty[t] * BWQ + tx[t]) * stepSize];
            sv[k] = sIn[ti];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const int t = half * 4 + k;
            constexpr float k3[2] = {0.44198f, 0.27901f};
            atrousTap(pt, cn, cMat, nParam, cLum, phiInv, ok[k], k3[tx[t] & 1] * k3[ty[t] & 1], sg[k], sm[k], sv[k], fx[k], fy[k], sumW, sum);
        }
    }
    const f4 res = sum / f4{sumW, sumW, sumW, sumW * sumW};
    if (kComposite) a.out[pix] = make_float4(res.x * al.x, res.y * al.y, res.z * al.z, 0.0f);
    else a.out[pix] = toFloat4(res);
}

// The tile arrives in TY/8 pieces, each with its own mbarrier: the HEAD box (8 + 2h rows: everything the threads' first pixel row
// group needs) and TY/8 - 1 BODY boxes of 8 more rows. A thread's k-th pixel (row ly0 + 8k) only waits for piece k, so the
// arithmetic of the first rows overlaps the transfer of the rest (ncu r2e: 25 % of the stall samples of the one-barrier version
// sat on the wait). Piece k is issued by lane 0 of warp k.
struct TileMaps { CUtensorMap inHead, gHead, mqHead, inBody, gBody, mqBody; };

template <int STEP, int TY, bool kComposite>
__global__ void __launch_bounds__(256, VPT_ATILE_MINB) atrousTileKernel(const __grid_constant__ AtrousArgs a, const __grid_constant__ TileMaps maps)
{
    constexpr int kHalo = AtrousTile<STEP>::kHalo, kHaloQ = AtrousTile<STEP>::kHaloQ, kPieces = TY / 8, kHeadRows = 8 + 2 * kHalo;
    constexpr int BW = kBX + 2 * kHalo, BH = TY + 2 * kHalo, NPX = BW * BH, BWQ = kBX + 2 * kHaloQ, NPXQ = BWQ * BH;
    static_assert((NPX * 16) % 128 == 0 && (BW * 8 * 16) % 128 == 0 && (BW * kHeadRows * 16) % 128 == 0, "pieces of the float4 tiles must stay 128-byte aligned");
    static_assert((BWQ * 8 * 4) % 128 == 0 && (BWQ * kHeadRows * 4) % 128 == 0, "pieces of the material tile must stay 128-byte aligned");
    extern __shared__ __align__(128) unsigned char tileSmem[];
    float4 *sIn = reinterpret_cast<float4 *>(tileSmem);
    float4 *sG = sIn + NPX;
    uint32_t *sMQ = reinterpret_cast<uint32_t *>(sG + NPX);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tileSmem + (size_t)NPX * 32 + (size_t)NPXQ * 4);
    const int x0 = blockIdx.x * kBX, y0 = a.rowBegin + blockIdx.y * TY;
    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
    if (lx == 0 && ly0 < kPieces)
    {
        const int k = ly0;
        const int row0 = k == 0 ? 0 : kHeadRows + 8 * (k - 1), rows = k == 0 ? kHeadRows : 8; // tile rows of this piece
        tma::barrierInit(bar + k, 1);
        tma::barrierExpectTx(bar + k, (unsigned)(rows * (BW * 32 + BWQ * 4)));
        tma::load2D(sG + row0 * BW, k == 0 ? &maps.gHead : &maps.gBody, 4 * (x0 - kHalo), y0 - kHalo + row0, bar + k);
        tma::load2D(sMQ + row0 * BWQ, k == 0 ? &maps.mqHead : &maps.mqBody, x0 - kHaloQ, y0 - kHalo + row0, bar + k);
        tma::load2D(sIn + row0 * BW, k == 0 ? &maps.inHead : &maps.inBody, 4 * (x0 - kHalo), y0 - kHalo + row0, bar + k);
    }
    const int x = x0 + lx;
    const bool interior = x0 - kHalo >= 0 && x0 + kBX - 1 + kHalo < a.W && y0 - kHalo >= 0 && y0 + TY - 1 + kHalo < a.H; // CTA-uniform
    // the per-pixel scalars that do not come through the tile are requested before the first wait
    float hl[kPieces];
#pragma unroll
    for (int k = 0; k < kPieces; ++k)
    {
        const int y = y0 + ly0 + 8 * k;
        hl[k] = (x < a.W && y < a.rowEnd) ? __ldg(a.histLen + (size_t)y * a.W + x) : 0.0f;
    }
    __syncthreads(); // the barrier words are initialised before anyone polls them
#pragma unroll
    for (int k = 0; k < kPieces; ++k)
    {
        const int ly = ly0 + 8 * k, y = y0 + ly;
        const bool live = x < a.W && y < a.rowEnd;
        float4 al = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (kComposite && live) al = __ldg(a.albedo + (size_t)y * a.W + x);
        if (!tma::barrierWait(bar + k, 0, &g_tmaTimeouts)) return;
        if (!live) continue;
        const int ci = (ly + kHalo) * BW + lx + kHalo, cq = (ly + kHalo) * BWQ + lx + kHaloQ;
        if (interior) atrousTilePixel<STEP, BW, BWQ, kComposite, true>(a, sIn, sG, sMQ, x, y, ci, cq, hl[k], al);
        else atrousTilePixel<STEP, BW, BWQ, kComposite, false>(a, sIn, sG, sMQ, x, y, ci, cq, hl[k], al);
    }
}

// ------------------------------------------------------------------------------------------------ first spatial pass over a TMA tile
#ifndef VPT_AFTILE_MINB
#define VPT_AFTILE_MINB 4
#endif
constexpr int kAfTY = 16, kAfHalo = 2, kAfHaloQ = 4, kAfBW = kBX + 2 * kAfHalo, kAfBH = kAfTY + 2 * kAfHalo, kAfNPX = kAfBW * kAfBH;
constexpr int kAfBWQ = kBX + 2 * kAfHaloQ, kAfNPXQ = kAfBWQ * kAfBH; // the 4-byte plane's box starts on a 16-byte boundary (see AtrousTile)

VPT_DEV void atrousFirstTilePixel(const AtrousArgs &a, const float4 *__restrict__ sIn, const float4 *__restrict__ sG, const uint32_t *__restrict__ sMQ, int x, int y, int ci,
                                  int cq, float hl)
{
    constexpr int BW = kAfBW, BWQ = kAfBWQ;
    const int W = a.W, H = a.H;
    const size_t pix = (size_t)y * W + x;
    const float4 g = sG[ci];
    if (g.w > kSkyZs) return;
    const f3 cn = {g.x, g.y, g.z};
    const uint32_t cMat = sMQ[cq] >> 16;
    const float nParam = a.nParamFull;
    if (hl >= 3.0f)
    {
        float4 sv[9];
#pragma unroll
        for (int cx = -1; cx <= 1; cx++)
#pragma unroll
            for (int cy = -1; cy <= 1; cy++)
                sv[(cx + 1) * 3 + (cy + 1)] = sIn[ci + cy * BW + cx];
        f4 vsum = F4(0.0f);
        const float kern[4] = {1.0f / 4.0f, 1.0f / 8.0f, 1.0f / 8.0f, 1.0f / 16.0f};
#pragma unroll
        for (int dx = -1; dx <= 1; dx++)
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
                vsum += F4(sv[(dx + 1) * 3 + (dy + 1)]) * kern[abs(dx) * 2 + abs(dy)];
        const float v1 = luminance(xyz(vsum));
        const float cVar = fmaxr(0.0f, subSq(vsum.w, v1));
        const float cLum = luminance(xyz(sv[4]));
        const float phiInv = 1.0f / fmaxr(1.0e-4f, a.phiLuminance * sqrtf(cVar));
        const PlaneTest pt = planeTest(a.view, x, y, cn, g.w, a.depthThreshold);
        float sumW = 0.0f; f4 sum = F4(0.0f);
        const float k3[2] = {0.44198f, 0.27901f};
#pragma unroll
        for (int cx = -1; cx <= 1; cx++)
#pragma unroll
            for (int cy = -1; cy <= 1; cy++)
            {
                const int sx = x + cx, sy = y + cy;
                const bool center = (cx == 0 && cy == 0);
                const bool inside = sx >= 0 && sy >= 0 && sx < W && sy < H;
                const float kernel = inside ? k3[abs(cx)] * k3[abs(cy)] : 0.0f;
                const int qx = clampi(sx, 0, W - 1), qy = clampi(sy, 0, H - 1);
                const int ti = ci + cy * BW + cx; // border tiles were patched to clamp-to-edge: entry (sx, sy) holds pixel (qx, qy)
                const float4 sg = sG[ti];
                const uint32_t sMat = sMQ[cq + cy * BWQ + cx] >> 16;
                const float geomW = planeNear(pt, sg.w, (float)qx, (float)qy) ? kernel : 0.0f;
                const float normalW = normalWeight(dot(cn, F3(sg.x, sg.y, sg.z)), nParam);
                const f4 v = F4(sv[(cx + 1) * 3 + (cy + 1)]);
                const float lumW = fabsf(cLum - luminance(xyz(v))) * phiInv;
                float w = geomW * normalW * __expf(-lumW);
                w = center ? kernel : w;
                w = (sMat == cMat) ? w : 0.0f;
                sumW += w;
                sum += w * v;
            }
        sumW = fmaxr(sumW, 1e-6f);
        sum = sum / sumW;
        const float m1 = luminance(xyz(sum));
        a.out[pix] = make_float4(sum.x, sum.y, sum.z, fmaxr(0.0f, subSq(sum.w, m1)));
    }
    else
    {
        float sumW = 0.0f; f3 sumI = F3(0.0f); float s1 = 0.0f, s2 = 0.0f;
        for (int cx = -2; cx <= 2; cx++)
            for (int cy = -2; cy <= 2; cy++)
            {
                const int ti = ci + cy * BW + cx;
                const float4 sg = sG[ti];
                const uint32_t sMat = sMQ[cq + cy * BWQ + cx] >> 16;
                const float normalW = normalWeight(dot(cn, F3(sg.x, sg.y, sg.z)), nParam);
                const f4 v = F4(sIn[ti]);
                const float l1 = luminance(xyz(v));
                const float w = (sMat == cMat) ? normalW : 0.0f;
                sumW += w; sumI += xyz(v) * w; s1 += l1 * w; s2 += v.w * w;
            }
        const float boost = fmaxr(1.0f, 4.0f / (hl + 1.0f));
        sumW = fmaxr(sumW, 1e-6f);
        sumI /= sumW; s1 /= sumW; s2 /= sumW;
        float var = fmaxr(0.0f, subSq(s2, s1));
        var *= boost;
        a.out[pix] = make_float4(sumI.x, sumI.y, sumI.z, var);
    }
}

__global__ void __launch_bounds__(256, VPT_AFTILE_MINB) atrousFirstTileKernel(const __grid_constant__ AtrousArgs a, const __grid_constant__ TileMaps maps)
{
    constexpr int BW = kAfBW, BH = kAfBH, NPX = kAfNPX, kHalo = kAfHalo, kHaloQ = kAfHaloQ, BWQ = kAfBWQ, NPXQ = kAfNPXQ;
    constexpr int kPieces = kAfTY / 8, kHeadRows = 8 + 2 * kHalo;
    static_assert((NPX * 16) % 128 == 0 && (BW * kHeadRows * 16) % 128 == 0 && (BWQ * kHeadRows * 4) % 128 == 0, "pieces must stay 128-byte aligned");
    static_assert(BH == kHeadRows + 8 * (kPieces - 1), "head + body pieces cover the tile");
    extern __shared__ __align__(128) unsigned char tileSmem[];
    float4 *sIn = reinterpret_cast<float4 *>(tileSmem);
    float4 *sG = sIn + NPX;
    uint32_t *sMQ = reinterpret_cast<uint32_t *>(sG + NPX);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tileSmem + (size_t)NPX * 32 + (size_t)NPXQ * 4);
    const int x0 = blockIdx.x * kBX, y0 = a.rowBegin + blockIdx.y * kAfTY;
    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
    if (lx == 0 && ly0 < kPieces)
    {
        const int k = ly0;
        const int row0 = k == 0 ? 0 : kHeadRows + 8 * (k - 1), rows = k == 0 ? kHeadRows : 8;
        tma::barrierInit(bar + k, 1);
        tma::barrierExpectTx(bar + k, (unsigned)(rows * (BW * 32 + BWQ * 4)));
        tma::load2D(sG + row0 * BW, k == 0 ? &maps.gHead : &maps.gBody, 4 * (x0 - kHalo), y0 - kHalo + row0, bar + k);
        tma::load2D(sMQ + row0 * BWQ, k == 0 ? &maps.mqHead : &maps.mqBody, x0 - kHaloQ, y0 - kHalo + row0, bar + k);
        tma::load2D(sIn + row0 * BW, k == 0 ? &maps.inHead : &maps.inBody, 4 * (x0 - kHalo), y0 - kHalo + row0, bar + k);
    }
    const int x = x0 + lx;
    const bool interior = x0 - kHalo >= 0 && x0 + kBX - 1 + kHalo < a.W && y0 - kHalo >= 0 && y0 + kAfTY - 1 + kHalo < a.H; // CTA-uniform
    float hl[kPieces];
#pragma unroll
    for (int k = 0; k < kPieces; ++k)
    {
        const int y = y0 + ly0 + 8 * k;
        hl[k] = (x < a.W && y < a.rowEnd) ? __ldg(a.histLen + (size_t)y * a.W + x) : 0.0f;
    }
    __syncthreads();
    if (!interior)
    {
        // clamp-to-edge (cudaBoundaryModeClamp): an out-of-image entry takes the value of the nearest in-image pixel, which is an
        // entry of this same tile; sources are in-image entries, which nobody writes. Needs the whole tile.
        for (int k = 0; k < kPieces; ++k)
            if (!tma::barrierWait(bar + k, 0, &g_tmaTimeouts)) return;
        for (int i = threadIdx.x; i < NPX; i += 256)
        {
            const int ey = i / BW, ex = i - ey * BW;
            const int gx = x0 - kHalo + ex, gy = y0 - kHalo + ey;
            const int qx = clampi(gx, 0, a.W - 1), qy = clampi(gy, 0, a.H - 1);
            if (qx != gx || qy != gy)
            {
                const int sy = qy - (y0 - kHalo), sx = qx - (x0 - kHalo);
                sIn[i] = sIn[sy * BW + sx]; sG[i] = sG[sy * BW + sx];
                sMQ[ey * BWQ + ex + (kHaloQ - kHalo)] = sMQ[sy * BWQ + sx + (kHaloQ - kHalo)];
            }
        }
        __syncthreads();
    }
#pragma unroll 1
    for (int k = 0; k < kPieces; ++k)
    {
        const int ly = ly0 + 8 * k, y = y0 + ly;
        if (interior && !tma::barrierWait(bar + k, 0, &g_tmaTimeouts)) return;
        if (x >= a.W || y >= a.rowEnd) continue;
        atrousFirstTilePixel(a, sIn, sG, sMQ, x, y, (ly + kHalo) * BW + lx + kHalo, (ly + kHalo) * BWQ + lx + kHaloQ, hl[k]);
    }
}

// ------------------------------------------------------------------------------------------------ history clamping, column-walking warps
// A warp owns 32 consecutive columns — 28 output columns and a 2-column apron either side — and walks down `rowsPerWarp` output
// rows (plus 2 apron rows above and below). Per pixel the transforms (responsive YCoCg, squares, noisy rgb, noisy luminance^2) are
// evaluated ONCE and pushed into a 5-row register ring; the 5x5 moments are the vertical sum of the ring followed by the
// horizontal sum over the neighbouring lanes (shuffles). Loads are one full 512-byte line per plane per row, two rows ahead.
#ifndef VPT_HCCOL_ROWS
#define VPT_HCCOL_ROWS 30
#endif
#ifndef VPT_HCCOL_MINB
#define VPT_HCCOL_MINB 3
#endif
constexpr int kHcWarps = 4, kHcOutCols = 28, kHcRows = VPT_HCCOL_ROWS;

struct Moments { float c[10]; };
VPT_DEV Moments momentsOf(float4 resp, float4 noisy)
{
    const f3 s = rgbToYCoCg(xyz(resp));
    const f3 nz = xyz(noisy);
    const float nl = luminance(nz);
    Moments m;
    m.c[0] = s.x; m.c[1] = s.y; m.c[2] = s.z; m.c[3] = s.x * s.x; m.c[4] = s.y * s.y; m.c[5] = s.z * s.z;
    m.c[6] = nz.x; m.c[7] = nz.y; m.c[8] = nz.z; m.c[9] = nl * nl;
    return m;
}

__global__ void __launch_bounds__(kHcWarps * 32, VPT_HCCOL_MINB) historyClampColKernel(const __grid_constant__ ClampArgs a)
{
    const int W = a.W, H = a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int colBase = (blockIdx.x * kHcWarps + warp) * kHcOutCols; // first output column of this warp
    if (colBase >= W) return;
    const int x = colBase - 2 + lane;                                 // this lane's column (apron lanes: 0, 1, 30, 31)
    const int xc = clampi(x, 0, W - 1);
    const int rBegin = a.rowBegin + blockIdx.y * kHcRows;
    const int rEnd = min(rBegin + kHcRows, a.rowEnd);
    const bool outLane = lane >= 2 && lane < 30 && x < W;
    Moments ring[5];
    // rows rBegin-2 .. rEnd+1 are consumed; a row's planes are requested two iterations before they are used (pf0 = next row,
    // pf1 = the row after it)
    float4 pf0Resp, pf0Noisy, pf1Resp, pf1Noisy;
    {
        const size_t p0 = (size_t)clampi(rBegin - 2, 0, H - 1) * W + xc, p1 = (size_t)clampi(rBegin - 1, 0, H - 1) * W + xc;
        pf0Resp = __ldg(a.pong + p0); pf0Noisy = __ldg(a.illum + p0);
        pf1Resp = __ldg(a.pong + p1); pf1Noisy = __ldg(a.illum + p1);
    }
    // the per-pixel planes of an OUTPUT row (depth, history length, accumulated history) are requested one row ahead as well
    float nDepth, nHl; float4 nPing;
    {
        const size_t p = (size_t)rBegin * W + xc;
        nDepth = __ldg(a.depth + p); nHl = __ldg(a.histLen + p); nPing = __ldg(a.ping + p);
    }
    // the ring slot of a row is (row - (rBegin - 2)) % 5; the loop is unrolled by 5 so every slot index is a compile-time constant
    for (int r0 = rBegin - 2; r0 < rEnd + 2; r0 += 5)
    {
#pragma unroll
        for (int j = 0; j < 5; ++j)
        {
            const int r = r0 + j;
            if (r >= rEnd + 2) break;
            const float4 resp = pf0Resp, noisy = pf0Noisy;
            pf0Resp = pf1Resp; pf0Noisy = pf1Noisy;
            {
                const size_t p = (size_t)clampi(r + 2, 0, H - 1) * W + xc;
                pf1Resp = __ldg(a.pong + p); pf1Noisy = __ldg(a.illum + p);
            }
            ring[j] = momentsOf(resp, noisy);
            const int y = r - 2; // the output row whose 5 rows are now in the ring (slots j-4 .. j, modulo 5)
            if (y < rBegin) continue;
            const float cDepth = nDepth, hl = nHl; const float4 cPing = nPing;
            {
                const size_t p = (size_t)min(y + 1, rEnd - 1) * W + xc;
                nDepth = __ldg(a.depth + p); nHl = __ldg(a.histLen + p); nPing = __ldg(a.ping + p);
            }
            // vertical sum, top row first (the slot after j holds the oldest row), then horizontal over lanes x-2 .. x+2
            float m[10];
#pragma unroll
            for (int c = 0; c < 10; ++c)
            {
                float v = ring[(j + 1) % 5].c[c];
                v += ring[(j + 2) % 5].c[c]; v += ring[(j + 3) % 5].c[c]; v += ring[(j + 4) % 5].c[c]; v += ring[j].c[c];
                // v(x-2) + v(x-1) + v(x) + v(x+1) + v(x+2) as pair(x-2) + pair(x) + v(x+2), pair(x) = v(x) + v(x+1): 3 shuffles
                const float pair = v + __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
                float h = __shfl_sync(0xffffffffu, pair, (lane + 30) & 31);
                h += pair;
                h += __shfl_sync(0xffffffffu, v, (lane + 2) & 31);
                m[c] = h;
            }
            if (!outLane) continue;
            const size_t pix = (size_t)y * W + x;
            if (cDepth > kDenoisingRange) continue;
            const Moments &ctr = ring[(j + 3) % 5]; // row y
            f3 rM1 = {m[0], m[1], m[2]}, rM2 = {m[3], m[4], m[5]}, nM1 = {m[6], m[7], m[8]};
            float nM2 = m[9];
            rM1 /= 25.0f; rM2 /= 25.0f; nM1 /= 25.0f; nM2 /= 25.0f;
            const f3 sigma = sqrt3(max3f(F3(0.0f), F3(subSq(rM2.x, rM1.x), subSq(rM2.y, rM1.y), subSq(rM2.z, rM1.z))));
            f3 cmin = rM1 - 2.0f * sigma, cmax = rM1 + 2.0f * sigma;
            const f3 centerY = {ctr.c[0], ctr.c[1], ctr.c[2]};
            cmin = (cmin.x < centerY.x) ? cmin : centerY;
            cmax = (cmax.x > centerY.x) ? cmax : centerY;
            const f4 acc = F4(cPing);
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
            const float tSigma = 0.5f * sqrtf(fmaxr(0.0f, subSq(nM2, noisyL)));
            const float sSigma = 4.5f * sigma.x;
            float reset = 0.5f * fmaxr(0.0f, fabsf(diffL - noisyL) - sSigma - tSigma) / (1.0e-6f + fmaxr(diffL, noisyL) + sSigma + tSigma);
            reset = saturate(reset);
            const f3 noisyC = {ctr.c[6], ctr.c[7], ctr.c[8]};
            f3 d3 = lerp3(xyz(outD), noisyC, reset), r3 = lerp3(xyz(outR), noisyC, reset);
            outD = F4(d3, outD.w); outR = F4(r3, outR.w);
            const float outL = luminance(xyz(outD));
            outD.w += diffSq(outL, diffL);
            outD.w = fmaxr(0.0f, outD.w);
            a.prevIllum[pix] = toFloat4(outD);
            a.prevFast[pix] = toFloat4(outR);
            a.prevHistLen[pix] = hl;
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
// VPT_DN_TILE_MASK (debug / A-B): bit 0 first a-trous pass, bit 1 a-trous passes, bit 2 column-walking history clamp
static unsigned tileMask()
{
    // default 3: the column-walking clamp is opt-in — measured on B200 it ties the shared-memory tile kernel at 4K (242 vs 248 us)
    // and loses at 1080p (81 vs 70 us: 139 registers, 12 warps per SM)
    static const unsigned m = [] { const char *e = std::getenv("VPT_DN_TILE_MASK"); return e ? (unsigned)std::strtoul(e, nullptr, 0) : 3u; }();
    return m;
}
bool tileClampEnabled() { return (tileMask() & 4u) != 0; }
static bool tileShapeOk(const DenoiseLaunch &d)
{
    // TMA: row pitches are multiples of 16 bytes (the 4-byte material plane needs W % 4 == 0); 16-byte aligned bases (cudaMalloc)
    return (d.width & 3) == 0 && d.width >= kBX;
}

template <int STEP, int TY>
static cudaError_t launchAtrousTileT(const DenoiseLaunch &d, const AtrousArgs &a, bool composite)
{
    constexpr int kHalo = AtrousTile<STEP>::kHalo, BW = kBX + 2 * kHalo, BH = TY + 2 * kHalo, BWQ = kBX + 2 * AtrousTile<STEP>::kHaloQ, kHeadRows = 8 + 2 * kHalo;
    constexpr size_t smem = (size_t)BW * BH * 32 + (size_t)BWQ * BH * 4 + 8 * (TY / 8);
    TileMaps m;
    cudaError_t e;
    if ((e = planeMap(&m.inHead, a.in, d.width, d.height, BW, kHeadRows, 4, false)) != cudaSuccess) return e;
    if ((e = planeMap(&m.gHead, a.G, d.width, d.height, BW, kHeadRows, 4, false)) != cudaSuccess) return e;
    if ((e = planeMap(&m.mqHead, a.MQ, d.width, d.height, BWQ, kHeadRows, 1, true)) != cudaSuccess) return e;
    if ((e = planeMap(&m.inBody, a.in, d.width, d.height, BW, 8, 4, false)) != cudaSuccess) return e;
    if ((e = planeMap(&m.gBody, a.G, d.width, d.height, BW, 8, 4, false)) != cudaSuccess) return e;
    if ((e = planeMap(&m.mqBody, a.MQ, d.width, d.height, BWQ, 8, 1, true)) != cudaSuccess) return e;
    const dim3 grid((d.width + kBX - 1) / kBX, (d.rowEnd - d.rowBegin + TY - 1) / TY);
    if (composite)
    {
        if ((e = cudaFuncSetAttribute(atrousTileKernel<STEP, TY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        atrousTileKernel<STEP, TY, true><<<grid, 256, smem, d.stream>>>(a, m);
    }
    else
    {
        if ((e = cudaFuncSetAttribute(atrousTileKernel<STEP, TY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        atrousTileKernel<STEP, TY, false><<<grid, 256, smem, d.stream>>>(a, m);
    }
    return cudaGetLastError();
}

#ifndef VPT_ATILE_TY2
#define VPT_ATILE_TY2 16
#endif
#ifndef VPT_ATILE_TY4
#define VPT_ATILE_TY4 16
#endif
#ifndef VPT_ATILE_TY8
#define VPT_ATILE_TY8 16 // 32 rows (97 KB, 2 CTAs per SM) measured slower: 142 vs 128 us for the three passes at 1080p; 8-row tiles slower again (144)
#endif
cudaError_t launchAtrousTiled(const DenoiseLaunch &d, const float4 *in, float4 *out, unsigned frameIndex, unsigned step, bool composite, bool *handled)
{
    *handled = false;
    if (!(tileMask() & 2u) || !tileShapeOk(d) || !(step == 2u || step == 4u || step == 8u)) return cudaSuccess;
    const AtrousArgs a = makeAtrousArgs(d, in, out, frameIndex, step);
    *handled = true;
    if (step == 2u) return launchAtrousTileT<2, VPT_ATILE_TY2>(d, a, composite);
    if (step == 4u) return launchAtrousTileT<4, VPT_ATILE_TY4>(d, a, composite);
    return launchAtrousTileT<8, VPT_ATILE_TY8>(d, a, composite);
}

cudaError_t launchAtrousSmemTiled(const DenoiseLaunch &d, const float4 *in, float4 *out, bool *handled)
{
    *handled = false;
    if (!(tileMask() & 1u) || !tileShapeOk(d)) return cudaSuccess;
    const AtrousArgs a = makeAtrousArgs(d, in, out, 0, 1);
    constexpr int kHeadRows = 8 + 2 * kAfHalo;
    TileMaps m;
    cudaError_t e;
    if ((e = planeMap(&m.inHead, a.in, d.width, d.height, kAfBW, kHeadRows, 4, false)) != cudaSuccess) return e;
    if ((e = planeMap(&m.gHead, a.G, d.width, d.height, kAfBW, kHeadRows, 4, false)) != cudaSuccess) return e;
    if ((e = planeMap(&m.mqHead, a.MQ, d.width, d.height, kAfBWQ, kHeadRows, 1, true)) != cudaSuccess) return e;
    if ((e = planeMap(&m.inBody, a.in, d.width, d.height, kAfBW, 8, 4, false)) != cudaSuccess) return e;
    if ((e = planeMap(&m.gBody, a.G, d.width, d.height, kAfBW, 8, 4, false)) != cudaSuccess) return e;
    if ((e = planeMap(&m.mqBody, a.MQ, d.width, d.height, kAfBWQ, 8, 1, true)) != cudaSuccess) return e;
    constexpr size_t smem = (size_t)kAfNPX * 32 + (size_t)kAfNPXQ * 4 + 8 * (kAfTY / 8);
    const dim3 grid((d.width + kBX - 1) / kBX, (d.rowEnd - d.rowBegin + kAfTY - 1) / kAfTY);
    atrousFirstTileKernel<<<grid, 256, smem, d.stream>>>(a, m);
    *handled = true;
    return cudaGetLastError();
}

cudaError_t launchHistoryClampingCols(const DenoiseLaunch &d)
{
    ClampArgs a;
    a.W = d.width; a.H = d.height; a.rowBegin = d.rowBegin; a.rowEnd = d.rowEnd;
    a.depth = d.b.cur.depth; a.illum = d.b.illumination; a.ping = d.b.ping; a.pong = d.b.pong; a.histLen = d.b.historyLength;
    a.prevIllum = d.b.prevIllum; a.prevFast = d.b.prevFastIllum; a.prevHistLen = d.b.prevHistoryLength;
    const int warpsX = (d.width + kHcOutCols - 1) / kHcOutCols;
    const dim3 grid((warpsX + kHcWarps - 1) / kHcWarps, (d.rowEnd - d.rowBegin + kHcRows - 1) / kHcRows);
    historyClampColKernel<<<grid, kHcWarps * 32, 0, d.stream>>>(a);
    return cudaGetLastError();
}

} // namespace vpt
