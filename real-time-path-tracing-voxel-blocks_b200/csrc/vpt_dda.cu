// Ray–voxel DDA engine for sm_100a: replaces optixTraverse over the triangle BVH
// (/root/reference/renderer/shaders/RayGen.cu:49-54, closesthit.cu:458,616,745,801) — B200 has no RT cores.
//
// Algorithm: Amanatides–Woo with the comparison order and the tMax += tDelta accumulation of
// VoxelEngine::performRayTraversal (/root/reference/voxelengine/VoxelEngine.cu:1133-1162), bit-exact vs the oracle.
//
// B200 shape:
//  * one persistent 1024-thread CTA per SM; the padded 1-bit occupancy mask (86 KiB for the 16-chunk world) is staged
//    whole in shared memory, so a step is one LDS + bit test; worlds whose mask exceeds shared memory walk it
//    through L1/L2 (kSmem = false).
//  * the mask has a solid one-voxel shell: leaving the grid is "hitting" the shell — the step loop has NO bounds
//    arithmetic. 18 SASS instructions per step (words are stored bit-reversed: shift + sign test). Rays with dir.y > 0 walk a second copy of the mask that is solid
//    from the highest solid voxel up (GridView::upH): they retire as soon as nothing can be above them.
//  * warp-level ray compaction: a warp reserves chunks of the prepared-ray queue (one atomic per 64 rays); whenever
//    at most kRefillBelow lanes still hold a live ray, the idle lanes are re-armed from the queue
//    (__ballot_sync/__popc slot assignment), so the step loop runs with most lanes active regardless of how
//    different the trip counts are. Lanes without a ray are parked on a spare all-zero mask word with zero
//    strides: they execute the same instructions harmlessly — no per-step predicate.
#include "vpt_dda.cuh"

namespace vpt {

#ifndef VPT_DDA_THREADS
#define VPT_DDA_THREADS 1024
#endif
constexpr int kDdaThreads = VPT_DDA_THREADS;
#ifndef VPT_DDA_CHUNK
#define VPT_DDA_CHUNK 64 // rays reserved per warp per atomic; measured (DDA ms/frame): 32 -> 0.963, 64 -> 0.923, 128 -> 0.934, 256 -> 0.997
#endif
#ifndef VPT_DDA_REFILL
#define VPT_DDA_REFILL 20
#endif
#ifndef VPT_DDA_UNROLL
#define VPT_DDA_UNROLL 32 // closest-hit launches (coherent primary rays): 16 -> 0.947, 32 -> 0.933, 48 -> 0.939, 64 -> 0.933 ms of DDA per frame. Both kinds, earlier sweep: measured on B200 (DDA ms/frame): 3 -> 1.196, 4 -> 1.121, 6 -> 1.037, 8 -> 0.993, 12 -> 0.961, 16 -> 0.945, 24 -> 0.946, 32 -> 0.963
#endif
#ifndef VPT_DDA_MULHI
#define VPT_DDA_MULHI 0 // lin >> 5 as IMAD.HI (fma pipe) instead of SHF + LOP3: measured slower (0.983 vs 0.943 ms)
#endif
#ifndef VPT_DDA_UNROLL_ANY
#define VPT_DDA_UNROLL_ANY 16 // any-hit launches (ray lengths vary more): 8 -> 0.981, 12 -> 0.953, 16 -> 0.947, 24 -> 0.951
#endif
#ifndef VPT_DDA_BREAK
#define VPT_DDA_BREAK 1
#endif
constexpr int kChunk = VPT_DDA_CHUNK;        // rays reserved per warp per atomic
constexpr int kRefillBelow = VPT_DDA_REFILL; // re-arm idle lanes when <= this many lanes are live
constexpr unsigned kFull = 0xffffffffu;

template <bool kSmem, bool kClosest, bool kStats, bool kTmax>
__global__ void __launch_bounds__(kDdaThreads, 1) ddaKernel(const __grid_constant__ DdaArgs a)
{
    extern __shared__ uint32_t occS[];
    if (kSmem)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.grid.occ);
        uint4 *dst = reinterpret_cast<uint4 *>(occS);
        const int n4 = a.grid.occWords >> 2; // occWords is a multiple of 4
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    const uint32_t *__restrict__ occG = a.grid.occ;
    const unsigned count = __ldg(a.count);
    const unsigned lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const int Wp = a.grid.Wp, Dp = a.grid.Dp, W = a.grid.W, H = a.grid.H, D = a.grid.D;
    const int strideY = Wp * Dp;
    const int parkLin = a.grid.parkLin, maskBits = a.grid.maskWords * 32, upH = a.grid.upH;
    // padded bit index -> voxel; shell = outside the grid, or at/above upH in the upward mask
    auto decode = [&](int l, int &x, int &y, int &z) -> bool {
        const bool up = l >= maskBits;
        if (up) l -= maskBits;
        uint32_t r, xp, yp, zp;
        a.grid.divWp.div((uint32_t)l, r, xp);
        a.grid.divDp.div(r, yp, zp);
        x = (int)xp - 1; y = (int)yp - 1; z = (int)zp - 1;
        return (unsigned)x >= (unsigned)W || (unsigned)z >= (unsigned)D || (unsigned)y >= (unsigned)(up ? upH : H);
    };

    // per-lane DDA state
    float tX = 0.0f, tY = 0.0f, tZ = 0.0f, dtX = 0.0f, dtY = 0.0f, dtZ = 0.0f, tCur = 0.0f, tmin = 0.0f, tmax = kRayMax;
    int lin = parkLin, dX = 0, dY = 0, dZ = 0, lastD = 0;
    uint32_t meta = 0, result = 0;
    bool live = false;
    // a finished ray's result is written later, together with the other idle lanes (convergent), not inside the step loop
    bool pending = false;
    int finLin = 0, finD = 0;
    float finT = 0.0f;
    // warp-uniform queue chunk
    unsigned chunkPos = 0, chunkEnd = 0;
    bool exhausted = (count == 0);
    unsigned raysAcc = 0, stepsAcc = 0;

    for (;;)
    {
        // ---- retire finished rays
        if (pending)
        {
            int x, y, z;
            // a solid voxel first met at or beyond the ray's far end is no hit (the oracle's walk stops at tCur >= tmax; every voxel
            // before that one was empty, so stopping there or walking on to the first solid voxel decides the same)
            const bool shell = decode(finLin, x, y, z) || (kTmax && finT >= tmax); // (the lane's tmax is only replaced by the re-arm below;
            // kTmax: only the visibility launches of scenes with local lights carry a finite far end)
            if (kClosest)
            {
                uint32_t packed = kHitMiss;
                if (!shell)
                {
                    uint32_t face = (meta >> 4) & 7u;
                    if (finD != 0)
                    {
                        const int ax = finD < 0 ? -finD : finD;
                        if (ax == 1) face = (meta & 1u) ? 2u : 3u;
                        else if (ax == strideY) face = (meta & 2u) ? 1u : 0u;
                        else face = (meta & 4u) ? 5u : 4u;
                    }
                    packed = ((uint32_t)((y * D + z) * W + x) << 3) | face;
                }
                a.hitT[result] = shell ? kRayMax : finT;
                a.hitPacked[result] = packed;
            }
            else
                a.vis[result] = shell ? (uint8_t)0 : (uint8_t)1;
            pending = false;
        }
        // ---- re-arm idle lanes
        unsigned idle = __ballot_sync(kFull, !live);
        while (idle != 0 && !exhausted)
        {
            if (chunkPos >= chunkEnd)
            {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(a.cursor, (unsigned)kChunk);
                base = __shfl_sync(kFull, base, 0);
                if (base >= count) { exhausted = true; break; }
                chunkPos = base;
                chunkEnd = min(base + (unsigned)kChunk, count);
            }
            const unsigned avail = chunkEnd - chunkPos;
            const unsigned rank = __popc(idle & ltMask);
            const bool take = !live && rank < avail;
            if (take)
            {
                const uint4 *q = a.queue + (size_t)(chunkPos + rank) * 3;
                const uint4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
                tX = __uint_as_float(q0.x); tY = __uint_as_float(q0.y); tZ = __uint_as_float(q0.z); tCur = __uint_as_float(q0.w);
                dtX = __uint_as_float(q1.x); dtY = __uint_as_float(q1.y); dtZ = __uint_as_float(q1.z); tmin = __uint_as_float(q1.w);
                lin = (int)q2.x; meta = q2.y; result = q2.z; if (kTmax) tmax = __uint_as_float(q2.w);
                dX = (meta & 1u) ? 1 : -1;
                dY = (meta & 2u) ? strideY : -strideY;
                dZ = (meta & 4u) ? Wp : -Wp;
                lastD = 0;
                live = true;
                ++raysAcc;
            }
            chunkPos += min((unsigned)__popc(idle), avail);
            idle = __ballot_sync(kFull, !live);
        }
        if (idle == kFull) break; // nothing live and nothing left to take

        // ---- step loop
        for (;;)
        {
#pragma unroll
            for (int u = 0; u < (kClosest ? VPT_DDA_UNROLL : VPT_DDA_UNROLL_ANY); ++u)
            {
                uint32_t word;
#if VPT_DDA_MULHI
                if (kSmem) word = occS[__umulhi((unsigned)lin, 0x08000000u)]; // lin >> 5 on the fma pipe (the alu pipe is the bottleneck)
#else
                if (kSmem) word = occS[(unsigned)lin >> 5];
#endif
                else word = __ldg(occG + ((unsigned)lin >> 5));
#define VPT_DDA_ADVANCE()                                                                          \
    do {                                                                                           \
        const bool xy = tX < tY;                                                                   \
        const float tA = xy ? tX : tY;                                                             \
        const bool az = tA < tZ;                                                                   \
        tCur = az ? tA : tZ;                                                                       \
        const int dd = az ? (xy ? dX : dY) : dZ;                                                   \
        /* tCur IS the advanced axis' tMax: three conditional adds, no dt select */                \
        if (az && xy) tX = __fadd_rn(tX, dtX);                                                     \
        if (az && !xy) tY = __fadd_rn(tY, dtY);                                                    \
        if (!az) tZ = __fadd_rn(tZ, dtZ);                                                          \
        lin += dd;                                                                                 \
        lastD = dd;                                                                                \
        if (kStats) stepsAcc += live ? 1u : 0u;                                                    \
    } while (0)
                if ((int)(word << (lin & 31)) < 0) // mask words are bit-reversed: voxel k of a word sits at bit 31-k
                {
                    // solid voxel or shell: once per ray (twice for a ray whose origin voxel lies before tmin)
                    bool fin = tCur >= tmin;
                    if (!fin) { int x, y, z; fin = decode(lin, x, y, z); }
                    if (fin)
                    {
                        finLin = lin; finT = tCur; finD = lastD;
                        pending = true;
                        live = false;
                        lin = parkLin; dX = 0; dY = 0; dZ = 0; // parked: steps in place on an empty spare word
                    }
#if VPT_DDA_BREAK
                    // leave the unrolled block: the divergent region then closes once per block, not once per step, and a
                    // finished lane has nothing to do before the next ballot anyway
                    else VPT_DDA_ADVANCE();
                    break;
#endif
                }
                // advance the axis with the smallest tMax: X<Y ? (X<Z ? X : Z) : (Y<Z ? Y : Z)  ==  A = min(X,Y); A<Z ? A : Z
                VPT_DDA_ADVANCE();
            }
            const unsigned act = __ballot_sync(kFull, live);
            if (act == 0u) break;
            if (!exhausted && __popc(act) <= kRefillBelow) break;
        }
    }
    // statistics: one atomic pair per warp
    unsigned long long r64 = raysAcc, s64 = stepsAcc;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        r64 += __shfl_down_sync(kFull, r64, off);
        s64 += __shfl_down_sync(kFull, s64, off);
    }
    if (lane == 0 && r64) { atomicAdd(a.counters + 0, r64); atomicAdd(a.counters + 2, r64); if (kStats) atomicAdd(a.counters + 1, s64); } // [2]: running total over frames
}

template <bool kSmem, bool kClosest, bool kStats, bool kTmax>
static cudaError_t launchDdaT(const DdaArgs &a, cudaStream_t s, int smCount)
{
    size_t smem = 0;
    if (kSmem)
    {
        smem = (size_t)a.grid.occWords * 4;
        cudaError_t e = cudaFuncSetAttribute(ddaKernel<kSmem, kClosest, kStats, kTmax>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    ddaKernel<kSmem, kClosest, kStats, kTmax><<<smCount, kDdaThreads, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launchDda(const DdaArgs &a, bool closest, bool occInSmem, bool countSteps, bool farEnd, cudaStream_t s, int smCount)
{
    const int sel = (occInSmem ? 4 : 0) | (closest ? 2 : 0) | (countSteps ? 1 : 0);
    if (farEnd && !closest) // finite tmax only exists on visibility rays
        switch (sel)
        {
        case 0: return launchDdaT<false, false, false, true>(a, s, smCount);
        case 1: return launchDdaT<false, false, true, true>(a, s, smCount);
        case 4: return launchDdaT<true, false, false, true>(a, s, smCount);
        default: return launchDdaT<true, false, true, true>(a, s, smCount);
        }
    switch (sel)
    {
    case 0: return launchDdaT<false, false, false, false>(a, s, smCount);
    case 1: return launchDdaT<false, false, true, false>(a, s, smCount);
    case 2: return launchDdaT<false, true, false, false>(a, s, smCount);
    case 3: return launchDdaT<false, true, true, false>(a, s, smCount);
    case 4: return launchDdaT<true, false, false, false>(a, s, smCount);
    case 5: return launchDdaT<true, false, true, false>(a, s, smCount);
    case 6: return launchDdaT<true, true, false, false>(a, s, smCount);
    default: return launchDdaT<true, true, true, false>(a, s, smCount);
    }
}

} // namespace vpt
