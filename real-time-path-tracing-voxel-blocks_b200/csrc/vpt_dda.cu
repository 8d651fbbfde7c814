// Ray–voxel DDA engine for sm_100a: replaces optixTraverse over the triangle BVH
// (/root/reference/renderer/shaders/RayGen.cu:49-54, closesthit.cu:458,616,745,801) — B200 has no RT cores.
//
// Algorithm: Amanatides–Woo with the comparison order and the tMax += tDelta accumulation of
// VoxelEngine::performRayTraversal (/root/reference/voxelengine/VoxelEngine.cu:1133-1162), bit-exact vs the oracle.
//
// B200 shape:
//  * one persistent 1024-thread CTA per SM; the padded 1-bit occupancy masks (2 x 86 KiB for the 16-chunk world) are staged
//    whole in shared memory, so a step is one LDS + bit test; worlds whose masks exceed shared memory walk them
//    through L1/L2 (ddaKernel<kSmem = false>).
//  * the mask has a solid one-voxel shell: leaving the grid is "hitting" the shell — the step loop has NO bounds
//    arithmetic. Rays with dir.y > 0 walk a second copy of the mask that is solid from the highest solid voxel up
//    (GridView::upH): they retire as soon as nothing can be above them.
//  * warp-level ray compaction: a warp reserves chunks of the prepared-ray queue (one atomic per 64 rays); whenever
//    at most kRefillBelow lanes still hold a live ray, the idle lanes are retired and re-armed from the queue together
//    (__ballot_sync/__popc slot assignment), so the step loop runs with most lanes active regardless of how
//    different the trip counts are.
//  * two kernels, same decisions and roundings, bit-identical results (DESIGN.md §3 has the profile that led here):
//      ddaFlatKernel  the any-hit launches: a block of 16 steps without a single branch — "alive" is a predicate threaded
//                     through the block, a finished lane's instructions are predicated off; 16 SASS instructions per step.
//      ddaKernel      the closest-hit launch (coherent primary rays) and global-memory masks: a hit leaves the unrolled block;
//                     lanes without a ray are parked on a spare all-zero mask word with zero strides.
#include "vpt_dda.cuh"

namespace vpt {

#ifndef VPT_DDA_THREADS
#define VPT_DDA_THREADS 1024
#endif
constexpr int kDdaThreads = VPT_DDA_THREADS;
#ifndef VPT_DDA_CHUNK
#define VPT_DDA_CHUNK 64 // rays reserved per warp per atomic; measured (DDA ms/frame): 32 -> 0.844, 64 -> 0.789, 128 -> 0.797 (round-1 engine: 0.963 / 0.923 / 0.934)
#endif
#ifndef VPT_DDA_REFILL
#define VPT_DDA_REFILL 20
#endif
#ifndef VPT_DDA_UNROLL
#define VPT_DDA_UNROLL 32 // closest-hit launches (coherent primary rays), steps per unrolled block: 16 -> 0.798, 32 -> 0.789, 48 -> 0.792 ms of DDA per frame
#endif
#ifndef VPT_DDA_UNROLL_ANY
#define VPT_DDA_UNROLL_ANY 16 // any-hit launches on ddaKernel (global-memory masks; shared-memory masks: ddaFlatKernel, VPT_DDA_FLAT_BLOCK)
#endif
#ifndef VPT_DDA_BREAK
#define VPT_DDA_BREAK 1
#endif
// The step is bound by the ALU pipe, not by issue slots: LOP3 / SHF / SEL / FSEL / FSETP / ISETP / FMNMX all go to the alu pipe
// (one warp instruction per 2 cycles per SM sub-partition), FADD / IMAD to the fma pipe. The select form below compiles to
// 12 alu + 4 fma instructions of 18: 24 alu cycles per step = the 73 % issue-slot utilisation ncu reports (r2p capture).
// VPT_DDA_PRED: the same decision as ONE 3-input minimum (FMNMX3) and equality predicates — Z wins ties, then Y, then X, exactly
// the nested '<' of VoxelEngine.cu:1133-1162 — with predicated FADD / IADD instead of selects: 16 instructions, 9 alu.
// (Zero signs: an axis' tMax is only ever -0, never +0 — (boundary - o)/d is zero only for a negative step — so the minimum
// returns the same bits as the select.)
// VPT_DDA_BYTE: the shared-memory mask is staged byte-swapped and read with LDS.U8 at lin >> 3 (one alu instruction instead of
// SHF + LOP3); the byte is replicated over the word by an IMAD (fma pipe) so the shift-by-lin sign test stays as it is.
#ifndef VPT_DDA_PRED
#define VPT_DDA_PRED 1
#endif
#ifndef VPT_DDA_BYTE
#define VPT_DDA_BYTE 1
#endif
constexpr int kChunk = VPT_DDA_CHUNK;        // rays reserved per warp per atomic
constexpr int kRefillBelow = VPT_DDA_REFILL; // re-arm idle lanes when <= this many lanes are live
constexpr unsigned kFull = 0xffffffffu;

// one mask byte replicated over the word: the shift-by-lin sign test then finds bit 7 - (lin & 7) of it wherever lin & 31 points
__device__ __forceinline__ uint32_t repByte(uint32_t b)
{
    uint32_t w;
    asm("mul.lo.u32 %0, %1, 0x01010101;" : "=r"(w) : "r"(b)); // IMAD (fma pipe), not PRMT / shifts (alu pipe)
    return w;
}

template <bool kSmem, bool kClosest, bool kStats, bool kTmax>
__global__ void __launch_bounds__(kDdaThreads, 1) ddaKernel(const __grid_constant__ DdaArgs a)
{
    extern __shared__ uint32_t occS[];
    if (kSmem)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.grid.occ);
        uint4 *dst = reinterpret_cast<uint4 *>(occS);
        const int n4 = a.grid.occWords >> 2; // occWords is a multiple of 4
        for (int i = threadIdx.x; i < n4; i += blockDim.x)
        {
            uint4 w = __ldg(src + i);
#if VPT_DDA_BYTE
            // byte-swapped: voxel k of a word (bit 31-k) then lives in byte k >> 3 of the word, i.e. at byte address lin >> 3
            w.x = __byte_perm(w.x, 0u, 0x0123u); w.y = __byte_perm(w.y, 0u, 0x0123u); w.z = __byte_perm(w.z, 0u, 0x0123u); w.w = __byte_perm(w.w, 0u, 0x0123u);
#endif
            dst[i] = w;
        }
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
    int lin = parkLin, dX = 0, dY = 0, dZ = 0;
    // kPred: the FMNMX3 / predicated-add step (masks in shared memory). Worlds walked through L1/L2 keep the select form: measured on
    // the cfg5-shaped trace (tools/cfg5_quick.py): select form 48.3 ms, predicated form 49.0 ms, branch-free engine 49.8 ms per frame.
    constexpr bool kPred = kSmem && VPT_DDA_PRED;
    int linB = parkLin; // kPred: the voxel before lin (lin - linB = the last step; equal: no step taken yet)
    int lastD = 0;      // select form: the last step's stride
    (void)lastD; (void)linB;
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
                linB = lin; lastD = 0; // (both forms of "no step taken yet")
                live = true;
                ++raysAcc;
            }
            chunkPos += min((unsigned)__popc(idle), avail);
            idle = __ballot_sync(kFull, !live);
        }
        if (idle == kFull) break; // nothing live and nothing left to take

        // ---- step loop
        if (kPred)
        {
        // Two voxel registers that swap roles every step (linB = step(lin), lin = step(linB)): the voxel before the current one
        // — what the entry face of a hit follows from — is then simply the other register, at no instruction per step.
        for (;;)
        {
            bool leave = false;
#pragma unroll
            for (int u = 0; u < (kClosest ? VPT_DDA_UNROLL : VPT_DDA_UNROLL_ANY) / 2; ++u)
            {
#define VPT_DDA_LOADWORD(L) \
    (kSmem ? (VPT_DDA_BYTE ? repByte(reinterpret_cast<const uint8_t *>(occS)[(unsigned)(L) >> 3]) : occS[(unsigned)(L) >> 5]) \
           : __ldg(occG + ((unsigned)(L) >> 5)))
// advance the axis with the smallest tMax (ties: Z, then Y, then X = X<Y ? (X<Z ? X : Z) : (Y<Z ? Y : Z)); TO = FROM + its step
#define VPT_DDA_ADVANCE(FROM, TO)                                                                  \
    do {                                                                                           \
        asm("{\n\t.reg .pred pz, pnz, py, px;\n\t"                                                \
            "min.f32 %0, %2, %3, %4;\n\t"                                                          \
            "setp.eq.f32 pz|pnz, %4, %0;\n\t"                                                      \
            "setp.eq.and.f32 py|px, %3, %0, pnz;\n\t"                                              \
            "@px add.rn.f32 %2, %2, %5;\n\t"                                                       \
            "@py add.rn.f32 %3, %3, %6;\n\t"                                                       \
            "@pz add.rn.f32 %4, %4, %7;\n\t"                                                       \
            "@px add.s32 %1, %8, %9;\n\t"                                                          \
            "@py add.s32 %1, %8, %10;\n\t"                                                         \
            "@pz add.s32 %1, %8, %11;\n\t}"                                                        \
            : "=f"(tCur), "=r"(TO), "+f"(tX), "+f"(tY), "+f"(tZ)                                   \
            : "f"(dtX), "f"(dtY), "f"(dtZ), "r"(FROM), "r"(dX), "r"(dY), "r"(dZ));                 \
        if (kStats) stepsAcc += live ? 1u : 0u;                                                    \
    } while (0)
// mask words are bit-reversed: voxel k of a word sits at bit 31-k. A set bit = solid voxel or shell: once per ray (twice for
// a ray whose origin voxel lies before tmin). Leaving the unrolled block on it closes the divergent region once per block,
// not once per step, and a finished lane has nothing to do before the next ballot anyway.
#define VPT_DDA_TEST(CUR, PRV, SWAPBACK)                                                           \
    if ((int)(VPT_DDA_LOADWORD(CUR) << ((CUR) & 31)) < 0)                                          \
    {                                                                                              \
        bool fin = tCur >= tmin;                                                                   \
        if (!fin) { int x, y, z; fin = decode(CUR, x, y, z); }                                     \
        if (fin)                                                                                   \
        {                                                                                          \
            finLin = CUR; finT = tCur; finD = (CUR) - (PRV);                                       \
            pending = true;                                                                        \
            live = false;                                                                          \
            lin = parkLin; linB = parkLin; dX = 0; dY = 0; dZ = 0; /* parked: steps in place on an empty spare word */ \
        }                                                                                          \
        else                                                                                       \
        {                                                                                          \
            VPT_DDA_ADVANCE(CUR, PRV);                                                             \
            if (SWAPBACK) { const int t_ = linB; linB = lin; lin = t_; }                           \
        }                                                                                          \
        leave = true;                                                                              \
        break;                                                                                     \
    }
                VPT_DDA_TEST(lin, linB, true)
                VPT_DDA_ADVANCE(lin, linB);
                VPT_DDA_TEST(linB, lin, false)
                VPT_DDA_ADVANCE(linB, lin);
            }
            (void)leave;
            const unsigned act = __ballot_sync(kFull, live);
            if (act == 0u) break;
            if (!exhausted && __popc(act) <= kRefillBelow) break;
        }
        }
        else
        {
        for (;;)
        {
#pragma unroll
            for (int u = 0; u < (kClosest ? VPT_DDA_UNROLL : VPT_DDA_UNROLL_ANY); ++u)
            {
                uint32_t word;
                if (kSmem) word = VPT_DDA_BYTE ? repByte(reinterpret_cast<const uint8_t *>(occS)[(unsigned)lin >> 3]) : occS[(unsigned)lin >> 5];
                else word = __ldg(occG + ((unsigned)lin >> 5));
#define VPT_DDA_ADVANCE_SEL()                                                                          \
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
                    else VPT_DDA_ADVANCE_SEL();
                    break;
#endif
                }
                // advance the axis with the smallest tMax: X<Y ? (X<Z ? X : Z) : (Y<Z ? Y : Z)  ==  A = min(X,Y); A<Z ? A : Z
                VPT_DDA_ADVANCE_SEL();
            }
            const unsigned act = __ballot_sync(kFull, live);
            if (act == 0u) break;
            if (!exhausted && __popc(act) <= kRefillBelow) break;
        }
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

// ------------------------------------------------------------------------------------------------ branch-free engine
// What the profile of the kernel above says (ncu r3b): 26 warp-instructions per warp-step although the step itself is 16 —
// a warp of ~24 live rays retires about one ray per step, and every retirement runs its bookkeeping (record the hit, park the
// lane, leave the unrolled block) as a divergent region for that one lane. ddaFlatKernel has NO branch inside a block of steps:
//  * "alive" is a predicate that is threaded through the block (each step's hit test produces the next step's alive); every
//    instruction of a finished lane is predicated off, so its registers simply keep the state of the step that ended it;
//  * the step advances into a CANDIDATE voxel (nxt = cur + step) and tests that, with the two voxel registers swapping roles
//    every step: a finished lane holds the hit voxel in one and the voxel before it in the other (the entry face follows from
//    their difference); which is which follows from the sign of that difference when the ray retires;
//  * tCur = min(tX, tY, tZ) is predicated on alive as well, so it stays the entry time of the hit voxel.
// The block is one asm statement (PTX predicates cannot cross statements). Same decisions, same roundings, same results as
// ddaKernel: 16 instructions per step, nothing per retirement.
// tmin > 0 (only the bias rays of the temporal ReSTIR pass): a solid voxel entered before tmin ends the block like a hit and
// is resolved after it — not the shell: the ray walks on from it.
#ifndef VPT_DDA_FLAT
#define VPT_DDA_FLAT 1
#endif
#ifndef VPT_DDA_FLAT_BLOCK
#define VPT_DDA_FLAT_BLOCK 16 // steps per block, any-hit launches
#endif
#ifndef VPT_DDA_FLAT_BLOCK_CLOSEST
#define VPT_DDA_FLAT_BLOCK_CLOSEST 8
#endif
#ifndef VPT_DDA_FLAT_CLOSEST
#define VPT_DDA_FLAT_CLOSEST 0 // closest-hit launches (coherent primary rays) on the branch-free engine too
#endif
// %0-%2 tX tY tZ, %3 tCur, %4 %5 the two voxel registers, %6 alive, %7 step counter; %8-%10 tDelta, %11-%13 voxel strides,
// %14 mask base (shared).
// Written the way ptxas keeps it (it turns a predicated min into min + select and the second output of a two-output setp into a
// PLOP3): one FSETP per predicate, each ANDed with alive; the next step's alive comes straight out of the mask test (ISETP.GE.AND).
// TC = per-instance extras: the tCur commit, only in the instances that look at the hit time (closest hits, rays with a near or
// far end); the step counter of vpt_get_counters, only while statistics are on.
#define VPT_FS(PA, PN, CUR, NXT, TC, LD)                           \
    "min.f32 tn, %0, %1, %2;\n\t"                                  \
    TC(PA)                                                         \
    "setp.eq.and.f32 pz, %2, tn, " PA ";\n\t"                      \
    "setp.neu.and.f32 pnz, %2, tn, " PA ";\n\t"                    \
    "setp.eq.and.f32 py, %1, tn, pnz;\n\t"                         \
    "setp.neu.and.f32 px, %1, tn, pnz;\n\t"                        \
    "@px add.s32 " NXT ", " CUR ", %11;\n\t"                       \
    "@py add.s32 " NXT ", " CUR ", %12;\n\t"                       \
    "@pz add.s32 " NXT ", " CUR ", %13;\n\t"                       \
    "@px add.rn.f32 %0, %0, %8;\n\t"                               \
    "@py add.rn.f32 %1, %1, %9;\n\t"                               \
    "@pz add.rn.f32 %2, %2, %10;\n\t"                              \
    LD(NXT)                                                        \
    "mul.lo.u32 wb, wb, 0x01010101;\n\t"                           \
    "shf.l.wrap.b32 wb, 0, wb, " NXT ";\n\t"                       \
    "setp.ge.and.s32 " PN ", wb, 0, " PA ";\n\t"
// the mask byte of voxel NXT from the byte-swapped shared-memory copy: byte lin >> 3
#define VPT_LD_S(NXT) "shr.u32 ad, " NXT ", 3;\n\t" "add.u32 ad, ad, %14;\n\t" "ld.shared.u8 wb, [ad];\n\t"
#define VPT_TC_ON(PA) "selp.f32 %3, tn, %3, " PA ";\n\t"
#define VPT_TC_OFF(PA)
#define VPT_TC_ON_STATS(PA) VPT_TC_ON(PA) "@" PA " add.u32 %7, %7, 1;\n\t"
#define VPT_TC_OFF_STATS(PA) "@" PA " add.u32 %7, %7, 1;\n\t"
#define VPT_FS2(TC, LD) VPT_FS("p0", "p1", "%4", "%5", TC, LD) VPT_FS("p1", "p0", "%5", "%4", TC, LD)
#define VPT_FS8(TC, LD) VPT_FS2(TC, LD) VPT_FS2(TC, LD) VPT_FS2(TC, LD) VPT_FS2(TC, LD)
#define VPT_FLAT_BLOCK(R, STEPS)                                                                                     \
    asm volatile("{\n\t.reg .pred p0, p1, pz, pnz, py, px;\n\t.reg .u32 ad, wb;\n\t.reg .f32 tn;\n\t"                \
                 "setp.ne.s32 p0, %6, 0;\n\t" STEPS "selp.s32 %6, 1, 0, p0;\n\t}"                                    \
                 : "+f"(R.tX), "+f"(R.tY), "+f"(R.tZ), "+f"(R.tCur), "+r"(R.linA), "+r"(R.linB), "+r"(R.alive), "+r"(R.steps) \
                 : "f"(R.dtX), "f"(R.dtY), "f"(R.dtZ), "r"(R.dX), "r"(R.dY), "r"(R.dZ), "r"(base))
#if VPT_DDA_FLAT_BLOCK == 32
#define VPT_FS_ANY(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD)
#elif VPT_DDA_FLAT_BLOCK == 24
#define VPT_FS_ANY(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD)
#elif VPT_DDA_FLAT_BLOCK == 8
#define VPT_FS_ANY(TC, LD) VPT_FS8(TC, LD)
#else
#define VPT_FS_ANY(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD)
#endif
#if VPT_DDA_FLAT_BLOCK_CLOSEST == 8
#define VPT_FS_CLOSEST(TC, LD) VPT_FS8(TC, LD)
#elif VPT_DDA_FLAT_BLOCK_CLOSEST == 16
#define VPT_FS_CLOSEST(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD)
#elif VPT_DDA_FLAT_BLOCK_CLOSEST == 24
#define VPT_FS_CLOSEST(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD)
#else
#define VPT_FS_CLOSEST(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD) VPT_FS8(TC, LD)
#endif

// kRays rays per lane: each is an independent dependency chain (min -> compare -> add -> LDS -> test -> next step's alive), and a
// block of steps is straight-line code, so ptxas interleaves the lane's rays and one warp covers its own latencies the way two
// warps would — the shared-memory masks leave room for one 1024-thread CTA per SM, this doubles the chains in flight.
#ifndef VPT_DDA_RAYS
#define VPT_DDA_RAYS 1 // measured: 2 rays per lane 0.816 vs 0.793 ms of DDA per frame (a warp of 64 slots refills less often; the engine is alu/issue bound, not latency bound)
#endif
struct LaneRay
{
    float tX, tY, tZ, dtX, dtY, dtZ, tCur, tmin, tmax;
    int linA, linB, dX, dY, dZ;
    int alive;      // the ray is still walking
    bool armed;     // the slot holds a ray (walking, or finished and not yet retired)
    uint32_t meta, result;
    unsigned steps;
};

template <bool kClosest, bool kTmax, bool kTmin, bool kStats, int kRays>
__global__ void __launch_bounds__(kDdaThreads, 1) ddaFlatKernel(const __grid_constant__ DdaArgs a)
{
    extern __shared__ uint32_t occS[];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.grid.occ);
        uint4 *dst = reinterpret_cast<uint4 *>(occS);
        const int n4 = a.grid.occWords >> 2; // occWords is a multiple of 4
        for (int i = threadIdx.x; i < n4; i += blockDim.x)
        {
            uint4 w = __ldg(src + i);
            // byte-swapped: voxel k of a word (bit 31-k) then lives in byte k >> 3 of the word, i.e. at byte address lin >> 3
            w.x = __byte_perm(w.x, 0u, 0x0123u); w.y = __byte_perm(w.y, 0u, 0x0123u); w.z = __byte_perm(w.z, 0u, 0x0123u); w.w = __byte_perm(w.w, 0u, 0x0123u);
            dst[i] = w;
        }
        __syncthreads();
    }
    const uint8_t *occB = reinterpret_cast<const uint8_t *>(occS);
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(occS);
    auto solid = [&](int l) -> bool { return ((uint32_t)occB[(unsigned)l >> 3] << (l & 7)) & 0x80u; };
    const unsigned count = __ldg(a.count);
    const unsigned lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const int Wp = a.grid.Wp, Dp = a.grid.Dp, W = a.grid.W, H = a.grid.H, D = a.grid.D;
    const int strideY = Wp * Dp;
    const int parkLin = a.grid.parkLin, maskBits = a.grid.maskWords * 32, upH = a.grid.upH;
    auto decode = [&](int l, int &x, int &y, int &z) -> bool {
        const bool up = l >= maskBits;
        if (up) l -= maskBits;
        uint32_t r, xp, yp, zp;
        a.grid.divWp.div((uint32_t)l, r, xp);
        a.grid.divDp.div(r, yp, zp);
        x = (int)xp - 1; y = (int)yp - 1; z = (int)zp - 1;
        return (unsigned)x >= (unsigned)W || (unsigned)z >= (unsigned)D || (unsigned)y >= (unsigned)(up ? upH : H);
    };

    LaneRay rays[kRays];
#pragma unroll
    for (int k = 0; k < kRays; ++k)
    {
        LaneRay &R = rays[k];
        R.tX = R.tY = R.tZ = R.dtX = R.dtY = R.dtZ = R.tCur = R.tmin = 0.0f; R.tmax = kRayMax;
        R.linA = R.linB = parkLin; R.dX = R.dY = R.dZ = 0; R.alive = 0; R.armed = false; R.meta = R.result = 0; R.steps = 0;
    }
    unsigned chunkPos = 0, chunkEnd = 0;
    bool exhausted = (count == 0);
    unsigned raysAcc = 0;
    // a finished slot holds the voxel that ended it in one register and the voxel before it in the other: the later one is a
    // forward step (dX, dY or dZ, signed) ahead. (Not "the solid one": a ray with a near end may have walked through solid voxels.)
    auto aheadA = [&](const LaneRay &R) -> bool { const int d = R.linA - R.linB; return d == 0 || d == R.dX || d == R.dY || d == R.dZ; };

    for (;;)
    {
#pragma unroll
        for (int k = 0; k < kRays; ++k)
        {
            LaneRay &R = rays[k];
            // ---- retire finished rays (all idle lanes together, convergent)
            if (R.armed && !R.alive)
            {
                const bool hitA = aheadA(R);
                const int finLin = hitA ? R.linA : R.linB, finD = hitA ? R.linA - R.linB : R.linB - R.linA;
                int x, y, z;
                // a solid voxel first met at or beyond the ray's far end is no hit (see ddaKernel)
                const bool shell = decode(finLin, x, y, z) || (kTmax && R.tCur >= R.tmax);
                if (kClosest)
                {
                    uint32_t packed = kHitMiss;
                    if (!shell)
                    {
                        uint32_t face = (R.meta >> 4) & 7u;
                        if (finD != 0)
                        {
                            const int ax = finD < 0 ? -finD : finD;
                            if (ax == 1) face = (R.meta & 1u) ? 2u : 3u;
                            else if (ax == strideY) face = (R.meta & 2u) ? 1u : 0u;
                            else face = (R.meta & 4u) ? 5u : 4u;
                        }
                        packed = ((uint32_t)((y * D + z) * W + x) << 3) | face;
                    }
                    a.hitT[R.result] = shell ? kRayMax : R.tCur;
                    a.hitPacked[R.result] = packed;
                }
                else
                    a.vis[R.result] = shell ? (uint8_t)0 : (uint8_t)1;
                R.armed = false;
            }
            // ---- re-arm idle slots
            unsigned idle = __ballot_sync(kFull, !R.armed);
            while (idle != 0 && !exhausted)
            {
                if (chunkPos >= chunkEnd)
                {
                    unsigned b0 = 0;
                    if (lane == 0) b0 = atomicAdd(a.cursor, (unsigned)kChunk);
                    b0 = __shfl_sync(kFull, b0, 0);
                    if (b0 >= count) { exhausted = true; break; }
                    chunkPos = b0;
                    chunkEnd = min(b0 + (unsigned)kChunk, count);
                }
                const unsigned avail = chunkEnd - chunkPos;
                const unsigned rank = __popc(idle & ltMask);
                const bool take = !R.armed && rank < avail;
                if (take)
                {
                    const uint4 *q = a.queue + (size_t)(chunkPos + rank) * 3;
                    const uint4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
                    R.tX = __uint_as_float(q0.x); R.tY = __uint_as_float(q0.y); R.tZ = __uint_as_float(q0.z); R.tCur = __uint_as_float(q0.w);
                    R.dtX = __uint_as_float(q1.x); R.dtY = __uint_as_float(q1.y); R.dtZ = __uint_as_float(q1.z); if (kTmin) R.tmin = __uint_as_float(q1.w);
                    R.linA = (int)q2.x; R.linB = R.linA; R.meta = q2.y; R.result = q2.z; if (kTmax) R.tmax = __uint_as_float(q2.w);
                    R.dX = (R.meta & 1u) ? 1 : -1;
                    R.dY = (R.meta & 2u) ? strideY : -strideY;
                    R.dZ = (R.meta & 4u) ? Wp : -Wp;
                    // a solid start voxel is the hit (entry face from the prepared ray) unless the ray has not reached tmin yet
                    R.alive = (solid(R.linA) && !(kTmin && R.tCur < R.tmin)) ? 0 : 1;
                    R.armed = true;
                    ++raysAcc;
                }
                chunkPos += min((unsigned)__popc(idle), avail);
                idle = __ballot_sync(kFull, !R.armed);
            }
        }
        {
            bool any = false;
#pragma unroll
            for (int k = 0; k < kRays; ++k) any = any || rays[k].armed;
            if (!__any_sync(kFull, any)) break; // nothing armed and nothing left to take
        }

        // ---- step loop
        for (;;)
        {
#pragma unroll
            for (int k = 0; k < kRays; ++k)
            {
                LaneRay &R = rays[k];
#define VPT_FLAT_RUN(BLOCK, LD)                                                                  \
    do {                                                                                         \
        if (kStats)                                                                              \
        {                                                                                        \
            if (kClosest) BLOCK(R, VPT_FS_CLOSEST(VPT_TC_ON_STATS, LD));                         \
            else if (kTmax || kTmin) BLOCK(R, VPT_FS_ANY(VPT_TC_ON_STATS, LD));                  \
            else BLOCK(R, VPT_FS_ANY(VPT_TC_OFF_STATS, LD));                                     \
        }                                                                                        \
        else                                                                                     \
        {                                                                                        \
            if (kClosest) BLOCK(R, VPT_FS_CLOSEST(VPT_TC_ON, LD));                               \
            else if (kTmax || kTmin) BLOCK(R, VPT_FS_ANY(VPT_TC_ON, LD));                        \
            else BLOCK(R, VPT_FS_ANY(VPT_TC_OFF, LD));                                           \
        }                                                                                        \
    } while (0)
                VPT_FLAT_RUN(VPT_FLAT_BLOCK, VPT_LD_S);
            }
            unsigned nAlive = 0;
#pragma unroll
            for (int k = 0; k < kRays; ++k)
            {
                LaneRay &R = rays[k];
                if (kTmin)
                {
                    // ended on a solid voxel entered before tmin: not a hit unless it is the shell — walk on from it
                    const bool early = R.armed && !R.alive && R.tCur < R.tmin;
                    if (__any_sync(kFull, early) && early)
                    {
                        const int cur = aheadA(R) ? R.linA : R.linB;
                        int x, y, z;
                        if (!decode(cur, x, y, z)) { R.linA = cur; R.linB = cur; R.alive = 1; }
                    }
                }
                nAlive += __popc(__ballot_sync(kFull, R.alive != 0));
            }
            if (nAlive == 0u) break;
            if (!exhausted && nAlive <= (unsigned)(kRefillBelow * kRays)) break;
        }
    }
    unsigned long long r64 = raysAcc, s64 = 0;
#pragma unroll
    for (int k = 0; k < kRays; ++k) s64 += rays[k].steps;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        r64 += __shfl_down_sync(kFull, r64, off);
        if (kStats) s64 += __shfl_down_sync(kFull, s64, off);
    }
    if (lane == 0 && r64) { atomicAdd(a.counters + 0, r64); atomicAdd(a.counters + 2, r64); if (kStats) atomicAdd(a.counters + 1, s64); } // [2]: running total over frames
}

template <bool kClosest, bool kTmax, bool kTmin, bool kStats>
static cudaError_t launchFlatK(const DdaArgs &a, cudaStream_t s, int smCount)
{
    const size_t smem = (size_t)a.grid.occWords * 4;
    cudaError_t e = cudaFuncSetAttribute(ddaFlatKernel<kClosest, kTmax, kTmin, kStats, VPT_DDA_RAYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ddaFlatKernel<kClosest, kTmax, kTmin, kStats, VPT_DDA_RAYS><<<smCount, kDdaThreads, smem, s>>>(a);
    return cudaGetLastError();
}
template <bool kClosest, bool kTmax, bool kTmin>
static cudaError_t launchFlatT(const DdaArgs &a, bool stats, cudaStream_t s, int smCount)
{
    return stats ? launchFlatK<kClosest, kTmax, kTmin, true>(a, s, smCount) : launchFlatK<kClosest, kTmax, kTmin, false>(a, s, smCount);
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

cudaError_t launchDda(const DdaArgs &a, bool closest, bool occInSmem, bool countSteps, bool farEnd, bool nearEnd, cudaStream_t s, int smCount)
{
#if VPT_DDA_FLAT
    // the branch-free engine: every any-hit launch whose masks are in shared memory (closest hits only with VPT_DDA_FLAT_CLOSEST).
    // Worlds walked through L1/L2 stay on ddaKernel: a global-memory instance of the block measured 49.8 vs 48.3 ms per frame on the
    // cfg5-shaped trace (tools/cfg5_quick.py; the finished lanes' loads are no longer free there) and was removed.
    if (occInSmem && !(closest && !VPT_DDA_FLAT_CLOSEST))
    {
        if (closest) return launchFlatT<true, false, false>(a, countSteps, s, smCount);
        if (farEnd) return nearEnd ? launchFlatT<false, true, true>(a, countSteps, s, smCount) : launchFlatT<false, true, false>(a, countSteps, s, smCount);
        return nearEnd ? launchFlatT<false, false, true>(a, countSteps, s, smCount) : launchFlatT<false, false, false>(a, countSteps, s, smCount);
    }
#endif
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
