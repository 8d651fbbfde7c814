// Prepared rays: the DDA set-up of VoxelEngine::performRayTraversal (/root/reference/voxelengine/VoxelEngine.cu:1040-1131:
// floor of the origin, step signs, tDelta = 1/|d| with the 1e-8 guard, tMax = (nextBoundary - o)/d), generalised
// with an entry clip for origins outside the grid, computed ONCE by the stage that spawns the ray —
// at full lane occupancy — and queued as 48 bytes. The DDA kernel (vpt_dda.cu) then only steps.
// Exact arithmetic class (explicit round-to-nearest intrinsics): bit-identical to the CPU oracle.
#pragma once
#include "vpt_kernels.h"
#include "vpt_math.cuh"

namespace vpt {

struct PreparedRay // 3 x 16 bytes
{
    float tMaxX, tMaxY, tMaxZ, tCur;
    float tDeltaX, tDeltaY, tDeltaZ, tmin;
    int lin;         // padded linear voxel index of the start voxel
    uint32_t meta;   // bits 0..2: step is +1 along x,y,z; bits 4..6: face id if the start voxel is the hit (6 = none)
    uint32_t result; // slot the result is written to
    float tmax;      // far end (visibility rays towards a local light stop 0.01 before it, closesthit.cu:616-617); kRayMax = none
};
constexpr uint32_t kHitMiss = 0xFFFFFFFFu;

// Returns false when the ray never enters the grid (the caller records a miss itself).
VPT_DEV bool prepareRay(const GridView &g, f3 o, f3 d, float tmin, uint32_t result, PreparedRay &r, float tmax = kRayMax)
{
    const int W = g.W, H = g.H, D = g.D;
    int x = (int)floorf(o.x), y = (int)floorf(o.y), z = (int)floorf(o.z);
    int hitAxis = -1;
    float tCur = 0.0f;
    if (x < 0 || x >= W || y < 0 || y >= H || z < 0 || z >= D)
    {
        const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
        const float dim[3] = {(float)W, (float)H, (float)D};
        float tEnter = -FLT_MAX, tExit = FLT_MAX;
        int axis = -1;
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            if (fabsf(dd[a]) < 1e-8f)
            {
                if (oo[a] < 0.0f || oo[a] >= dim[a]) return false;
                continue;
            }
            float ta = ex::divf(ex::subf(0.0f, oo[a]), dd[a]), tb = ex::divf(ex::subf(dim[a], oo[a]), dd[a]);
            float tn = fminr(ta, tb), tf = fmaxr(ta, tb);
            if (tn > tEnter) { tEnter = tn; axis = a; }
            if (tf < tExit) tExit = tf;
        }
        if (axis < 0 || tEnter > tExit || tExit < 0.0f || tEnter < 0.0f) return false;
        f3 p = ex::pointAt(o, d, tEnter);
        x = clampi((int)floorf(p.x), 0, W - 1);
        y = clampi((int)floorf(p.y), 0, H - 1);
        z = clampi((int)floorf(p.z), 0, D - 1);
        if (axis == 0) x = dd[0] > 0.0f ? 0 : W - 1;
        if (axis == 1) y = dd[1] > 0.0f ? 0 : H - 1;
        if (axis == 2) z = dd[2] > 0.0f ? 0 : D - 1;
        hitAxis = axis;
        tCur = tEnter;
    }
    const bool px = d.x > 0.0f, py = d.y > 0.0f, pz = d.z > 0.0f;
    const bool zx = fabsf(d.x) < 1e-8f, zy = fabsf(d.y) < 1e-8f, zz = fabsf(d.z) < 1e-8f;
    r.tDeltaX = zx ? FLT_MAX : ex::divf(1.0f, fabsf(d.x));
    r.tDeltaY = zy ? FLT_MAX : ex::divf(1.0f, fabsf(d.y));
    r.tDeltaZ = zz ? FLT_MAX : ex::divf(1.0f, fabsf(d.z));
    const float nbX = px ? (float)(x + 1) : (float)x;
    const float nbY = py ? (float)(y + 1) : (float)y;
    const float nbZ = pz ? (float)(z + 1) : (float)z;
    r.tMaxX = zx ? FLT_MAX : ex::divf(ex::subf(nbX, o.x), d.x);
    r.tMaxY = zy ? FLT_MAX : ex::divf(ex::subf(nbY, o.y), d.y);
    r.tMaxZ = zz ? FLT_MAX : ex::divf(ex::subf(nbZ, o.z), d.z);
    r.tCur = tCur;
    r.tmin = tmin;
    // rays that climb (dir.y > 0) walk the upward mask (solid from GridView::upH up): same hits, earlier exits
    r.lin = ((y + 1) * g.Dp + (z + 1)) * g.Wp + (x + 1) + (py ? g.maskWords * 32 : 0);
    uint32_t face0 = 6u;
    if (hitAxis == 0) face0 = px ? 2u : 3u;
    else if (hitAxis == 1) face0 = py ? 1u : 0u;
    else if (hitAxis == 2) face0 = pz ? 5u : 4u;
    r.meta = (px ? 1u : 0u) | (py ? 2u : 0u) | (pz ? 4u : 0u) | (face0 << 4);
    r.result = result;
    r.tmax = tmax;
    return true;
}

VPT_DEV void storePreparedRay(uint4 *queue, unsigned pos, const PreparedRay &r)
{
    uint4 *q = queue + (size_t)pos * 3;
    q[0] = make_uint4(__float_as_uint(r.tMaxX), __float_as_uint(r.tMaxY), __float_as_uint(r.tMaxZ), __float_as_uint(r.tCur));
    q[1] = make_uint4(__float_as_uint(r.tDeltaX), __float_as_uint(r.tDeltaY), __float_as_uint(r.tDeltaZ), __float_as_uint(r.tmin));
    q[2] = make_uint4((uint32_t)r.lin, r.meta, r.result, __float_as_uint(r.tmax));
}

} // namespace vpt
