// TMA (cp.async.bulk.tensor) + mbarrier plumbing for the shared-memory tile kernels of the denoiser (vpt_dn_tiles.cu).
//
// A frame plane is a dense row-major array. A float4 plane of W x H pixels is described to the TMA unit as a 2-D FLOAT32 tensor
// of {4W, H} elements (a uint32 / float plane as {W, H}); one `cp.async.bulk.tensor.2d` then lands a (tile + halo) box in shared
// memory as a dense [rows][cols] array, signalling an mbarrier with the byte count. Coordinates are signed and elements outside
// the tensor are ZERO-filled, which is exactly what the edge-stopping stencils want where they give out-of-image taps zero
// weight; kernels with clamp-to-edge semantics (cudaBoundaryModeClamp in the reference) patch the border tiles in shared memory.
// SASS evidence: UTMALDG + SYNCS (profiles/).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vpt {
namespace tma {

// ---- host: tensor-map encoding through the driver entry point (libvpt.so does not link libcuda)
// dimX / boxX in ELEMENTS (4-byte), rowPitchBytes a multiple of 16, boxX * 4 a multiple of 16, boxX, boxY <= 256.
cudaError_t encode2D(CUtensorMap *map, bool asUint32, const void *base, uint64_t dimX, uint64_t dimY, uint64_t rowPitchBytes, uint32_t boxX, uint32_t boxY);

#ifdef __CUDACC__
// ---- device
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void barrierInit(uint64_t *bar, unsigned arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); // make the init visible to the async proxy
}
__device__ __forceinline__ void barrierExpectTx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
// box whose first element is (c0, c1) of the tensor -> dst (128-byte aligned shared memory), completion counted on `bar`
__device__ __forceinline__ void load2D(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smemAddr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void prefetchDescriptor(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// Wait for the phase with the given parity. try_wait suspends the warp in hardware up to a time limit, so this is not a busy
// spin. A transfer that never completes (a descriptor / byte-count bug) must not hang the GPU: after ~1 M probes the wait gives
// up, counts the event in *timeouts (read back by vpt_debug_tma_timeouts) and returns false.
__device__ __forceinline__ bool barrierWait(uint64_t *bar, unsigned parity, unsigned *timeouts)
{
    unsigned done = 0, spins = 0;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smemAddr(bar)), "r"(parity)
                     : "memory");
        if (!done && ++spins > (1u << 20)) { atomicAdd(timeouts, 1u); return false; }
    } while (!done);
    return true;
}
#endif

} // namespace tma
} // namespace vpt
