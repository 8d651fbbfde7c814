// Stand-alone probe of the TMA tile load used by csrc/vpt_dn_tiles.cu: one (tile + halo) box of a float4 plane and of a uint32
// plane into shared memory, at an interior position and hanging over the image corner (negative coordinates -> zero fill),
// checked element by element. Prints one line per case; exit code 0 when every case passes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I real-time-path-tracing-voxel-blocks_b200/csrc -I include -o tools/tma_probe tools/tma_probe.cu
#include "vpt_tma.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace vpt;

namespace vpt { namespace tma {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
cudaError_t encode2D(CUtensorMap *map, bool asUint32, const void *base, uint64_t dimX, uint64_t dimY, uint64_t rowPitchBytes, uint32_t boxX, uint32_t boxY)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) { std::printf("entry point: %s q=%d\n", cudaGetErrorString(e), (int)q); return cudaErrorNotSupported; }
    const cuuint64_t dims[2] = {dimX, dimY};
    const cuuint64_t strides[1] = {rowPitchBytes};
    const cuuint32_t box[2] = {boxX, boxY};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = reinterpret_cast<EncodeTiledFn>(p)(map, asUint32 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims,
                                                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) std::printf("cuTensorMapEncodeTiled -> %d (dims %llu x %llu pitch %llu box %u x %u)\n", (int)r, (unsigned long long)dimX,
                                       (unsigned long long)dimY, (unsigned long long)rowPitchBytes, boxX, boxY);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
} }

__device__ unsigned g_timeouts = 0;

template <int BW, int BH>
__global__ void probeKernel(const __grid_constant__ CUtensorMap m4, const __grid_constant__ CUtensorMap m1, int x0, int y0, float4 *out4, uint32_t *out1)
{
    constexpr int NPX = BW * BH;
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *s4 = reinterpret_cast<float4 *>(smem);
    uint32_t *s1 = reinterpret_cast<uint32_t *>(s4 + NPX);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)NPX * 20);
    if (threadIdx.x == 0)
    {
        tma::barrierInit(bar, 1);
        tma::barrierExpectTx(bar, (unsigned)NPX * 20u);
        tma::load2D(s4, &m4, 4 * x0, y0, bar);
        tma::load2D(s1, &m1, x0, y0, bar);
    }
    __syncthreads();
    const bool ok = tma::barrierWait(bar, 0, &g_timeouts);
    if (!ok) return;
    for (int i = threadIdx.x; i < NPX; i += blockDim.x) { out4[i] = s4[i]; out1[i] = s1[i]; }
}

// the same box with one cp.async.bulk (1-D bulk copy, no descriptor) per row and plane
template <int BW, int BH>
__global__ void probeBulkKernel(const float4 *p4, const uint32_t *p1, int W, int x0, int y0, float4 *out4, uint32_t *out1)
{
    constexpr int NPX = BW * BH;
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *s4 = reinterpret_cast<float4 *>(smem);
    uint32_t *s1 = reinterpret_cast<uint32_t *>(s4 + NPX);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)NPX * 20);
    if (threadIdx.x == 0)
    {
        tma::barrierInit(bar, 1);
        tma::barrierExpectTx(bar, (unsigned)NPX * 20u);
    }
    __syncthreads();
    if (threadIdx.x < BH)
    {
        const int r = threadIdx.x;
        const float4 *g4 = p4 + (size_t)(y0 + r) * W + x0;
        const uint32_t *g1 = p1 + (size_t)(y0 + r) * W + x0;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tma::smemAddr(s4 + r * BW)), "l"(g4), "r"(BW * 16), "r"(tma::smemAddr(bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tma::smemAddr(s1 + r * BW)), "l"(g1), "r"(BW * 4), "r"(tma::smemAddr(bar)) : "memory");
    }
    const bool ok = tma::barrierWait(bar, 0, &g_timeouts);
    if (!ok) return;
    for (int i = threadIdx.x; i < NPX; i += blockDim.x) { out4[i] = s4[i]; out1[i] = s1[i]; }
}

template <int BW, int BH> static int runBulk(int W, int H, int x0, int y0, const float4 *d4, const uint32_t *d1, const std::vector<float4> &h4, const std::vector<uint32_t> &h1)
{
    constexpr int NPX = BW * BH;
    float4 *o4; uint32_t *o1;
    cudaMalloc(&o4, NPX * 16); cudaMalloc(&o1, NPX * 4);
    cudaMemset(o4, 0xff, NPX * 16); cudaMemset(o1, 0xff, NPX * 4);
    const size_t smem = (size_t)NPX * 20 + 16;
    cudaFuncSetAttribute(probeBulkKernel<BW, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probeBulkKernel<BW, BH><<<1, 256, smem>>>(d4, d1, W, x0, y0, o4, o1);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned to = 0;
    cudaMemcpyFromSymbol(&to, g_timeouts, sizeof to);
    std::vector<float4> r4(NPX); std::vector<uint32_t> r1(NPX);
    cudaMemcpy(r4.data(), o4, NPX * 16, cudaMemcpyDeviceToHost); cudaMemcpy(r1.data(), o1, NPX * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < BH; ++j)
        for (int i = 0; i < BW; ++i)
        {
            const float4 e4 = h4[(size_t)(y0 + j) * W + x0 + i];
            const float4 g = r4[j * BW + i];
            if (g.x != e4.x || g.y != e4.y || g.z != e4.z || g.w != e4.w || r1[j * BW + i] != h1[(size_t)(y0 + j) * W + x0 + i]) ++bad;
        }
    std::printf("BULK box %dx%d at (%d,%d) of %dx%d: sync=%s timeouts=%u mismatches=%d/%d\n", BW, BH, x0, y0, W, H, cudaGetErrorString(e), to, bad, NPX);
    return (e != cudaSuccess || to != 0 || bad != 0) ? 1 : 0;
}

template <int BW, int BH> static int runCase(int W, int H, int x0, int y0, const float4 *d4, const uint32_t *d1, const std::vector<float4> &h4, const std::vector<uint32_t> &h1)
{
    constexpr int NPX = BW * BH;
    CUtensorMap m4, m1;
    if (tma::encode2D(&m4, false, d4, (uint64_t)W * 4, H, (uint64_t)W * 16, BW * 4, BH) != cudaSuccess) return 1;
    if (tma::encode2D(&m1, true, d1, (uint64_t)W, H, (uint64_t)W * 4, BW, BH) != cudaSuccess) return 1;
    float4 *o4; uint32_t *o1;
    cudaMalloc(&o4, NPX * 16); cudaMalloc(&o1, NPX * 4);
    cudaMemset(o4, 0xff, NPX * 16); cudaMemset(o1, 0xff, NPX * 4);
    const size_t smem = (size_t)NPX * 20 + 16;
    cudaFuncSetAttribute(probeKernel<BW, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probeKernel<BW, BH><<<1, 256, smem>>>(m4, m1, x0, y0, o4, o1);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned to = 0;
    cudaMemcpyFromSymbol(&to, g_timeouts, sizeof to);
    std::vector<float4> r4(NPX); std::vector<uint32_t> r1(NPX);
    cudaMemcpy(r4.data(), o4, NPX * 16, cudaMemcpyDeviceToHost); cudaMemcpy(r1.data(), o1, NPX * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < BH; ++j)
        for (int i = 0; i < BW; ++i)
        {
            const int gx = x0 + i, gy = y0 + j;
            const bool in = gx >= 0 && gy >= 0 && gx < W && gy < H;
            const float4 e4 = in ? h4[(size_t)gy * W + gx] : make_float4(0, 0, 0, 0);
            const uint32_t e1 = in ? h1[(size_t)gy * W + gx] : 0u;
            const float4 g = r4[j * BW + i];
            if (g.x != e4.x || g.y != e4.y || g.z != e4.z || g.w != e4.w || r1[j * BW + i] != e1) ++bad;
        }
    std::printf("box %dx%d at (%d,%d) of %dx%d: sync=%s timeouts=%u mismatches=%d/%d\n", BW, BH, x0, y0, W, H, cudaGetErrorString(e), to, bad, NPX);
    cudaFree(o4); cudaFree(o1);
    return (e != cudaSuccess || to != 0 || bad != 0) ? 1 : 0;
}

int main(int argc, char **argv)
{
    const int W = 256, H = 160;
    std::vector<float4> h4((size_t)W * H); std::vector<uint32_t> h1((size_t)W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) { h4[(size_t)y * W + x] = make_float4((float)x, (float)y, (float)(x + y), 1.0f); h1[(size_t)y * W + x] = (uint32_t)(y * 65536 + x + 1); }
    float4 *d4; uint32_t *d1;
    cudaMalloc(&d4, h4.size() * 16); cudaMalloc(&d1, h1.size() * 4);
    cudaMemcpy(d4, h4.data(), h4.size() * 16, cudaMemcpyHostToDevice); cudaMemcpy(d1, h1.data(), h1.size() * 4, cudaMemcpyHostToDevice);
    int fails = 0;
    const int which = argc > 1 ? std::atoi(argv[1]) : -1;
    switch (which)
    {
    case 0: fails += runCase<4, 8>(W, H, 64, 32, d4, d1, h4, h1); break;      // 64-byte inner rows
    case 1: fails += runCase<8, 8>(W, H, 64, 32, d4, d1, h4, h1); break;      // 128
    case 2: fails += runCase<16, 8>(W, H, 64, 32, d4, d1, h4, h1); break;     // 256
    case 3: fails += runCase<36, 20>(W, H, 64, 32, d4, d1, h4, h1); break;    // 576, aligned origin
    case 4: fails += runCase<36, 20>(W, H, 62, 30, d4, d1, h4, h1); break;    // 576
    case 5: fails += runCase<36, 20>(W, H, -2, -2, d4, d1, h4, h1); break;
    case 6: fails += runCase<36, 20>(W, H, 222, 142, d4, d1, h4, h1); break;
    case 7: fails += runCase<52, 52>(W, H, -10, -10, d4, d1, h4, h1); break;
    case 8: fails += runBulk<36, 20>(W, H, 64, 32, d4, d1, h4, h1); break;
    case 9: fails += runBulk<40, 24>(W, H, 28, 12, d4, d1, h4, h1); break;
    case 10: fails += runBulk<52, 52>(W, H, 52, 54, d4, d1, h4, h1); break;
    default: std::printf("usage: tma_probe <case 0..10>\n"); return 2;
    }
    std::printf(fails ? "TMA PROBE: %d case(s) FAILED\n" : "TMA PROBE: all cases pass\n", fails);
    return fails ? 1 : 0;
}
