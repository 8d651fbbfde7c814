// Launch-invariant unsigned division (no CUDA headers: shared with the host-side unit-test hook).
#pragma once
#include <stdint.h>

namespace vpt {

// Unsigned division by a launch-invariant divisor without the ~20-instruction runtime division sequence
// (magic multiply: t = mulhi(m, n); q = (((n - t) >> 1) + t) >> s; a plain shift for powers of two).
struct FastDiv
{
    uint32_t d, m, s, pow2;
#if defined(__CUDACC__)
    __host__ __device__
#endif
    void div(uint32_t n, uint32_t &q, uint32_t &r) const
    {
#if defined(__CUDA_ARCH__)
        const uint32_t t = __umulhi(m, n);
#else
        const uint32_t t = (uint32_t)(((uint64_t)m * n) >> 32);
#endif
        q = pow2 ? (n >> s) : ((((n - t) >> 1) + t) >> s);
        r = n - q * d;
    }
};
inline FastDiv makeFastDiv(uint32_t d)
{
    FastDiv f;
    f.d = d; f.m = 0; f.s = 0; f.pow2 = 0;
    uint32_t fl = 0;
    while ((2u << fl) <= d && fl < 31) ++fl; // floor(log2 d)
    if ((d & (d - 1)) == 0) { f.pow2 = 1; f.s = fl; return f; }
    const uint64_t num = (uint64_t)1 << (32 + fl);
    uint64_t pm = num / d;
    const uint64_t rem = num % d;
    pm += pm;
    const uint64_t twice = rem + rem;
    if (twice >= d) pm += 1;
    f.m = (uint32_t)(pm + 1);
    f.s = fl;
    return f;
}

} // namespace vpt
