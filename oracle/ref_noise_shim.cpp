// C shim over the REFERENCE's own PerlinNoiseGenerator (voxelengine/Noise.{h,cpp} + ext/PerlinNoise.hpp),
// compiled from /root/reference by `make ref` into oracle/_ref/libref_noise.so. Test infrastructure only:
// used to pin the oracle's Perlin restatement bit-exactly.
#include "Noise.h"
extern "C" float ref_perlin_noise(int octaves, unsigned seed, float x, float y)
{
    PerlinNoiseGenerator gen(octaves, seed);
    return gen.getNoise(x, y);
}
extern "C" void ref_perlin_noise_map(int octaves, unsigned seed, int n, const float *xs, const float *ys, float *out)
{
    PerlinNoiseGenerator gen(octaves, seed);
    for (int i = 0; i < n; ++i) out[i] = gen.getNoise(xs[i], ys[i]);
}
