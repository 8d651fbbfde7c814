// C shim over the REFERENCE's own ImageDiff (renderer/util/ImageDiff.{h,cpp} + renderer/ext/stb), compiled
// from /root/reference by `make ref` into oracle/_ref/libref_imagediff.so. Test infrastructure only: used to
// pin oracle/imagediff.py (the travelling numpy restatement) against the real thing.
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "ext/stb/stb_image.h"
#include "ext/stb/stb_image_write.h"
#include "util/ImageDiff.h"
extern "C" int ref_imagediff(const unsigned char *a, const unsigned char *b, int w, int h, int channels, float *out /*rmse, ssim, ratio*/, int *flags /*diffPixels, identical, veryClose, close*/)
{
    ImageData A, B;
    A.width = B.width = w; A.height = B.height = h; A.channels = B.channels = channels;
    A.data.assign(a, a + (size_t)w * h * channels);
    B.data.assign(b, b + (size_t)w * h * channels);
    ImageDiffResult r = ImageDiff::compare(A, B);
    out[0] = r.rmse; out[1] = r.ssim; out[2] = r.pixelDifferenceRatio;
    flags[0] = r.differentPixels; flags[1] = r.isIdentical; flags[2] = r.isVeryClose; flags[3] = r.isClose;
    return 0;
}
