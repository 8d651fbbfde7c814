"""ORACLE — test infrastructure only. numpy restatement of the reference's image-diff tool
(/root/reference/renderer/util/ImageDiff.cpp:95-372; thresholds :119-121, doc docs/image-diffing-system.md):
different-pixel count (per-channel |a-b|/255 > 0.01), RMSE over all samples, global SSIM on 3x3-gaussian
filtered luma with K1=0.01, K2=0.03, L=255; IDENTICAL / VERY CLOSE (SSIM>0.99 & RMSE<1) / CLOSE (SSIM>0.95 & RMSE<5).
Pinned against the reference's own ImageDiff.cpp (oracle/_ref/libref_imagediff.so) through tests/golden/imagediff_ref.json.
Sums are taken in float64 and rounded once where the reference accumulates serially in float32, so values agree to
~1e-4 relative, which is far inside the class thresholds."""
import numpy as np


def _gauss3(img):
    k = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], np.float32) / np.float32(16)
    p = np.pad(img, 1, mode="edge")
    out = np.zeros_like(img, dtype=np.float32)
    h, w = img.shape
    for ky in range(3):
        for kx in range(3):
            out += p[ky:ky + h, kx:kx + w] * k[ky, kx]
    return out


def _gray(img):
    img = img.astype(np.float32)
    if img.ndim == 2 or img.shape[2] == 1:
        return img.reshape(img.shape[0], img.shape[1])
    if img.shape[2] >= 3:
        return np.float32(0.299) * img[..., 0] + np.float32(0.587) * img[..., 1] + np.float32(0.114) * img[..., 2]
    return img[..., 0]


def compare(a, b):
    """a, b: uint8 arrays [H, W, C]. Returns the fields of ImageDiffResult."""
    a = np.asarray(a, np.uint8)
    b = np.asarray(b, np.uint8)
    assert a.shape[:2] == b.shape[:2]
    c = min(a.shape[2], b.shape[2])
    fa, fb = a[..., :c].astype(np.float32), b[..., :c].astype(np.float32)
    diff = np.abs(fa - fb) / np.float32(255.0)
    different = int((diff > np.float32(0.01)).any(-1).sum())
    total = a.shape[0] * a.shape[1]
    d64 = fa.astype(np.float64) - fb.astype(np.float64)
    rmse = float(np.sqrt((d64 * d64).sum() / d64.size))
    ga, gb = _gauss3(_gray(a)), _gauss3(_gray(b))
    C1, C2 = (0.01 * 255.0) ** 2, (0.03 * 255.0) ** 2
    ma, mb = float(ga.astype(np.float64).mean()), float(gb.astype(np.float64).mean())
    n = ga.size
    va = float(((ga - ma).astype(np.float64) ** 2).sum() / (n - 1))
    vb = float(((gb - mb).astype(np.float64) ** 2).sum() / (n - 1))
    cov = float(((ga - ma).astype(np.float64) * (gb - mb).astype(np.float64)).sum() / (n - 1))
    ssim = ((2 * ma * mb + C1) * (2 * cov + C2)) / ((ma * ma + mb * mb + C1) * (va + vb + C2))
    return dict(differentPixels=different, totalPixels=total, pixelDifferenceRatio=different / total, rmse=rmse, ssim=float(ssim),
                isIdentical=different == 0, isVeryClose=(ssim > 0.99 and rmse < 1.0), isClose=(ssim > 0.95 and rmse < 5.0))


def to_png8(hdr_rgb, exposure=0.8):
    """Deterministic 8-bit conversion used to apply the reference's PNG thresholds to linear HDR output
    (the reference's own tone-mapper is wall-clock dependent and out of scope): x*exposure -> x/(1+x) -> ^(1/2.2) -> *255,
    y flipped like OfflineBackend::writeFrameBufferToPNG (OfflineBackend.cpp:191-221)."""
    v = np.maximum(hdr_rgb[..., :3].astype(np.float32) * np.float32(exposure), 0)
    v = v / (1 + v)
    v = np.clip(v ** np.float32(1 / 2.2), 0, 1)
    return (v[::-1] * 255.0).astype(np.uint8)
