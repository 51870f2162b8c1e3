"""Oracle: the OpenCV primitives the reference calls, restated in NumPy.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference's arithmetic for blur / warp / zoom lives in a third-party dependency that is not
under /root/reference: opencv-contrib-python, UNPINNED in requirements.txt:16; the version
installed in this image and used to pin everything here is 4.13.0.  Each `*_np` function restates
the published OpenCV algorithm (imgproc/src/smooth.dispatch.cpp, filter.simd.hpp, imgwarp.cpp,
resize.cpp); tests/test_oracle_cvops.py checks every restatement against the installed cv2 at the
reference's own call sites' parameters.  The wrappers without the `_np` suffix do what the
reference does -- call cv2 -- and fall back to the restatement only if cv2 is not importable.
"""
from __future__ import annotations

import numpy as np

try:  # the reference imports cv2 unconditionally (animal_utils.py:2)
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


# ----------------------------------------------------------------------------- Gaussian kernels
def gaussian_ksize(sigma: float, *, u8: bool = False) -> int:
    """Kernel size cv2.GaussianBlur picks for ksize=(0,0): cvRound(sigma*{3|4}*2+1)|1.

    Factor 4 for float images, 3 for CV_8U (smooth.dispatch.cpp createGaussianKernels).
    cvRound rounds half to even, like Python's round().
    """
    return int(round(sigma * (3 if u8 else 4) * 2 + 1)) | 1


def gaussian_taps(ksize: int, sigma: float, dtype=np.float32) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma>0): exp(-x^2/(2 sigma^2)) normalised in double, then cast."""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    w = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return (w / w.sum()).astype(dtype)


def reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    """cv::borderInterpolate(BORDER_REFLECT_101), valid for any overshoot (period 2n-2)."""
    idx = np.asarray(idx)
    if n == 1:
        return np.zeros_like(idx)
    p = 2 * n - 2
    m = np.mod(idx, p)
    return np.where(m >= n, p - m, m)


def sep_filter_np(img: np.ndarray, kx: np.ndarray, ky: np.ndarray) -> np.ndarray:
    """Separable correlation, BORDER_REFLECT_101, rows (x) first then columns (y).

    Accumulates in the image dtype, tap by tap.  cv2 pairs symmetric taps before multiplying, so
    results agree to a few ulp (<= 6e-7 relative, measured), not bit-for-bit.
    """
    H, W = img.shape[:2]
    rx, ry = len(kx) // 2, len(ky) // 2
    cols = reflect101(np.arange(-rx, W + rx), W)
    acc = np.zeros_like(img)
    for t, w in enumerate(kx.astype(img.dtype)):
        acc += w * img[:, cols[t:t + W]]
    rows = reflect101(np.arange(-ry, H + ry), H)
    out = np.zeros_like(img)
    for t, w in enumerate(ky.astype(img.dtype)):
        out += w * acc[rows[t:t + H]]
    return out


def gaussian_blur_np(img: np.ndarray, sigma_x: float, sigma_y: float | None = None,
                     ksize: tuple[int, int] = (0, 0)) -> np.ndarray:
    """cv2.GaussianBlur for float images, default border (REFLECT_101)."""
    sigma_y = sigma_x if not sigma_y or sigma_y <= 0 else sigma_y
    kw = ksize[0] if ksize[0] > 0 else gaussian_ksize(sigma_x)
    kh = ksize[1] if ksize[1] > 0 else gaussian_ksize(sigma_y)
    dt = np.float64 if img.dtype == np.float64 else np.float32
    return sep_filter_np(img, gaussian_taps(kw, sigma_x, dt), gaussian_taps(kh, sigma_y, dt))


def gaussian_blur(img: np.ndarray, sigma: float, ksize: tuple[int, int] = (0, 0)) -> np.ndarray:
    if cv2 is not None:
        return cv2.GaussianBlur(img, ksize, sigmaX=sigma, sigmaY=sigma,
                                borderType=cv2.BORDER_REFLECT101)
    return gaussian_blur_np(img, sigma, sigma, ksize)


# ----------------------------------------------------------------------------- remap (bilinear)
REMAP_BITS = 5            # INTER_BITS
REMAP_SCALE = 1 << REMAP_BITS


def remap_rows_linear_np(img: np.ndarray, xmap_row: np.ndarray) -> np.ndarray:
    """cv2.remap(img, xmap, ymap, INTER_LINEAR, BORDER_CONSTANT 0) for the special case the
    reference uses (cat_widevision_utils.py:84-92): ymap[y, x] == y exactly and xmap depends only
    on the column, so the warp is a per-row 1-D gather.

    OpenCV converts the float maps to fixed point first: sx = cvRound(x * 32), integer part
    sx >> 5, fraction (sx & 31) / 32; the two horizontal neighbours are blended with float
    weights (1 - f) and f, out-of-image neighbours contribute the border value 0.
    """
    H, W = img.shape[:2]
    sx = np.rint(xmap_row.astype(np.float32) * np.float32(REMAP_SCALE)).astype(np.int64)
    ix = sx >> REMAP_BITS
    f = ((sx & (REMAP_SCALE - 1)).astype(np.float32) / np.float32(REMAP_SCALE))
    w0 = (np.float32(1.0) - f)[None, :, None]
    w1 = f[None, :, None]
    ok0 = (ix >= 0) & (ix < W)
    ok1 = (ix + 1 >= 0) & (ix + 1 < W)
    a = np.where(ok0[None, :, None], img[:, np.clip(ix, 0, W - 1)], 0).astype(np.float32)
    b = np.where(ok1[None, :, None], img[:, np.clip(ix + 1, 0, W - 1)], 0).astype(np.float32)
    return a * w0 + b * w1


# ----------------------------------------------------------------------------- resize (u8 bilinear)
RESIZE_COEF_BITS = 11     # INTER_RESIZE_COEF_BITS
RESIZE_COEF_SCALE = 1 << RESIZE_COEF_BITS


def _resize_axis_tables(src: int, dst: int, *, vertical: bool):
    """Source index and 11-bit fixed-point weight pairs of cv::resize INTER_LINEAR (resize.cpp).

    Horizontal axis: an out-of-range left/right neighbour resets the fraction to 0.  Vertical
    axis: only the ROW INDICES are clamped, the weights keep their fractional split (so the first
    and last output rows blend a row with itself and pick up different truncation).
    """
    scale = 1.0 / (dst / src)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)   # float fx = (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(np.float32)
    if not vertical:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    w1 = np.rint(f * np.float32(RESIZE_COEF_SCALE)).astype(np.int64)          # saturate_cast<short>
    w0 = np.rint((np.float32(1.0) - f) * np.float32(RESIZE_COEF_SCALE)).astype(np.int64)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), w0, w1


def resize_linear_u8_np(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """cv2.resize(u8, (out_w,out_h), INTER_LINEAR): fixed-point HResize / VResize pair.

    rows: S = a*w0 + b*w1 (int32, scale 2^11);  cols: ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
    """
    H, W = img.shape[:2]
    x0, x1, ax0, ax1 = _resize_axis_tables(W, out_w, vertical=False)
    y0, y1, by0, by1 = _resize_axis_tables(H, out_h, vertical=True)
    src = img.astype(np.int64)
    rows = src[:, x0] * ax0[None, :, None] + src[:, x1] * ax1[None, :, None]
    s0, s1 = rows[y0], rows[y1]
    v = (((by0[:, None, None] * (s0 >> 4)) >> 16) + ((by1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


def resize_linear_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    if cv2 is not None:
        return cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR)
    return resize_linear_u8_np(img, out_w, out_h)
