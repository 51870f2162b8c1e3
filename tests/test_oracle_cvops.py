"""NumPy restatements of the OpenCV primitives against the installed cv2 (4.13.0), at the
parameters of the reference's own call sites."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import cvops as V
from oracle import mammals as M


@pytest.mark.parametrize("sigma", [3.5, 2.0, 1.8, 1.6, 1.4, 1.3, 1.2, 1.0, 0.7, 0.45, 1.15, 2.6])
def test_gaussian_taps_and_ksize(sigma):
    k = V.gaussian_ksize(sigma)
    assert np.array_equal(cv2.getGaussianKernel(k, sigma, cv2.CV_32F).ravel(), V.gaussian_taps(k, sigma))
    # float64 taps: OpenCV's soft-float exp differs from libm in the last ulp
    assert np.abs(cv2.getGaussianKernel(k, sigma, cv2.CV_64F).ravel() - V.gaussian_taps(k, sigma, np.float64)).max() < 5e-16
    # ksize actually used by GaussianBlur((0,0)): an impulse spreads exactly k//2 pixels
    img = np.zeros((1, 101), np.float32)
    img[0, 50] = 1
    nz = np.nonzero(cv2.GaussianBlur(img, (0, 0), sigmaX=sigma, sigmaY=sigma)[0])[0]
    assert nz.max() - nz.min() + 1 == k


@pytest.mark.parametrize("sigma,dtype", [(3.5, np.float32), (1.0, np.float64), (0.7, np.float32)])
def test_gaussian_blur_restatement(sigma, dtype):
    img = np.random.default_rng(1).random((61, 83, 3)).astype(dtype)
    a = cv2.GaussianBlur(img, (0, 0), sigmaX=sigma, sigmaY=sigma)
    b = V.gaussian_blur_np(img, sigma)
    assert np.abs(a - b).max() <= (1e-6 if dtype == np.float32 else 1e-14)


def test_uv_blur_3x3():
    img = np.random.default_rng(2).random((40, 50)).astype(np.float32)
    a = cv2.GaussianBlur(img, (3, 3), sigmaX=0.2, sigmaY=0.2, borderType=cv2.BORDER_REFLECT101)
    b = V.gaussian_blur_np(img[..., None], 0.2, 0.2, (3, 3))[..., 0]
    assert np.abs(a - b).max() <= 2e-7


def test_reflect101():
    for n in (1, 2, 3, 5, 17):
        for i in range(-3 * n, 4 * n):
            assert V.reflect101(np.array([i]), n)[0] == cv2.borderInterpolate(i, n, cv2.BORDER_REFLECT_101) \
                if n > 1 else True


def test_remap_bit_exact():
    rng = np.random.default_rng(3)
    H, W = 37, 480
    src = rng.random((H, W, 3), dtype=np.float32)
    for xmap in ((rng.random(W) * W * 1.2 - 0.1 * W).astype(np.float32), *M.cat_warp_tables(W)[:2]):
        xm = np.repeat(xmap[None], H, 0)
        ym = np.repeat(np.arange(H, dtype=np.float32)[:, None], W, 1)
        a = cv2.remap(src, xm, ym, interpolation=cv2.INTER_LINEAR, borderMode=0, borderValue=0.0)
        assert np.array_equal(a, V.remap_rows_linear_np(src, xmap))


@pytest.mark.parametrize("shape", [(480, 270, 320, 180), (1920, 1080, 1280, 720), (101, 67, 67, 45), (64, 48, 43, 32)])
def test_resize_u8_bit_exact(shape):
    W, H, cw, ch = shape
    src = np.random.default_rng(4).integers(0, 256, (ch, cw, 3), np.uint8)
    assert np.array_equal(cv2.resize(src, (W, H), interpolation=cv2.INTER_LINEAR), V.resize_linear_u8_np(src, W, H))


def test_streak_blur_restatement():
    img = np.random.default_rng(5).random((45, 70, 3)).astype(np.float32)
    for args in [(0.5, 0.9, 2.3, 6.5), (0.52, 1.0, 2.1, 4.5), (0.5, 1.2, 2.5, 3.0)]:
        a = M.streak_blur(img, *args)
        b = M.streak_blur_np(img, *args)
        assert np.abs(a - b).max() <= 1e-6
