"""Input-contract cases closed in round 2: float and wide-integer frames
for Cat and HoneyBee (reference cat.py:24,80-112; honeybee.py:106,166-173; uv_helpers.py:15-23), HoneyBee's
hsi_downsample route (uv_helpers.py:155-183) and blur sigmas beyond the fused walker (uv_helpers.py:67-73).
Golden vectors: the unmodified reference (tools/make_golden_r2.py).  CPU: the oracle reproduces them;
GPU: float data within 1e-5 relative (plus a 2e-6 floor at encoded black), integer data within 1 unit."""
import numpy as np
import pytest

import frames
from oracle import mammals as M
from oracle import uv

HW = (54, 76)
REL_TOL = 1e-5
# absolute floor near encoded black: the mappers cancel (val * (1 - sat), 1 - f * sat) and the OETF's linear segment
# multiplies what is left by 12.92, so one float32 ulp of a value near 1 (6e-8) becomes ~8e-7 of output -- the
# reference's own float32 result carries that uncertainty.  5 such ulps:
ABS_FLOOR = 4e-6
BEE_VARIANTS = {
    "default": {},
    "down10": dict(hsi_downsample=True, hsi_scale=0.1),
    "down25_falsecolor": dict(hsi_downsample=True, hsi_scale=0.25, mapping_mode="falsecolor"),
    "sigma1p5": dict(blur_sigma_px=1.5),
    "sigma2p2_gray_mixed": dict(blur_sigma_px=2.2, adaptation="gray_world", mapping_mode="falsecolor_uv_mixed"),
}


def _u8_cases():
    h, w = HW
    return [("natural", frames.natural(h, w)), ("bars", frames.bars(h, w)), ("noise", frames.noise(h, w, 2))]


def _bee_cases(vname):
    return ([] if vname == "default" else _u8_cases()) + frames.float_set(*HW)


def _close(out, ref, what, int_frac=0.03, scale=1.0):
    assert out.dtype == ref.dtype and out.shape == ref.shape, what
    if np.issubdtype(ref.dtype, np.integer):
        d = np.abs(out.astype(np.int64) - ref.astype(np.int64))
        assert d.max() <= 1 and (d > 0).mean() <= int_frac, f"{what}: max {d.max()}, {(d > 0).mean():.4f} of values differ"
    else:
        err = np.abs(out.astype(np.float64) - ref.astype(np.float64))
        bound = REL_TOL * np.abs(ref.astype(np.float64)) + ABS_FLOOR * scale
        bad = err > bound
        assert not bad.any(), (f"{what}: {int(bad.sum())} values beyond the bar, max abs err {err.max():.3e}, "
                               f"worst excess at ref={ref[np.unravel_index(np.argmax(err - bound), err.shape)]:.3e}")


# ------------------------------------------------------------------ CPU: the oracle against the reference's outputs
@pytest.mark.parametrize("vname", list(BEE_VARIANTS))
def test_oracle_honeybee_variants(vname, golden):
    g = golden("boundary_r2")
    for name, f in _bee_cases(vname):
        base, out = uv.honeybee_visualize(f.copy(), **BEE_VARIANTS[vname])
        ref = g[f"bee/{vname}/{name}"]
        if np.issubdtype(ref.dtype, np.integer):
            assert np.array_equal(out, ref), (vname, name)
        else:
            np.testing.assert_allclose(out, ref, rtol=1e-6, atol=1e-7, err_msg=f"{vname}/{name}")


def test_oracle_cat_float_frames(golden):
    g = golden("boundary_r2")
    for name, f in frames.float_set(*HW):
        human, cat = M.cat_visualize(f.copy())
        for got, key in ((human, "human"), (cat, "cat")):
            ref = g[f"cat/{name}/{key}"]
            assert got.dtype == ref.dtype
            if np.issubdtype(ref.dtype, np.integer):
                assert np.array_equal(got, ref), (name, key)
            else:
                np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7, err_msg=f"{name}/{key}")


def test_resize_taps_restates_cv2_resize():
    """tables.resize_taps (what avb_img_resample consumes) against cv2.resize on float frames."""
    import cv2
    from animal_vision_b200 import tables as T
    rng = np.random.default_rng(0)

    def run(img, idx, w, axis):
        g = img[:, idx] if axis == 1 else img[idx]
        sub = "hdtc,dt->hdc" if axis == 1 else "dtwc,dt->dwc"
        return np.einsum(sub, g.astype(np.float64), w.astype(np.float64)).astype(np.float32)
    for (H, W, h, w, interp, flag) in [(37, 53, 9, 13, "area", cv2.INTER_AREA), (40, 60, 10, 15, "area", cv2.INTER_AREA),
                                       (54, 76, 5, 8, "area", cv2.INTER_AREA), (9, 13, 37, 53, "linear", cv2.INTER_LINEAR),
                                       (5, 8, 54, 76, "linear", cv2.INTER_LINEAR), (37, 53, 37, 69, "cubic", cv2.INTER_CUBIC),
                                       (25, 35, 37, 53, "linear", cv2.INTER_LINEAR)]:
        img = rng.random((H, W, 3), dtype=np.float32)
        ref = cv2.resize(img, (w, h), interpolation=flag)
        ix, wx = T.resize_taps(W, w, interp)
        iy, wy = T.resize_taps(H, h, interp, vertical=True)
        out = run(run(img, ix, wx, 1), iy, wy, 0)
        assert np.abs(out - ref).max() <= 5e-6, (interp, H, W, h, w)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("vname", list(BEE_VARIANTS))
def test_gpu_honeybee_variants(vname, golden):
    from animal_vision_b200.animals import HoneyBee
    g = golden("boundary_r2")
    bee = HoneyBee(**BEE_VARIANTS[vname])
    for name, f in _bee_cases(vname):
        src = f.copy()
        base, out = bee.visualize(src)
        assert base is src and np.array_equal(src, f)
        # float outputs of the mappers: the opponent channels are DIFFERENCES of catches (G-B, B-U: an ulp of a catch
        # is 10-30 ulp of the radius), val*(1-sat) cancels again and the OETF multiplies what is left by up to 12.92 --
        # the reference's own float32 result is only defined to ~1e-5 absolute there.  The catches themselves are held
        # to 1e-5 relative (tests/test_gpu_honeybee.py); the encoded output to 1e-5 relative + 2e-5 (0.005 uint8 LSB).
        # hsi_downsample adds cv2's IPP resize (sample coordinates ~1e-6 px off OpenCV's C++ arithmetic): 4e-5 (0.01 LSB)
        _close(out, g[f"bee/{vname}/{name}"], f"bee/{vname}/{name}", scale=10.0 if vname.startswith("down") else 5.0)


@pytest.mark.gpu
def test_gpu_honeybee_plane_route_equals_fused_route():
    """The same uint8 frame through the fused K3 kernel and through the float32 plane route (forced)."""
    import torch
    from animal_vision_b200.animals import HoneyBee
    from animal_vision_b200.engine import get_engine
    eng = get_engine()
    x = torch.from_numpy(np.stack([frames.natural(90, 130), frames.checker(90, 130)])).cuda()
    for kw in ({}, dict(mapping_mode="falsecolor", adaptation="gray_world"), dict(mapping_mode="uv_purple_yellow", blur_sigma_px=0.5)):
        bee = HoneyBee(**kw)
        fused = bee.visualize_batch(x)[1]
        planes = torch.empty_like(x)
        bee._run_planes(eng, x, planes, False)
        d = (fused.int() - planes.int()).abs()
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) <= 0.01, kw


@pytest.mark.gpu
def test_gpu_cat_float_frames(golden):
    from animal_vision_b200.animals import Cat
    g = golden("boundary_r2")
    for name, f in frames.float_set(*HW):
        src = f.copy()
        human, cat = Cat().visualize(src)
        assert np.array_equal(src, f)
        # the zoomed frame keeps the caller's value range (0..255 here): an interpolation between unrelated neighbours is
        # only defined to a weight error times the RANGE (the installed cv2 resizes float frames through its IPP path,
        # whose sample coordinates sit ~1e-6 px away from OpenCV's own C++ arithmetic that tables.resize_taps restates)
        _close(human, g[f"cat/{name}/human"], f"cat/{name}/human", scale=max(1.0, float(np.abs(f).max())))
        _close(cat, g[f"cat/{name}/cat"], f"cat/{name}/cat")


@pytest.mark.gpu
def test_gpu_imgops_against_cv2_and_numpy():
    """The K6 operators one by one: resize (area / linear / cubic, crop), blur, stats, percentile."""
    import cv2
    import torch
    from animal_vision_b200.engine import get_engine
    from animal_vision_b200.imgops import get_imgops
    ops = get_imgops(get_engine())
    rng = np.random.default_rng(3)
    img = rng.random((2, 54, 76, 3), dtype=np.float32)
    d = torch.from_numpy(img).cuda()
    for (hw, interp, flag) in [((5, 8), "area", cv2.INTER_AREA), ((14, 19), "area", cv2.INTER_AREA), ((54, 99), "cubic", cv2.INTER_CUBIC),
                               ((108, 152), "linear", cv2.INTER_LINEAR)]:
        out = ops.resize(d, hw, interp).cpu().numpy()
        for k in range(2):
            ref = cv2.resize(img[k], (hw[1], hw[0]), interpolation=flag)
            assert np.abs(out[k] - ref).max() <= 5e-6, (hw, interp)
    crop = (7, 5, 50, 36)
    out = ops.resize(d, (54, 76), "linear", crop=crop).cpu().numpy()
    ref = cv2.resize(np.ascontiguousarray(img[1][5:41, 7:57]), (76, 54), interpolation=cv2.INTER_LINEAR)
    assert np.abs(out[1] - ref).max() <= 5e-6
    for sigma in (0.2, 1.2, 3.0):
        k = int(2 * np.ceil(3 * sigma) + 1)
        out = ops.gaussian_blur(d, sigma).cpu().numpy()
        ref = cv2.GaussianBlur(img[0], (k, k), sigmaX=sigma, sigmaY=sigma, borderType=cv2.BORDER_REFLECT101)
        assert np.abs(out[0] - ref).max() <= 2e-6, sigma
    st = ops.stats(d).cpu().numpy()
    for k in range(2):
        for c in range(3):
            assert st[k, c, 0] == img[k, :, :, c].min() and st[k, c, 1] == img[k, :, :, c].max()
            assert abs(st[k, c, 2] - img[k, :, :, c].mean(dtype=np.float64)) <= 1e-6
    sgn = torch.from_numpy(img - 0.5).cuda()                  # negative values order correctly
    reqs = [(0, 0, 95.0), (1, 2, 50.0), (0, 1, 0.0), (1, 1, 100.0), (0, 2, 99.0), (1, 0, 37.5)]
    got = ops.percentile(sgn, reqs).cpu().numpy()
    for (f, c, q), v in zip(reqs, got):
        assert v == np.float32(np.percentile((img - 0.5)[f, :, :, c], q)), (f, c, q)


@pytest.mark.gpu
def test_gpu_percentile_large_plane_matches_numpy_float32_index():
    """NumPy keeps the virtual index (n-1)*q/100 in float32: at 4K sizes that decides WHICH order statistics are
    blended.  The device restates it (csrc/avb_common.cuh numpy_percentile_index / numpy_lerp): bit-equal."""
    import torch
    from animal_vision_b200.engine import get_engine
    from animal_vision_b200.imgops import get_imgops
    ops = get_imgops(get_engine())
    x = np.random.default_rng(5).random((1, 2160, 3840, 1), dtype=np.float32)
    d = torch.from_numpy(x).cuda()
    reqs = [(0, 0, q) for q in (95.0, 98.0, 99.0, 50.0, 12.3)]
    got = ops.percentile(d, reqs).cpu().numpy()
    for (_, _, q), v in zip(reqs, got):
        assert v == np.percentile(x[0, :, :, 0], q), q
