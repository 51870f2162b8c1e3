"""CPU oracle of the UV species.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy / OpenCV restatement of animals/{reindeer,goldfish,damselfish,rat_uv,anableps,anchovy,guppy,morpho,heliconius,
pieris,kestrel,jumping_spider,dragonfly,hummingbird,mantis_shrimp}.py of the reference: each function takes the frame
and the species' constructor arguments (a dict, defaults = the reference's) and returns (baseline, view).  Pinned by
tests/golden/uv_species.npz, generated from the unmodified reference by tools/make_golden_uv.py.

Deliberately kept: NumPy's promotion quirks that change results at the 1e-7 level are NOT reproduced on the GPU (it
computes in float32 throughout) but ARE reproduced here, e.g. Goldfish's `np.clip(python_float, 0, 1)` returns a
float64 scalar and silently makes the rest of that species float64.
"""
from __future__ import annotations

import cv2
import numpy as np

from . import uv as U

F32 = np.float32


# ----------------------------------------------------------------------------- shared steps (uv_helpers.py)
def srgb_to_linear(s):                      # uv_helpers.py:33-37
    return np.where(s <= 0.04045, s / 12.92, ((s + 0.055) / (1 + 0.055)) ** 2.4).astype(F32)


def linear_to_srgb(l):                      # uv_helpers.py:40-44
    return np.where(l <= 0.0031308, l * 12.92, (1 + 0.055) * np.power(np.clip(l, 0.0, None), 1 / 2.4) - 0.055).astype(F32)


def from_float01(img01, dtype):             # uv_helpers.py:26-30
    if np.issubdtype(dtype, np.integer):
        return np.clip(img01 * 255.0 + 0.5, 0.0, 255.0).astype(dtype)
    return img01.astype(dtype)


def blur(img, sigma):                       # uv_helpers.py:67-73
    if sigma <= 0:
        return img
    k = int(2 * np.ceil(3 * sigma) + 1)
    return cv2.GaussianBlur(img, (k, k), sigmaX=sigma, sigmaY=sigma, borderType=cv2.BORDER_REFLECT101)


def tone_compress(img, strength, knee=0.8):             # uv_helpers.py:110-121 snow_glare_tone_compress
    if strength <= 0.0:
        return img
    x = np.clip(img, 0.0, 1.0)
    t = (x - knee) / (1.0 - knee)
    return np.where(x <= knee, x, knee + (1.0 - knee) * (t / (1.0 + strength * t))).astype(x.dtype)


def scatter_blue(img, sigma, blue_bias):                # uv_helpers.py:101-107 apply_scatter_and_blue_bias
    out = img.copy()
    if sigma > 0.15:
        out = blur(out, sigma)
    out[..., 2] = np.clip(out[..., 2] + float(blue_bias), 0.0, 1.0)
    return out


def radial_t(H, W, softness, radius):                   # goldfish.py:166-172 and every sibling
    yy = np.linspace(-1.0, 1.0, H, dtype=F32)[:, None]
    xx = np.linspace(-1.0, 1.0, W, dtype=F32)[None, :]
    r = np.sqrt(xx * xx + yy * yy)
    return 1.0 / (1.0 + np.exp(-softness * (r - radius)))


def unsharp(img, sigma, amount):                        # anchovy.py:122-127 `_unsharp`
    if sigma <= 0.0 or amount <= 0.0:
        return img
    return np.clip(img + amount * np.clip(img - blur(img, sigma), -1.0, 1.0), 0.0, 1.0)


def sobel(ch):                                          # mantis_shrimp.py:122-131
    gx = cv2.Sobel(ch, cv2.CV_32F, 1, 0, ksize=3, borderType=cv2.BORDER_REFLECT101)
    gy = cv2.Sobel(ch, cv2.CV_32F, 0, 1, ksize=3, borderType=cv2.BORDER_REFLECT101)
    return gx.astype(F32), gy.astype(F32)


def luma(x):
    return (0.2126 * x[..., 0] + 0.7152 * x[..., 1] + 0.0722 * x[..., 2]).astype(F32)


def sat_apply(lin, scale):                              # heliconius.py:62-64 / guppy.py:111-113
    Y = luma(lin)[..., None]
    return np.clip(Y + (lin - Y) * scale[..., None], 0.0, 1.0).astype(F32)


def hsi_of(baseline_lin, lambdas, hsi_scale, *, nocast=False):
    """reindeer.py:101-109: the scaled route when 0 < hsi_scale < 1, else the full-resolution analytic spectrum."""
    lam32 = np.asarray(lambdas).astype(F32)
    if 0.0 < hsi_scale < 1.0:
        return U.analytic_hsi_scaled(baseline_lin.astype(F32, copy=False), lam32, hsi_scale)
    return U.analytic_hsi(baseline_lin, lam32)


def front(image, panorama_scale):
    """to_float01 -> srgb_to_linear -> panorama_warp -> baseline_out (reindeer.py:88-99)."""
    img01 = U.to_float01(image)
    img_lin = srgb_to_linear(img01)
    base_lin = U.panorama_warp(img_lin, panorama_scale) if (panorama_scale and panorama_scale != 1.0) else img_lin
    base_out = from_float01(linear_to_srgb(np.clip(base_lin, 0.0, 1.0)), image.dtype)
    return img01, base_lin, base_out


def back(render, dtype):
    return from_float01(linear_to_srgb(np.clip(render, 0.0, 1.0)), dtype)


def band(hsi, lam, lohi):
    return U.integrate_band(hsi, np.asarray(lam), float(lohi[0]), float(lohi[1]))


def nband(hsi, lam, lohi):
    return U.safe_norm(band(hsi, lam, lohi))


def lam81():
    return np.linspace(300.0, 700.0, 81, dtype=F32)


def _p(defaults, kw):
    p = dict(defaults)
    unknown = set(kw) - set(p)
    assert not unknown, f"unknown parameters {unknown}"
    p.update(kw)
    return type("P", (), p)


# ----------------------------------------------------------------------------- Reindeer (animals/reindeer.py:70-135)
REINDEER = dict(lambdas=None, hsi_scale=0.25, uv_band=(300.0, 410.0), uv_boost=3.5, snow_glare_compression=0.55, winter_mode=True,
                scatter_sigma=1.2, blue_bias=0.08, panorama_scale=1.3, return_uv_heatmap=True)


def reindeer(image, **kw):
    p = _p(REINDEER, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    _, base, base_out = front(image, p.panorama_scale)
    hsi = hsi_of(base, lam, p.hsi_scale)
    uv_map = nband(hsi, lam, p.uv_band)
    vis = nband(hsi, lam, (420.0, 680.0))
    sal = U.safe_norm(uv_map / (1e-6 + 0.6 * vis))
    r = base.copy()
    r[..., 2] = np.clip(r[..., 2] + p.uv_boost * 0.35 * sal, 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + p.uv_boost * 0.15 * sal, 0.0, 1.0)
    r = tone_compress(r, p.snow_glare_compression)
    if p.winter_mode:
        r = scatter_blue(r, p.scatter_sigma, p.blue_bias)
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Goldfish (animals/goldfish.py:84-180)
GOLDFISH = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
                red_band=(600.0, 680.0), uv_boost=3.0, panorama_scale=1.45, haze_strength=0.12, haze_tint=(0.78, 0.92, 1.0),
                red_kill=0.55, green_lift=0.12, blue_lift=0.06, base_blur_sigma=0.8, periph_blur_sigma=1.8, periph_radius=0.65,
                periph_softness=6.0)


def goldfish(image, **kw):
    p = _p(GOLDFISH, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    _, base, base_out = front(image, p.panorama_scale)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Uv, Bv, Gv, Rv = (nband(hsi, lam, b) for b in (p.uv_band, p.blue_band, p.green_band, p.red_band))
    sal = U.safe_norm(Uv / (1e-6 + 0.45 * Gv + 0.35 * Bv + 0.15 * Rv))
    r = base.copy()
    r[..., 0] = np.clip(r[..., 0] * (1.0 - p.red_kill), 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + p.green_lift, 0.0, 1.0)
    r[..., 2] = np.clip(r[..., 2] + p.blue_lift, 0.0, 1.0)
    if p.haze_strength > 0.0:
        a = np.clip(p.haze_strength, 0.0, 1.0)                  # a float64 SCALAR: the render is float64 from here on
        r = (1.0 - a) * r + a * np.array(p.haze_tint, F32)[None, None, :]
    if p.base_blur_sigma > 0.0:
        r = blur(r, p.base_blur_sigma)
    r[..., 0] = np.clip(r[..., 0] + p.uv_boost * 0.42 * sal, 0.0, 1.0)
    r[..., 2] = np.clip(r[..., 2] + p.uv_boost * 0.35 * sal, 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + p.uv_boost * 0.12 * sal, 0.0, 1.0)
    r[..., 2] = np.clip(r[..., 2] + 0.22 * Bv, 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + 0.30 * Gv, 0.0, 1.0)
    if p.periph_blur_sigma > 0.0:
        per = blur(r, p.periph_blur_sigma)
        t = radial_t(*r.shape[:2], p.periph_softness, p.periph_radius)[..., None]
        r = (1.0 - t) * r + t * per
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Damselfish (animals/damselfish.py:87-181)
DAMSELFISH = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), yellow_band=(560.0, 600.0),
                  red_band=(600.0, 680.0), uv_edge_boost=0.45, uv_gloss_boost=0.30, blue_chroma_gain=0.22, yellow_chroma_gain=0.28,
                  red_kill=0.35, base_blur_sigma=0.35, unsharp_sigma=1.2, panorama_scale=1.25, periph_radius=0.70,
                  periph_softness=7.0, periph_extra_blur=0.8)


def damselfish(image, **kw):
    p = _p(DAMSELFISH, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    _, base, base_out = front(image, p.panorama_scale)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Un = U.safe_norm(nband(hsi, lam, p.uv_band))
    Bn, Yn = nband(hsi, lam, p.blue_band), nband(hsi, lam, p.yellow_band)
    r = base.copy()
    r[..., 0] = np.clip(r[..., 0] * (1.0 - p.red_kill), 0.0, 1.0)
    if p.base_blur_sigma > 0.0:
        r = blur(r, p.base_blur_sigma)
    if p.unsharp_sigma > 0.0 and p.uv_edge_boost > 0.0:
        high = np.clip(r - blur(r, p.unsharp_sigma), -1.0, 1.0)
        r = np.clip(r + (1.0 + p.uv_edge_boost * Un[..., None]) * high, 0.0, 1.0)
    if p.uv_gloss_boost > 0.0:
        lift = p.uv_gloss_boost * Un
        r[..., 2] = np.clip(r[..., 2] + 0.60 * lift, 0.0, 1.0)
        r[..., 1] = np.clip(r[..., 1] + 0.30 * lift, 0.0, 1.0)
        r[..., 0] = np.clip(r[..., 0] + 0.15 * lift, 0.0, 1.0)
    r[..., 2] = np.clip(r[..., 2] + p.blue_chroma_gain * Bn, 0.0, 1.0)
    yb = p.yellow_chroma_gain * Yn
    r[..., 1] = np.clip(r[..., 1] + 0.65 * yb, 0.0, 1.0)
    r[..., 0] = np.clip(r[..., 0] + 0.35 * yb, 0.0, 1.0)
    if p.periph_extra_blur > 0.0:
        per = blur(r, p.periph_extra_blur)
        t = radial_t(*r.shape[:2], p.periph_softness, p.periph_radius)[..., None]
        r = (1.0 - t) * r + t * per
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- RatUV (animals/rat_uv.py:131-214)
RAT_UV = dict(lambdas=None, hsi_scale=0.55, panorama_scale=1.45, uv_boost_alpha=0.55, day_blur_sigma=0.8, night_blur_sigma=1.25,
              blue_bias_day=0.03, blue_bias_night=0.05, tone_knee=0.82, tone_strength=0.65, ground_vignette_day=0.10,
              ground_vignette_night=0.14)


def rat_uv(image, mode="auto", **kw):
    p = _p(RAT_UV, kw)
    if p.lambdas is None:
        lam = np.linspace(320.0, 700.0, 129, dtype=np.float64)                 # rat_uv.py:48
    else:
        wl = np.asarray(p.lambdas, np.float64).ravel()
        lam = np.linspace(float(wl[0]), float(wl[-1]), wl.size, dtype=np.float64)
    img01, base, base_out = front(image, p.panorama_scale)
    hsi = hsi_of(base, lam, p.hsi_scale)                                        # :113-127 (wavelengths reach torch as float32 either way)
    Uv = nband(hsi, lam, (330.0, 400.0))
    Bv, Gv = band(hsi, lam, (400.0, 500.0)), band(hsi, lam, (500.0, 600.0))
    n95 = lambda x: x / max(1e-8, float(np.percentile(x, 95.0)))               # noqa: E731  (:171-172)
    Un, Bn, Gn = n95(Uv), n95(Bv), n95(Gv)
    false = np.stack([np.clip(0.85 * Un + 0.10 * Gn, 0.0, 1.0), np.clip(0.80 * Gn + 0.20 * Bn, 0.0, 1.0),
                      np.clip(0.70 * Bn + 0.40 * Un, 0.0, 1.0)], axis=2).astype(F32)
    a = float(np.clip(p.uv_boost_alpha, 0.0, 1.0))
    r = np.clip((1.0 - a) * base + a * false, 0.0, 1.0)
    if mode == "auto":                                                          # :99-104
        mode = "night" if float(np.median(luma(img01))) < 0.12 else "day"
    night = mode == "night"
    r = scatter_blue(r, p.night_blur_sigma if night else p.day_blur_sigma, p.blue_bias_night if night else p.blue_bias_day)
    if not night:
        r = tone_compress(r, p.tone_strength, p.tone_knee)
    else:
        Y = 0.2126 * r[..., 0] + 0.7152 * r[..., 1] + 0.0722 * r[..., 2]
        r = np.clip(r * ((Y + 0.18) / (Y + 1e-6))[..., None], 0.0, 1.0)
    yy = np.linspace(0.0, 1.0, r.shape[0], dtype=F32)[:, None]                 # :106-111 ground-focus vignette
    gain = 1.0 - (p.ground_vignette_night if night else p.ground_vignette_day) * (1.0 - np.clip(1.0 - yy, 0.0, 1.0))
    r = np.clip(r * gain[..., None], 0.0, 1.0)
    return base_out, back(r, image.dtype)
