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


DEBUG = {}          # last call's orientation inputs (gx, gy): tests mask pixels whose gradient is rounding noise


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


def lam81_f64():
    return np.linspace(300, 700, 81)                     # pieris.py:52, heliconius.py:54, morpho.py:54: float64 (cast where used)


# ----------------------------------------------------------------------------- Pieris (animals/pieris.py:69-124)
PIERIS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
              panorama_scale=1.05, guide_sigma=1.2, guide_gain=0.75, foliage_opponent_gain=0.25, petal_warmth=0.08,
              clarity_unsharp_sigma=0.8, clarity_amount=0.22, center_bias=0.12, bias_radius=0.8, bias_softness=7.0)


def pieris(image, **kw):
    p = _p(PIERIS, kw)
    lam = lam81_f64() if p.lambdas is None else np.asarray(p.lambdas, F32)
    _, base, base_out = front(image, p.panorama_scale if p.panorama_scale != 1.0 else 0)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Uv = U.safe_norm(nband(hsi, lam, p.uv_band))
    Bv, Gv = nband(hsi, lam, p.blue_band), nband(hsi, lam, p.green_band)
    r = base.copy()
    Us = blur(Uv, p.guide_sigma)
    Us = np.clip(Us / (np.percentile(Us, 95.0) + 1e-8), 0.0, 1.0)
    r = np.clip(r + (p.guide_gain * Us)[..., None] * np.array([0.35, 0.35 + p.petal_warmth, 0.25], F32), 0.0, 1.0)
    foliage = np.clip(Gv - 0.5 * (Uv + Bv), 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + p.foliage_opponent_gain * foliage, 0.0, 1.0)
    if p.clarity_unsharp_sigma > 0.0 and p.clarity_amount > 0.0:
        r = np.clip(r + p.clarity_amount * (r - blur(r, p.clarity_unsharp_sigma)), 0.0, 1.0)
    t = radial_t(*r.shape[:2], p.bias_softness, p.bias_radius)
    r = np.clip(r * (1.0 + p.center_bias * (1.0 - t))[..., None], 0.0, 1.0)
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Heliconius (animals/heliconius.py:66-135)
HELICONIUS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), red_band=(600.0, 680.0), green_band=(500.0, 570.0),
                  panorama_scale=1.05, conj_sigma_small=0.8, conj_sigma_large=2.2, conj_gain=1.0, sat_boost=0.45, red_gain=0.4,
                  bg_desat=0.2, bg_cool=0.04, base_soft_sigma=0.3, unsharp_sigma=1.0, unsharp_amount=0.25)


def heliconius(image, **kw):
    p = _p(HELICONIUS, kw)
    lam = lam81_f64() if p.lambdas is None else np.asarray(p.lambdas, F32)
    _, base, base_out = front(image, p.panorama_scale if p.panorama_scale != 1.0 else 0)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Uv = U.safe_norm(nband(hsi, lam, p.uv_band))
    Rb = nband(hsi, lam, p.red_band)
    uv_dog = np.clip(blur(Uv, p.conj_sigma_small) - blur(Uv, p.conj_sigma_large), 0.0, 1.0)
    r_dog = np.clip(blur(Rb, p.conj_sigma_small) - blur(Rb, p.conj_sigma_large), 0.0, 1.0)
    conj = uv_dog * r_dog
    conj = np.clip(conj / (np.percentile(conj, 95.0) + 1e-8), 0.0, 1.0)
    r = base.copy()
    if p.base_soft_sigma > 0.0:
        r = blur(r, p.base_soft_sigma)
    bg = 1.0 - conj
    r[..., 2] = np.clip(r[..., 2] + p.bg_cool * bg, 0.0, 1.0)
    r = sat_apply(r, (1.0 - p.bg_desat * bg).astype(F32))
    if p.unsharp_sigma > 0.0 and p.unsharp_amount > 0.0:
        r = np.clip(r + (p.unsharp_amount * conj[..., None]) * (r - blur(r, p.unsharp_sigma)), 0.0, 1.0)
    r[..., 0] = np.clip(r[..., 0] + p.red_gain * conj, 0.0, 1.0)
    r = sat_apply(r, (1.0 + p.sat_boost * conj).astype(F32))
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Morpho (animals/morpho.py:95-154)
MORPHO = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), green_band=(500.0, 570.0),
              panorama_scale=1.05, sheen_strength=0.55, hue_shift_strength=0.45, gloss_sigma=1.0, mosaic_downscale=0.35,
              center_clarity=0.25, vignette_softness=7.0, vignette_radius=0.82)


def morpho(image, **kw):
    p = _p(MORPHO, kw)
    lam = lam81_f64() if p.lambdas is None else np.asarray(p.lambdas, F32)
    mosaic = float(np.clip(p.mosaic_downscale, 0.15, 1.0))
    _, base, base_out = front(image, p.panorama_scale if p.panorama_scale != 1.0 else 0)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Uv = U.safe_norm(nband(hsi, lam, p.uv_band))
    Bv = nband(hsi, lam, p.blue_band)
    r = base.copy()
    gx, gy = sobel(Bv.astype(F32))
    DEBUG["grad"] = (gx, gy)
    ori = np.arctan2(gy, gx).astype(F32)
    align = 0.5 * (1.0 + np.cos(2.0 * ori))
    gloss = blur(Uv, p.gloss_sigma)
    gloss = np.clip(gloss / (np.percentile(gloss, 95.0) + 1e-8), 0.0, 1.0)
    cyan, deep = p.hue_shift_strength * align, p.hue_shift_strength * (1.0 - align)
    r[..., 2] = np.clip(r[..., 2] + 0.40 * deep + 0.25 * cyan, 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + 0.35 * cyan, 0.0, 1.0)
    r = np.clip(r + p.sheen_strength * gloss[..., None] * np.array([0.10, 0.25, 0.45], F32), 0.0, 1.0)
    if mosaic < 0.999:                                                          # :85-93
        H, W = r.shape[:2]
        small = cv2.resize(r, (max(1, int(round(W * mosaic))), max(1, int(round(H * mosaic)))), interpolation=cv2.INTER_AREA)
        r = cv2.resize(small, (W, H), interpolation=cv2.INTER_NEAREST)
    t = radial_t(*r.shape[:2], p.vignette_softness, p.vignette_radius)
    sharp = r + 0.22 * (r - blur(r, 1.0))
    r = np.clip((1.0 - t[..., None]) * sharp + t[..., None] * r, 0.0, 1.0)
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Guppy (animals/guppy.py:132-235)
GUPPY = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
             red_band=(600.0, 680.0), panorama_scale=1.22, red_kill=0.28, haze_strength=0.06, haze_tint=(0.92, 0.98, 1.0),
             warm_tint=(1.03, 1.01, 0.99), base_soft_sigma=0.35, unsharp_sigma=0.9, unsharp_amount=0.28, dog_small_sigma=0.8,
             dog_large_sigma=2.4, dog_gain=0.85, uv_chroma_boost=0.4, uv_blue_gain=0.55, uv_green_gain=0.35, uv_red_gain=0.12,
             background_desat=0.18, vignette_strength=0.12, vignette_radius=0.78, vignette_softness=7.0)


def guppy(image, **kw):
    p = _p(GUPPY, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    _, base, base_out = front(image, p.panorama_scale)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Un = U.safe_norm(nband(hsi, lam, p.uv_band))
    Bn, Gn = nband(hsi, lam, p.blue_band), nband(hsi, lam, p.green_band)
    r = base.copy()
    r[..., 0] = np.clip(r[..., 0] * (1.0 - p.red_kill), 0.0, 1.0)
    if p.haze_strength > 0.0:
        a = float(np.clip(p.haze_strength, 0.0, 1.0))
        r = (1.0 - a) * r + a * np.array(p.haze_tint, F32)[None, None, :]
    r = np.clip(r * np.array(p.warm_tint, F32)[None, None, :], 0.0, 1.0)
    if p.base_soft_sigma > 0.0:
        r = blur(r, p.base_soft_sigma)
    dog = np.clip(blur(Un, p.dog_small_sigma) - blur(Un, p.dog_large_sigma), 0.0, 1.0)
    spot = np.clip(dog / (np.percentile(dog, 95.0) + 1e-8), 0.0, 1.0)
    if p.unsharp_sigma > 0.0 and p.unsharp_amount > 0.0:
        high = np.clip(r - blur(r, p.unsharp_sigma), -1.0, 1.0)
        r = np.clip(r + (p.unsharp_amount * spot[..., None]) * high, 0.0, 1.0)
    lift = p.uv_chroma_boost * spot
    r[..., 2] = np.clip(r[..., 2] + p.uv_blue_gain * lift * Bn, 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + p.uv_green_gain * lift * Gn, 0.0, 1.0)
    r[..., 0] = np.clip(r[..., 0] + p.uv_red_gain * lift * Un, 0.0, 1.0)
    mc = np.mean(np.abs(r - luma(r)[..., None]), axis=2)                       # :106-109 `_saturation`
    sat = (mc / (np.percentile(mc, 95.0) + 1e-8)).astype(F32)
    r = sat_apply(r, (1.0 - p.background_desat * (1.0 - Un) * (1.0 - sat)).astype(F32))
    if p.vignette_strength > 0.0:
        t = radial_t(*r.shape[:2], p.vignette_softness, p.vignette_radius)
        r = np.clip(r * (1.0 - p.vignette_strength * t)[..., None], 0.0, 1.0)
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Anchovy (animals/anchovy.py:130-253)
ANCHOVY = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), green_band=(500.0, 570.0),
               red_band=(600.0, 680.0), panorama_scale=1.2, red_kill=0.25, base_soft_sigma=0.3, unsharp_sigma=1.0, unsharp_amount=0.35,
               haze_strength=0.04, haze_tint=(0.9, 0.97, 1.0), evec_angle_deg=0.0, pol_strength=0.55, pol_gamma=1.2,
               orientation_mix=0.35, uv_gloss_gain=0.28, blue_chroma_gain=0.18, green_chroma_gain=0.1, periph_blur_sigma=0.6,
               periph_radius=0.78, periph_softness=7.0)


def anchovy(image, **kw):
    p = _p(ANCHOVY, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    evec = np.deg2rad(float(p.evec_angle_deg))
    mix = float(np.clip(p.orientation_mix, 0.0, 1.0))
    _, base, base_out = front(image, p.panorama_scale)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Un = U.safe_norm(nband(hsi, lam, p.uv_band))
    Bn, Gn = nband(hsi, lam, p.blue_band), nband(hsi, lam, p.green_band)
    gx, gy = sobel(Un.astype(F32))
    DEBUG["grad"] = (gx, gy)
    theta = np.arctan2(gy, gx).astype(F32)
    align = (1.0 - mix) * float(np.cos(2.0 * evec)) + mix * np.cos(2.0 * theta)
    align01 = np.clip(0.5 * (align + 1.0), 0.0, 1.0) ** float(p.pol_gamma)
    mag = np.sqrt(gx * gx + gy * gy)
    mag = np.clip(mag / (np.percentile(mag, 95.0) + 1e-8), 0.0, 1.0)
    pol_gain = 1.0 + p.pol_strength * (align01 * Un * mag)
    r = base.copy()
    r[..., 0] = np.clip(r[..., 0] * (1.0 - p.red_kill), 0.0, 1.0)
    if p.haze_strength > 0.0:
        a = float(np.clip(p.haze_strength, 0.0, 1.0))
        r = (1.0 - a) * r + a * np.array(p.haze_tint, F32)[None, None, :]
    if p.base_soft_sigma > 0.0:
        r = blur(r, p.base_soft_sigma)
    if p.unsharp_sigma > 0.0 and p.unsharp_amount > 0.0:
        high = np.clip(r - blur(r, p.unsharp_sigma), -1.0, 1.0)
        r = np.clip(r + (p.unsharp_amount * pol_gain[..., None]) * high, 0.0, 1.0)
    gloss = p.uv_gloss_gain * (align01 * Un)
    r[..., 2] = np.clip(r[..., 2] + 0.70 * gloss, 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + 0.30 * gloss, 0.0, 1.0)
    r[..., 2] = np.clip(r[..., 2] + p.blue_chroma_gain * (Bn * Un), 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + p.green_chroma_gain * (Gn * Un), 0.0, 1.0)
    if p.periph_blur_sigma > 0.0:
        per = blur(r, p.periph_blur_sigma)
        t = radial_t(*r.shape[:2], p.periph_softness, p.periph_radius)[..., None]
        r = (1.0 - t) * r + t * per
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Kestrel (animals/kestrel.py:113-234)
KESTREL = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), green_band=(500.0, 570.0),
               red_band=(600.0, 680.0), panorama_scale=1.1, sky_cool_tint=(0.95, 0.98, 1.03), sky_haze=0.1,
               ground_warm_tint=(1.02, 1.01, 0.99), ground_contrast=0.08, uv_overlay_strength=0.55, uv_magenta=(0.6, 0.12, 0.7),
               ridge_sigma=3, ridge_gain=1.0, unsharp_sigma=1.0, unsharp_amount=0.3, periph_blur_sigma=0.7, periph_radius=0.82,
               periph_softness=7.0)


def ridge_measure(u, sigma):                                                    # kestrel.py:113-136 structure-tensor coherence
    gx, gy = sobel(u)
    gxx, gyy, gxy = blur(gx * gx, sigma), blur(gy * gy, sigma), blur(gx * gy, sigma)
    trace, diff = gxx + gyy, gxx - gyy
    root = np.sqrt(np.maximum((0.5 * diff) ** 2 + gxy * gxy, 0.0)).astype(F32)
    lam1, lam2 = 0.5 * trace + root, 0.5 * trace - root
    coh = (lam1 - lam2) / (lam1 + lam2 + 1e-8)
    energy = np.clip(trace, 0.0, None)
    energy /= np.percentile(energy, 95.0) + 1e-8
    return np.clip(coh * energy, 0.0, 1.0).astype(F32)


def kestrel(image, **kw):
    p = _p(KESTREL, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    cool, warm, magenta = (np.array(v, F32) for v in (p.sky_cool_tint, p.ground_warm_tint, p.uv_magenta))
    _, base, base_out = front(image, p.panorama_scale if p.panorama_scale != 1.0 else 0)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Uv = U.safe_norm(nband(hsi, lam, p.uv_band))
    Bv, Gv = nband(hsi, lam, p.blue_band), nband(hsi, lam, p.green_band)
    H, W = base.shape[:2]
    prior = np.linspace(1.0, 0.0, H, dtype=F32)[:, None]
    sky = blur(0.6 * prior + 0.4 * np.clip(Bv - 0.6 * Gv, 0.0, 1.0), 3.0)
    sky = np.clip(sky / (np.percentile(sky, 98.0) + 1e-8), 0.0, 1.0)
    sky_w = 1.0 / (1.0 + np.exp(-6.0 * (sky - 0.45)))
    ground_w = 1.0 - sky_w
    s3, g3 = sky_w[..., None], ground_w[..., None]
    trail = np.clip(float(p.ridge_gain) * ridge_measure(Uv, float(p.ridge_sigma)) * ground_w, 0.0, 1.0)
    r = base.copy()
    if p.sky_haze > 0.0:
        a = float(np.clip(p.sky_haze, 0.0, 1.0))
        r = s3 * ((1.0 - a) * np.clip(r * cool[None, None, :], 0.0, 1.0) + a * np.array([0.90, 0.97, 1.00], F32)) + g3 * r
    else:
        r = s3 * np.clip(r * cool[None, None, :], 0.0, 1.0) + g3 * r
    gp = np.clip(r.copy() * warm[None, None, :], 0.0, 1.0)
    if p.ground_contrast > 0.0:
        gp = np.clip(gp + p.ground_contrast * (gp - blur(gp, 1.2)), 0.0, 1.0)
    r = s3 * r + g3 * gp
    U95 = np.clip(Uv / (np.percentile(Uv, 95.0) + 1e-8), 0.0, 1.0)
    uv_rgb = U95[..., None] * magenta[None, None, :]
    r = np.clip((1.0 - p.uv_overlay_strength * g3) * r + (p.uv_overlay_strength * g3) * uv_rgb, 0.0, 1.0)
    if p.unsharp_sigma > 0.0 and p.unsharp_amount > 0.0:
        high = np.clip(r - blur(r, p.unsharp_sigma), -1.0, 1.0)
        r = np.clip(r + (p.unsharp_amount * trail[..., None]) * high, 0.0, 1.0)
    if p.periph_blur_sigma > 0.0:
        per = blur(r, p.periph_blur_sigma)
        t = radial_t(H, W, p.periph_softness, p.periph_radius)
        r = (1.0 - t[..., None]) * r + t[..., None] * per
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- JumpingSpider (animals/jumping_spider.py:135-236)
JUMPING_SPIDER = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), green_band=(500.0, 570.0), red_band=(600.0, 680.0),
                      blue_band=(430.0, 500.0), panorama_scale=1.02, dog_small_sigma=0.9, dog_large_sigma=2.2, uv_patch_gain=0.95,
                      opponent_gain=0.3, red_kill=0.25, base_soft_sigma=0.25, clarity_sigma=0.9, clarity_amount=0.24, fovea_radius=0.38,
                      fovea_softness=10.0, periph_blur_sigma=2.2, periph_vignette_strength=0.22, scan_row_freq=22.0, scan_row_gain=0.08,
                      scan_soften=0.9, spots=((0.5, 0.52), (0.57, 0.48)), spot_sigma=0.08, spot_gain=0.2)


def attention_spots(H, W, spots, spot_sigma):                                   # jumping_spider.py:122-133
    yy = np.linspace(0.0, 1.0, H, dtype=F32)[:, None]
    xx = np.linspace(0.0, 1.0, W, dtype=F32)[None, :]
    mask = np.zeros((H, W), F32)
    s2 = max(spot_sigma, 1e-4) ** 2
    for yc, xc in spots:
        mask += np.exp(-((yy - yc) ** 2 + (xx - xc) ** 2) / (2.0 * s2))
    m95 = max(1e-8, float(np.percentile(mask, 95.0)))
    return np.clip(mask / m95, 0.0, 1.0).astype(F32)


def scan_rows(H, W, freq, soften):                                              # jumping_spider.py:196-202, mantis_shrimp.py:254-260
    y = np.linspace(0.0, 1.0, H, dtype=F32)[:, None]
    rows = 0.5 + 0.5 * np.sin(2.0 * np.pi * freq * y)
    rows = rows * np.ones((1, W), dtype=F32)
    return blur(rows, soften) if soften > 0.0 else rows


def jumping_spider(image, **kw):
    p = _p(JUMPING_SPIDER, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    spots = tuple((float(y), float(x)) for (y, x) in p.spots)
    _, base, base_out = front(image, p.panorama_scale if p.panorama_scale != 1.0 else 0)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Uv = U.safe_norm(nband(hsi, lam, p.uv_band))
    Gv, Bv = nband(hsi, lam, p.green_band), nband(hsi, lam, p.blue_band)
    r = base.copy()
    r[..., 0] = np.clip(r[..., 0] * (1.0 - p.red_kill), 0.0, 1.0)
    if p.base_soft_sigma > 0.0:
        r = blur(r, p.base_soft_sigma)
    dog = np.clip(blur(Uv, p.dog_small_sigma) - blur(Uv, p.dog_large_sigma), 0.0, 1.0)
    patch = np.clip(dog / (np.percentile(dog, 95.0) + 1e-8), 0.0, 1.0)
    opp = Gv - Uv
    opp = np.clip(opp / (np.percentile(np.abs(opp), 95.0) + 1e-8), -1.0, 1.0)
    gb, ub = np.clip(opp, 0.0, 1.0) * p.opponent_gain, np.clip(-opp, 0.0, 1.0) * p.opponent_gain
    r[..., 1] = np.clip(r[..., 1] + 0.40 * gb, 0.0, 1.0)
    r[..., 2] = np.clip(r[..., 2] + 0.30 * ub * Bv, 0.0, 1.0)
    r[..., 0] = np.clip(r[..., 0] + 0.12 * ub * Uv, 0.0, 1.0)
    if p.clarity_sigma > 0.0 and p.clarity_amount > 0.0:
        high = np.clip(r - blur(r, p.clarity_sigma), -1.0, 1.0)
        r = np.clip(r + (p.clarity_amount * p.uv_patch_gain * patch[..., None]) * high, 0.0, 1.0)
    H, W = r.shape[:2]
    if p.scan_row_gain != 0.0:
        rows = scan_rows(H, W, p.scan_row_freq, p.scan_soften)
        r = np.clip(r * (1.0 + p.scan_row_gain * (rows - 0.5))[..., None], 0.0, 1.0)
    sm = attention_spots(H, W, spots, p.spot_sigma)
    if p.spot_gain > 0.0:
        r = np.clip(r + p.spot_gain * sm[..., None], 0.0, 1.0)
        sharp = unsharp(r, 0.8, 0.25)
        r = np.clip((1.0 - 0.6 * sm[..., None]) * r + (0.6 * sm[..., None]) * sharp, 0.0, 1.0)
    if p.periph_blur_sigma > 0.0 or p.periph_vignette_strength > 0.0:
        edge = radial_t(H, W, p.fovea_softness, p.fovea_radius)
        if p.periph_blur_sigma > 0.0:
            per = blur(r, p.periph_blur_sigma)
            r = (1.0 - edge[..., None]) * r + edge[..., None] * per
        if p.periph_vignette_strength > 0.0:
            r = np.clip(r * (1.0 - p.periph_vignette_strength * edge)[..., None], 0.0, 1.0)
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Dragonfly (animals/dragonfly.py:146-251)
DRAGONFLY = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), green_band=(500.0, 570.0),
                 red_band=(600.0, 680.0), panorama_scale=1.15, sky_prior_strength=0.6, sky_blue_weight=0.4, sky_sigmoid_mid=0.46,
                 sky_sigmoid_steepness=6.0, sky_pol_strength=0.65, sky_pol_gamma=1.3, water_pol_strength=0.55, water_pol_gamma=1.2,
                 sky_evec_base_deg=90.0, sky_evec_sweep_deg=-45.0, red_kill=0.22, sky_uv_blue_gain=(0.25, 0.2),
                 water_uv_blue_gain=(0.3, 0.24), ventral_green_gain=0.12, base_soft_sigma=0.3, unsharp_sigma=1.0, unsharp_amount=0.3,
                 highlight_knee=0.85, highlight_strength=0.35, periph_blur_sigma=0.7, periph_radius=0.8, periph_softness=7.0)


def soft_knee(lin, knee, amount):                                               # dragonfly.py:133-144
    if amount <= 0.0:
        return lin
    x = np.clip(lin, 0.0, 1.0)
    t = (x - knee) / (1.0 - knee + 1e-8)
    return np.where(x <= knee, x, knee + (1.0 - knee) * (t / (1.0 + amount * t))).astype(x.dtype)


def dragonfly(image, **kw):
    p = _p(DRAGONFLY, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    evec_base, evec_sweep = np.deg2rad(float(p.sky_evec_base_deg)), np.deg2rad(float(p.sky_evec_sweep_deg))   # numpy float64 SCALARS
    sky_gain_ub, water_gain_ub = tuple(map(float, p.sky_uv_blue_gain)), tuple(map(float, p.water_uv_blue_gain))
    _, base, base_out = front(image, p.panorama_scale if p.panorama_scale != 1.0 else 0)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Uv = U.safe_norm(nband(hsi, lam, p.uv_band))
    Bv, Gv = nband(hsi, lam, p.blue_band), nband(hsi, lam, p.green_band)
    H, W = base.shape[:2]
    prior = np.linspace(1.0, 0.0, H, dtype=F32)[:, None]
    score = blur(p.sky_prior_strength * prior + p.sky_blue_weight * np.clip(Bv - 0.6 * Gv, 0.0, 1.0), 2.5)
    score = score / (np.percentile(score, 98.0) + 1e-8)
    sky_w = 1.0 / (1.0 + np.exp(-p.sky_sigmoid_steepness * (score - p.sky_sigmoid_mid)))
    ground_w = 1.0 - sky_w
    gx, gy = sobel((0.6 * Bv + 0.4 * Uv).astype(F32))
    DEBUG["grad"] = (gx, gy)
    theta = np.arctan2(gy, gx).astype(F32)
    y_norm = np.linspace(0.0, 1.0, H, dtype=F32)[:, None]
    sky_evec = evec_base + evec_sweep * y_norm                                  # float64 (H,1): everything downstream is float64
    c2, s2 = np.cos(2.0 * theta), np.sin(2.0 * theta)
    align_sky01 = np.clip(0.5 * ((c2 * np.cos(2.0 * sky_evec) + s2 * np.sin(2.0 * sky_evec)) + 1.0), 0.0, 1.0) ** p.sky_pol_gamma
    align_water01 = np.clip(0.5 * ((c2 * 1.0 + s2 * 0.0) + 1.0), 0.0, 1.0) ** p.water_pol_gamma
    r = base.copy()
    r[..., 0] = np.clip(r[..., 0] * (1.0 - p.red_kill), 0.0, 1.0)
    if p.base_soft_sigma > 0.0:
        r = blur(r, p.base_soft_sigma)
    sky_gain = (1.0 + p.sky_pol_strength * (align_sky01 * sky_w))[..., None]
    r = np.clip(r * (0.95 + 0.05 * sky_w[..., None]), 0.0, 1.0)
    r[..., 2] = np.clip(r[..., 2] + sky_gain_ub[1] * (Bv * sky_w * align_sky01), 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + 0.10 * (Uv * sky_w * align_sky01), 0.0, 1.0)
    r = np.clip(r * sky_gain, 0.0, 1.0)
    water_gain = (1.0 + p.water_pol_strength * (align_water01 * ground_w))[..., None]
    r[..., 2] = np.clip(r[..., 2] + water_gain_ub[1] * (Bv * ground_w * align_water01), 0.0, 1.0)
    r[..., 2] = np.clip(r[..., 2] + water_gain_ub[0] * (Uv * ground_w * align_water01), 0.0, 1.0)
    r[..., 1] = np.clip(r[..., 1] + p.ventral_green_gain * (Gv * ground_w), 0.0, 1.0)
    r = np.clip(r * water_gain, 0.0, 1.0)
    if p.unsharp_sigma > 0.0 and p.unsharp_amount > 0.0:
        r = np.clip(r + p.unsharp_amount * np.clip(r - blur(r, p.unsharp_sigma), -1.0, 1.0), 0.0, 1.0)
    r = soft_knee(r, p.highlight_knee, p.highlight_strength)
    if p.periph_blur_sigma > 0.0:
        per = blur(r, p.periph_blur_sigma)
        t = radial_t(H, W, p.periph_softness, p.periph_radius)
        r = (1.0 - t[..., None]) * r + t[..., None] * per
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Hummingbird (animals/hummingbird.py:128-227)
HUMMINGBIRD = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
                   red_band=(600.0, 680.0), panorama_scale=1.05, red_kill=0.1, base_soft_sigma=0.25, unsharp_sigma=0.9,
                   unsharp_amount=0.24, combo_opacity=0.55, combo_saturation=0.45, combo_sheen=0.28, tgt_uvb_srgb=(120, 150, 255),
                   tgt_uvg_srgb=(110, 255, 170), tgt_uvr_srgb=(255, 110, 210), guide_sigma=1.0, guide_gain=0.25, periph_blur_sigma=0.6,
                   periph_radius=0.82, periph_softness=7.0)


def _s2l_target(rgb):                                                           # hummingbird.py:96-99
    v = np.array(rgb, F32) / 255.0
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / (1 + 0.055)) ** 2.4).astype(F32)


def hummingbird(image, **kw):
    p = _p(HUMMINGBIRD, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    _, base, base_out = front(image, p.panorama_scale if p.panorama_scale != 1.0 else 0)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Uv = U.safe_norm(nband(hsi, lam, p.uv_band))
    Bv, Gv, Rv = (nband(hsi, lam, b) for b in (p.blue_band, p.green_band, p.red_band))

    def bandpass(m):
        d = np.clip(blur(m, 0.8) - blur(m, 2.0), 0.0, 1.0)
        return np.clip(d / (np.percentile(d, 95.0) + 1e-8), 0.0, 1.0).astype(F32)
    cb, cg, cr = (bandpass(U.safe_norm(Uv * x)) for x in (Bv, Gv, Rv))
    r = base.copy()
    r[..., 0] = np.clip(r[..., 0] * (1.0 - p.red_kill), 0.0, 1.0)
    if p.base_soft_sigma > 0.0:
        r = blur(r, p.base_soft_sigma)
    cmax = np.maximum.reduce([cb, cg, cr])
    if p.unsharp_sigma > 0.0 and p.unsharp_amount > 0.0:
        high = np.clip(r - blur(r, p.unsharp_sigma), -1.0, 1.0)
        r = np.clip(r + (p.unsharp_amount * cmax[..., None]) * high, 0.0, 1.0)
    if p.combo_sheen > 0.0:
        r = np.clip(r + p.combo_sheen * (0.55 * cb + 0.65 * cg + 0.75 * cr)[..., None], 0.0, 1.0)
    ws = cb + cg + cr + 1e-8
    tint = ((cb / ws)[..., None] * _s2l_target(p.tgt_uvb_srgb)[None, None, :] + (cg / ws)[..., None] * _s2l_target(p.tgt_uvg_srgb)[None, None, :]
            + (cr / ws)[..., None] * _s2l_target(p.tgt_uvr_srgb)[None, None, :]).astype(F32)
    s = 1.0 + p.combo_saturation
    if s != 1.0:
        Y = (0.2126 * tint[..., 0] + 0.7152 * tint[..., 1] + 0.0722 * tint[..., 2])[..., None]
        tint = np.clip(Y + (tint - Y) * s, 0.0, 1.0).astype(F32)
    r = np.clip((1.0 - p.combo_opacity) * r + p.combo_opacity * tint, 0.0, 1.0)
    if p.guide_gain > 0.0:
        Us = blur(Uv, p.guide_sigma)
        Us = np.clip(Us / (np.percentile(Us, 95.0) + 1e-8), 0.0, 1.0)
        r = np.clip(r + p.guide_gain * Us[..., None] * np.array([0.20, 0.25, 0.10], F32), 0.0, 1.0)
    if p.periph_blur_sigma > 0.0:
        H, W = r.shape[:2]
        per = blur(r, p.periph_blur_sigma)
        t = radial_t(H, W, p.periph_softness, p.periph_radius)
        r = (1.0 - t[..., None]) * r + t[..., None] * per
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- Anableps (animals/anableps.py:124-255)
ANABLEPS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
                red_band=(600.0, 680.0), panorama_scale=1.2, horizon_y=0.44, seam_softness_px=8.0, ripple_amp_px=6.0, ripple_waves=2.5,
                refract_push_px=3.0, air_warmth=(1.06, 1.03, 0.99), air_clarity_unsharp=0.35, air_unsharp_sigma=1.0, red_kill=0.55,
                blue_lift=0.08, green_lift=0.12, haze_strength=0.1, haze_tint=(0.8, 0.92, 1.0), base_blur_sigma_water=0.7, uv_boost=3.4,
                uv_R_gain=0.36, uv_G_gain=0.18, uv_B_gain=0.42, periph_blur_sigma=1.2, periph_radius=0.7, periph_softness=6.0)


def anableps_geometry(H, W, p):
    """anableps.py:171-186, :224-231: horizon with ripple, air / water weights, refraction map (all pixel independent)."""
    y0 = int(np.clip(p.horizon_y * H, 0, H - 1))
    if p.ripple_amp_px > 0.0:
        x = np.linspace(0, 2.0 * np.pi * p.ripple_waves, W, dtype=F32)
        ripple = (p.ripple_amp_px * np.sin(x)).astype(F32)
    else:
        ripple = np.zeros((W,), F32)
    yy = np.arange(H, dtype=F32)[:, None]
    seam = max(1.0, float(p.seam_softness_px))
    horizon = y0 + ripple[None, :]
    air_w = 1.0 / (1.0 + np.exp(+(yy - horizon) / seam))
    maps = None
    if p.refract_push_px > 0.0:
        yi = np.repeat(np.arange(H, dtype=F32)[:, None], W, axis=1)
        xi = np.repeat(np.arange(W, dtype=F32)[None, :], H, axis=0)
        push = p.refract_push_px * np.exp(-np.maximum(yi - horizon, 0.0) / (2.5 * p.seam_softness_px))
        maps = (xi.astype(F32), np.clip(yi + push, 0, H - 1).astype(F32))
    return air_w, maps


def anableps(image, **kw):
    p = _p(ANABLEPS, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    _, base, base_out = front(image, p.panorama_scale)
    hsi = hsi_of(base, lam, p.hsi_scale)
    Un = U.safe_norm(nband(hsi, lam, p.uv_band))
    Bv, Gv = nband(hsi, lam, p.blue_band), nband(hsi, lam, p.green_band)
    H, W = base.shape[:2]
    air_w, maps = anableps_geometry(H, W, p)
    air = unsharp(np.clip(base.copy() * np.array(p.air_warmth, F32)[None, None, :], 0.0, 1.0), p.air_unsharp_sigma, p.air_clarity_unsharp)
    wt = base.copy()
    wt[..., 0] = np.clip(wt[..., 0] * (1.0 - p.red_kill), 0.0, 1.0)
    wt[..., 1] = np.clip(wt[..., 1] + p.green_lift, 0.0, 1.0)
    wt[..., 2] = np.clip(wt[..., 2] + p.blue_lift, 0.0, 1.0)
    if p.haze_strength > 0.0:
        a = np.clip(p.haze_strength, 0.0, 1.0)                  # float64 scalar (as in Goldfish)
        wt = (1.0 - a) * wt + a * np.array(p.haze_tint, F32)[None, None, :]
    if p.base_blur_sigma_water > 0.0:
        wt = blur(wt, p.base_blur_sigma_water)
    wt[..., 0] = np.clip(wt[..., 0] + p.uv_boost * p.uv_R_gain * Un, 0.0, 1.0)
    wt[..., 1] = np.clip(wt[..., 1] + p.uv_boost * p.uv_G_gain * Un, 0.0, 1.0)
    wt[..., 2] = np.clip(wt[..., 2] + p.uv_boost * p.uv_B_gain * Un, 0.0, 1.0)
    wt[..., 2] = np.clip(wt[..., 2] + 0.20 * Bv, 0.0, 1.0)
    wt[..., 1] = np.clip(wt[..., 1] + 0.26 * Gv, 0.0, 1.0)
    if maps is not None:
        wt = cv2.remap(wt.astype(F32), maps[0], maps[1], interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT101)
    r = air * air_w[..., None] + wt * (1.0 - air_w)[..., None]
    if p.periph_blur_sigma > 0.0:
        per = blur(r, p.periph_blur_sigma)
        t = radial_t(H, W, p.periph_softness, p.periph_radius)[..., None]
        r = (1.0 - t) * r + t * per
    return base_out, back(r, image.dtype)


# ----------------------------------------------------------------------------- MantisShrimp (animals/mantis_shrimp.py:143-279)
MANTIS_SHRIMP = dict(lambdas=None, hsi_scale=0.25, panorama_scale=1.12,
                     bands=((320.0, 360.0), (360.0, 400.0), (400.0, 430.0), (430.0, 460.0), (460.0, 490.0), (490.0, 520.0), (520.0, 550.0),
                            (550.0, 580.0), (580.0, 610.0), (610.0, 680.0)),
                     red_kill=0.18, haze_strength=0.03, haze_tint=(0.92, 0.98, 1.0), pre_soft_sigma=0.25, unsharp_sigma=1.0,
                     unsharp_amount=0.32, evec_angle_deg=30.0, pol_linear_strength=0.55, pol_linear_gamma=1.2, pol_circular_strength=0.35,
                     orientation_mix=0.5, barcode_saturation=0.4, barcode_opacity=0.55, winner_take_most=0.35, scan_row_freq=26.0,
                     scan_row_gain=0.08, scan_soften=0.8, periph_blur_sigma=0.7, periph_radius=0.8, periph_softness=7.0)


def barcode_lut(N):                                                             # mantis_shrimp.py:172-196: N hues, s = 0.95, v = 1
    h = (np.arange(N, dtype=F32) / max(N, 1)).astype(F32)
    s, v = np.full_like(h, 0.95, F32), np.ones_like(h, F32)
    i = np.floor(h * 6.0).astype(np.int32)
    f = h * 6.0 - i
    pp, q, t = v * (1.0 - s), v * (1.0 - f * s), v * (1.0 - (1.0 - f) * s)
    i = i % 6
    conds = [i == 0, i == 1, i == 2, i == 3, i == 4, i == 5]
    return np.stack([np.select(conds, [v, q, pp, pp, t, v], default=v), np.select(conds, [t, v, v, q, pp, pp], default=v),
                     np.select(conds, [pp, pp, t, v, v, q], default=v)], axis=-1).astype(F32)


def mantis_shrimp(image, **kw):
    p = _p(MANTIS_SHRIMP, kw)
    lam = lam81() if p.lambdas is None else np.asarray(p.lambdas, F32)
    evec = np.deg2rad(float(p.evec_angle_deg))
    _, base, base_out = front(image, p.panorama_scale if p.panorama_scale != 1.0 else 0)
    hsi = hsi_of(base, lam, p.hsi_scale)
    H, W = base.shape[:2]
    S = np.stack([nband(hsi, lam, (float(a), float(b))) for a, b in p.bands], axis=2).astype(F32)
    lut = barcode_lut(S.shape[2])
    Sn = np.clip(S / (np.percentile(S, 95.0) + 1e-8), 0.0, 1.0)
    mi = np.argmax(Sn, axis=2)
    w = Sn / (np.sum(Sn, axis=2, keepdims=True) + 1e-8)
    bar = (1.0 - p.winner_take_most) * (w @ lut) + p.winner_take_most * lut[mi]
    Yb = (0.2126 * bar[..., 0] + 0.7152 * bar[..., 1] + 0.0722 * bar[..., 2])[..., None]
    bar = np.clip(Yb + (bar - Yb) * (1.0 + p.barcode_saturation), 0.0, 1.0)
    r = base.copy()
    r[..., 0] = np.clip(r[..., 0] * (1.0 - p.red_kill), 0.0, 1.0)
    if p.haze_strength > 0.0:
        a = float(np.clip(p.haze_strength, 0.0, 1.0))
        r = (1.0 - a) * r + a * np.array(p.haze_tint, F32)[None, None, :]
    if p.pre_soft_sigma > 0.0:
        r = blur(r, p.pre_soft_sigma)
    gx, gy = sobel(np.mean(Sn, axis=2).astype(F32))
    DEBUG["grad"] = (gx, gy)
    theta = np.arctan2(gy, gx).astype(F32)
    mix = p.orientation_mix
    c_mix = (1.0 - mix) * float(np.cos(2.0 * evec)) + mix * np.cos(2.0 * theta)
    s_mix = (1.0 - mix) * float(np.sin(2.0 * evec)) + mix * np.sin(2.0 * theta)
    pol = 1.0 + p.pol_linear_strength * (np.clip(0.5 * (c_mix + 1.0), 0.0, 1.0) ** p.pol_linear_gamma) \
        + p.pol_circular_strength * np.clip(0.5 * (s_mix + 1.0), 0.0, 1.0)
    if p.unsharp_sigma > 0.0 and p.unsharp_amount > 0.0:
        high = np.clip(r - blur(r, p.unsharp_sigma), -1.0, 1.0)
        r = np.clip(r + (p.unsharp_amount * pol[..., None]) * high, 0.0, 1.0)
    r = np.clip((1.0 - p.barcode_opacity) * r + p.barcode_opacity * bar, 0.0, 1.0)
    if p.scan_row_gain != 0.0:
        rows = scan_rows(H, W, p.scan_row_freq, p.scan_soften)
        r = np.clip(r * (1.0 + p.scan_row_gain * (rows - 0.5))[..., None], 0.0, 1.0)
    if p.periph_blur_sigma > 0.0:
        per = blur(r, p.periph_blur_sigma)
        t = radial_t(H, W, p.periph_softness, p.periph_radius)[..., None]
        r = (1.0 - t) * r + t * per
    return base_out, back(r, image.dtype)
