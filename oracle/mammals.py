"""Oracle: the 20 non-UV mammal species (dichromat recipe) and Cat.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates reference animals/dog.py:14-61 and its 18
sibling files, animals/animal_utils.py:121-259, animals/cat.py:73-114 (the runnable `Tina-animals`
side of the unresolved merge conflict) and animals/cat_widevision_utils.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import colorimetry as C
from . import cvops as V

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


# ----------------------------------------------------------------------------- species table
@dataclass(frozen=True)
class Recipe:
    """Step 4 matrix parameters and the step-5 filter of one species."""
    alpha: float
    s_scale: float
    kind: str                       # "gauss" | "streak" | "scone"
    sigma: float = 0.0              # gauss
    streak: Tuple[float, float, float, float] = (0.5, 0.8, 2.2, 6.0)  # y_center, s_streak, s_far, falloff
    chroma: float = 0.0             # chroma compression applied after the streak blur (0 = none)
    scone: Tuple[float, float, float, float] = (1.0, 0.6, 1.0, 0.0)   # s_top, s_bottom, power, extra_boost


# file:line of the two parameter lines in each reference species file
RECIPES = {
    "dog":      Recipe(0.58, 0.65, "gauss", sigma=3.5),                    # dog.py:46,51
    "bear":     Recipe(0.60, 0.95, "gauss", sigma=1.6),                    # bear.py:29,34
    "lion":     Recipe(0.60, 0.95, "gauss", sigma=1.2),                    # lion.py
    "tiger":    Recipe(0.60, 0.95, "gauss", sigma=1.2),                    # tiger.py
    "elephant": Recipe(0.60, 0.95, "gauss", sigma=1.8),                    # elephant.py
    "fox":      Recipe(0.65, 0.98, "gauss", sigma=1.3),                    # fox.py
    "wolf":     Recipe(0.65, 0.95, "gauss", sigma=1.4),                    # wolf.py
    "raccoon":  Recipe(0.60, 0.98, "gauss", sigma=2.0),                    # raccoon.py
    "squirrel": Recipe(0.55, 1.05, "gauss", sigma=0.7),                    # squirrel.py
    "rat":      Recipe(0.05, 0.86, "scone", scone=(1.3, 0.5, 1.4, 0.25)),  # rat.py:29,34
    "cow":      Recipe(0.84, 1.07, "streak", streak=(0.5, 0.9, 2.3, 6.5)),  # cow.py:29,34
    "deer":     Recipe(0.60, 0.95, "streak", streak=(0.5, 0.8, 2.6, 8.0)),
    "goat":     Recipe(0.75, 1.06, "streak", streak=(0.5, 0.8, 2.4, 8.0)),
    "horse":    Recipe(0.30, 1.02, "streak", streak=(0.5, 0.8, 2.2, 6.0)),
    "kangaroo": Recipe(0.60, 0.98, "streak", streak=(0.55, 0.8, 2.3, 8.0)),
    "sheep":    Recipe(0.74, 1.06, "streak", streak=(0.48, 0.8, 2.2, 6.0)),
    "panda":    Recipe(0.58, 0.74, "streak", streak=(0.52, 1.0, 2.1, 4.5), chroma=0.06),  # panda.py:29-37
    "rabbit":   Recipe(0.20, 1.01, "streak", streak=(0.52, 0.9, 2.5, 5.0), chroma=0.06),
    # pig.py:30-38: blur mutates its float32 argument in place (result used although the return
    # value is dropped); apply_chroma_compression's result IS dropped -> no chroma step.
    "pig":      Recipe(0.89, 1.32, "streak", streak=(0.5, 1.2, 2.5, 3.0)),
}


# ----------------------------------------------------------------------------- step-5 filters
def acuity_blur(lin: np.ndarray, sigma: float) -> np.ndarray:
    """animal_utils.py:121-145: cv2.GaussianBlur(img,(0,0),sigma,sigma); ksize = cvRound(8s+1)|1."""
    return V.gaussian_blur(lin, sigma)


def streak_sigmas(H: int, y_center: float, s_streak: float, s_far: float, falloff: float):
    """animal_utils.py:155-162, float32 arithmetic exactly as NumPy promotes it. Returns
    (sigmaX[H], sigmaY[H]) as Python floats (the reference passes float(sigma[y,0]) to OpenCV)."""
    yy = np.linspace(0, 1, H, dtype=np.float32)[:, None]
    d = np.abs(yy - y_center)
    smap = s_streak + (s_far - s_streak) * (1.0 - np.exp(-falloff * d ** 2))
    sx = np.maximum(0.4, 0.5 * smap)
    return [float(v) for v in sx[:, 0]], [float(v) for v in smap[:, 0]]


def streak_blur(lin: np.ndarray, y_center=0.5, s_streak=0.8, s_far=2.2, falloff=6.0) -> np.ndarray:
    """animal_utils.py:147-172, done the reference's way: 2*H OpenCV calls on W x 3 single-channel
    row matrices.  Does not mutate its argument (callers that rely on the in-place side effect,
    pig.py:35, simply use the return value here)."""
    assert cv2 is not None
    H = lin.shape[0]
    sx, sy = streak_sigmas(H, y_center, s_streak, s_far, falloff)
    src = lin.astype(np.float32, copy=True)
    tmp = np.empty_like(src)
    for y in range(H):
        tmp[y] = cv2.GaussianBlur(src[y], (0, 0), sigmaX=sx[y], sigmaY=0.0)
    for y in range(H):
        src[y] = cv2.GaussianBlur(tmp[y], (0, 0), sigmaX=1e-16, sigmaY=sy[y])
    return src.astype(lin.dtype, copy=False)


def streak_blur_np(lin: np.ndarray, y_center=0.5, s_streak=0.8, s_far=2.2, falloff=6.0) -> np.ndarray:
    """What those 2*H calls actually compute, without OpenCV.

    Row y, seen by OpenCV as a W-row x 3-column single-channel image:
      pass 1: sigmaY=0 -> copies sigmaX, so the SAME taps g1 (ksize from sigmaX(y)) run along the
              3 colour channels (REFLECT_101 over a width of 3) and along image x;
      pass 2: sigmaX=1e-16 -> 1 tap across channels; taps g2 (ksize from sigmaY(y)) along image x.
    No vertical mixing ever happens.
    """
    H, W = lin.shape[:2]
    sx, sy = streak_sigmas(H, y_center, s_streak, s_far, falloff)
    out = np.empty((H, W, 3), np.float32)
    src = lin.astype(np.float32, copy=False)
    for y in range(H):
        g1 = V.gaussian_taps(V.gaussian_ksize(sx[y]), sx[y])
        g2 = V.gaussian_taps(V.gaussian_ksize(sy[y]), sy[y])
        r1, r2 = len(g1) // 2, len(g2) // 2
        row = src[y]                                            # (W,3)
        cidx = V.reflect101(np.arange(-r1, 3 + r1), 3)
        a = np.zeros_like(row)
        for t, w in enumerate(g1):                              # across channels
            a += w * row[:, cidx[t:t + 3]]
        xidx = V.reflect101(np.arange(-r1, W + r1), W)
        b = np.zeros_like(row)
        for t, w in enumerate(g1):                              # along x, sigmaX
            b += w * a[xidx[t:t + W]]
        xidx = V.reflect101(np.arange(-r2, W + r2), W)
        c = np.zeros_like(row)
        for t, w in enumerate(g2):                              # along x again, sigmaY
            c += w * b[xidx[t:t + W]]
        out[y] = c
    return out


def chroma_compression(lin: np.ndarray, strength: float) -> np.ndarray:
    """animal_utils.py:174-181."""
    gray = lin.mean(axis=2, keepdims=True)
    return gray + (lin - gray) * (1 - strength)


def scone_row_gain(H: int, s_top=1.0, s_bottom=0.6, power=1.0, extra_boost=0.0) -> np.ndarray:
    """Row weights of animal_utils.py:236-247 (band=None)."""
    w = np.linspace(s_top, s_bottom, H, dtype=np.float32)
    if power != 1.0:
        t = (w - s_bottom) / max(1e-8, (s_top - s_bottom))
        t = np.clip(t, 0.0, 1.0) ** power
        w = s_bottom + (s_top - s_bottom) * t
    if extra_boost != 0.0:
        w = 1.0 + extra_boost * (w - 1.0)
    return w


def scone_vertical_gain(lin: np.ndarray, s_top, s_bottom, power, extra_boost) -> np.ndarray:
    """animal_utils.py:206-259 with clamp=True, band=None: channel 2 times a per-row weight, clipped."""
    out = lin.astype(np.float32, copy=True)
    w = scone_row_gain(out.shape[0], s_top, s_bottom, power, extra_boost)
    out[..., 2] = np.clip(out[..., 2] * w[:, None], 0.0, 1.0)
    return out


# ----------------------------------------------------------------------------- the recipe
def dichromat_linear(image: np.ndarray, recipe: Recipe) -> np.ndarray:
    """Steps 2-5 of dog.py:35-51: linear-light result before the encode tail (may be negative)."""
    lin = C.decode_srgb(C.normalize_frame(image))
    lin = C.apply_matrix(lin, C.dichromat_matrix(recipe.alpha, recipe.s_scale))
    if recipe.kind == "gauss":
        lin = acuity_blur(lin, recipe.sigma)
    elif recipe.kind == "streak":
        lin = streak_blur(lin, *recipe.streak)
        if recipe.chroma:
            lin = chroma_compression(lin, recipe.chroma)
    elif recipe.kind == "scone":
        lin = scone_vertical_gain(lin, *recipe.scone)
    else:  # pragma: no cover
        raise ValueError(recipe.kind)
    return lin


def mammal_visualize(image: np.ndarray, species: str):
    """`X.visualize(image)` for the 19 mammals that share dog.py's recipe. Baseline is the input
    object itself (dog.py:61)."""
    assert C.is_frame(image)
    lin = dichromat_linear(image, RECIPES[species])
    return image, C.encode_tail(lin, image.dtype)


# ----------------------------------------------------------------------------- Cat
CAT_CAMERA_HFOV = 100.0      # cat.py:17-21
CAT_HALF_FOV = 105.0
CAT_OVERLAP = 40.0
CAT_TO_HUMAN = 1.30


def cat_zoom_scale(camera_hfov=CAT_CAMERA_HFOV, half_fov=CAT_HALF_FOV, ratio=CAT_TO_HUMAN) -> float:
    """cat_widevision_utils.py:31-44."""
    eff = min(float(camera_hfov), 2.0 * float(half_fov))
    ratio = max(1.01, float(ratio))
    cam = math.tan(math.radians(camera_hfov) * 0.5)
    hum = math.tan(math.radians(eff / ratio) * 0.5)
    return float(cam / max(hum, 1e-6))


def center_zoom_box(W: int, H: int, scale: float):
    """Crop rectangle of cat_widevision_utils.py:19-25 -> (x0, y0, cw, ch)."""
    cw = max(1, int(round(W / scale)))
    ch = max(1, int(round(H / scale)))
    return (W - cw) // 2, (H - ch) // 2, cw, ch


def center_zoom(image: np.ndarray, scale: float) -> np.ndarray:
    """cat_widevision_utils.py:11-29 (u8 frames: cv2.resize INTER_LINEAR, 11-bit fixed point)."""
    if scale <= 1.0:
        return image
    H, W = image.shape[:2]
    x0, y0, cw, ch = center_zoom_box(W, H, scale)
    crop = image[y0:y0 + ch, x0:x0 + cw]
    if image.dtype == np.uint8:
        return V.resize_linear_u8(np.ascontiguousarray(crop), W, H)
    assert cv2 is not None
    return cv2.resize(crop, (W, H), interpolation=cv2.INTER_LINEAR)


def cat_warp_tables(W: int, fov_in=CAT_CAMERA_HFOV, half_fov=CAT_HALF_FOV, overlap=CAT_OVERLAP):
    """Per-column quantities of cat_widevision_utils.py:61-96 (every map is constant down a column).

    Returns float32 vectors of length W: xL, xR (source x of the left/right eye views) and the
    blend weights wL, wR (cos^2 window times validity).  dtype promotion follows NumPy 2: `u` is
    float32 and stays float32 against Python/NumPy float64 *scalars*.
    """
    phi = np.deg2rad(half_fov)
    psi = np.deg2rad(fov_in * 0.5)
    O = np.deg2rad(overlap)
    alpha = max(0.0, phi - 0.5 * O)
    u = np.linspace(-1.0, 1.0, W, dtype=np.float32)
    theta = u * phi
    gL, gR = theta - alpha, theta + alpha
    xL = ((gL / psi) * (W * 0.5) + (W * 0.5)).astype(np.float32)
    xR = ((gR / psi) * (W * 0.5) + (W * 0.5)).astype(np.float32)
    vL = (np.abs(gL) <= psi).astype(np.float32)
    vR = (np.abs(gR) <= psi).astype(np.float32)
    win = (np.cos(0.5 * np.pi * (theta / phi)) ** 2).astype(np.float32)
    return xL, xR, win * vL, win * vR


def cat_binocular_warp(srgb01: np.ndarray) -> np.ndarray:
    """cat_widevision_utils.py:46-99 with the arguments cat.py:84-92 passes (out_size = input size,
    BORDER_CONSTANT 0)."""
    H, W = srgb01.shape[:2]
    xL, xR, wL, wR = cat_warp_tables(W)
    if cv2 is not None:
        ymap = np.repeat(np.linspace(0, H - 1, H, dtype=np.float32)[:, None], W, axis=1)
        left = cv2.remap(srgb01, np.repeat(xL[None], H, 0), ymap, interpolation=cv2.INTER_LINEAR,
                         borderMode=0, borderValue=0.0)
        right = cv2.remap(srgb01, np.repeat(xR[None], H, 0), ymap, interpolation=cv2.INTER_LINEAR,
                          borderMode=0, borderValue=0.0)
    else:  # pragma: no cover
        left, right = V.remap_rows_linear_np(srgb01, xL), V.remap_rows_linear_np(srgb01, xR)
    wsum = (wL + wR + 1e-8)[None, :, None]
    out = (left * wL[None, :, None] + right * wR[None, :, None]) / wsum
    return np.clip(out, 0.0, 1.0).astype(np.float32)


def cat_linear(image: np.ndarray, fov_warp: bool = True) -> np.ndarray:
    """cat.py:83-102: [warped] frame -> linear -> LMS (f32) -> L/M merge -> RGB (float64 from here,
    because LMS_TO_RGB is a float64 matrix) -> 9x9 acuity blur in CV_64F.  fov_warp=False is the
    class switch ENABLE_FOV_WARP = False (cat.py:21, :84)."""
    H, W = image.shape[:2]
    s01 = C.normalize_frame(image).astype(np.float32)
    if fov_warp:
        s01 = cat_binocular_warp(s01)
    lms = C.decode_srgb(s01).reshape(-1, 3) @ C.RGB_TO_LMS.T
    lm = 0.5 * lms[:, 0] + (1.0 - 0.5) * lms[:, 1]
    merged = np.stack([lm, lm, lms[:, 2]], axis=1)
    rgb = (merged @ C.LMS_TO_RGB.T).reshape(H, W, 3)
    return acuity_blur(rgb, 1.0)


def cat_visualize(image: np.ndarray, fov_warp: bool = True):
    """Cat.visualize (cat.py:23-114). Returns (human_zoomed, cat_view) -- NOT the input object."""
    assert isinstance(image, np.ndarray) and image.ndim == 3 and image.shape[2] == 3
    dt = image.dtype
    human = center_zoom(image, cat_zoom_scale())
    cat_srgb = np.clip(C.encode_srgb(np.clip(cat_linear(image, fov_warp), 0.0, 1.0)), 0.0, 1.0)
    if np.issubdtype(dt, np.integer):
        if not np.issubdtype(human.dtype, np.integer):  # pragma: no cover
            human = (np.clip(human, 0, 1) * 255.0 + 0.5).astype(dt)
        return human, (cat_srgb * 255.0 + 0.5).astype(dt)
    return human.astype(dt), cat_srgb.astype(dt)
