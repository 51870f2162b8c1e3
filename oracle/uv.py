"""Oracle: the UV path -- analytic RGB->HSI reconstruction, receptor projection, (U,B,G) mappers,
HoneyBee.visualize and the mantis-shrimp band projection.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates reference
ml/classic_rgb_to_hsi/classic_rgb_to_hsi.py:47-82 (analytic branch, run on CPU tensors),
uv_helpers.py, uv_mappers.py, animals/honeybee.py:99-192, animals/mantis_shrimp.py:49-60.
The Mallett-2019 branch (classic_rgb_to_hsi.py:84-115) is NOT restated: it needs colour-science
(unpinned in requirements.txt:27, not installed) and its basis tables.
"""
from __future__ import annotations

import numpy as np
import torch

from . import cvops as V

EPS = 1e-8  # uv_helpers.py:11 / uv_mappers.py:5

# classic_rgb_to_hsi.py:63-64: lobe centres / widths in nm, in INPUT CHANNEL order 2,1,0
LOBE_CENTRES = (610.0, 545.0, 460.0)
LOBE_SIGMAS = (60.0, 60.0, 55.0)


def default_wavelengths() -> np.ndarray:
    return np.linspace(400.0, 700.0, 31, dtype=np.float32)   # honeybee.py:81-85


# ----------------------------------------------------------------------------- dtype / transfer
def to_float01(x: np.ndarray) -> np.ndarray:
    """uv_helpers.py:15-23: uint8 is ALWAYS divided by 255; other dtypes only if max > 1.001."""
    if x.dtype == np.uint8:
        return x.astype(np.float32) / 255.0
    y = x.astype(np.float32)
    if y.max() > 1.001:
        y = np.clip(y / 255.0, 0.0, 1.0)
    return y


def encode_srgb_f32(l: np.ndarray) -> np.ndarray:
    """uv_helpers.py:40-44 (float32-returning variant; clips negatives inside the pow)."""
    s = np.where(l <= 0.0031308, l * 12.92, (1 + 0.055) * np.power(np.clip(l, 0.0, None), 1 / 2.4) - 0.055)
    return s.astype(np.float32)


def decode_srgb_torch(t: torch.Tensor) -> torch.Tensor:
    """classic_rgb_to_hsi.py:16-22 (torch.jit.script in the reference; same eager ops)."""
    return torch.where(t <= 0.04045, t / 12.92, ((t + 0.055) / (1.0 + 0.055)) ** 2.4)


# ----------------------------------------------------------------------------- RGB -> HSI
def lobe_table(wavelengths: np.ndarray):
    """The three Gaussian lobes sampled on `wavelengths` and the scalar normaliser, computed with
    the reference's own torch float32 expressions (classic_rgb_to_hsi.py:60-78).

    Returns (g2, g1, g0, denom): g2 multiplies input channel 2, g1 channel 1, g0 channel 0.
    Note :75 writes the third denominator term as (wl-c)^2/s^2 instead of ((wl-c)/s)^2 -- kept.
    """
    wl = torch.as_tensor(wavelengths.astype(np.float32)).view(-1, 1, 1)
    (c2, c1, c0), (s2, s1, s0) = LOBE_CENTRES, LOBE_SIGMAS
    g2 = torch.exp(-0.5 * ((wl - c2) / s2) ** 2)
    g1 = torch.exp(-0.5 * ((wl - c1) / s1) ** 2)
    g0 = torch.exp(-0.5 * ((wl - c0) / s0) ** 2)
    w = wl.squeeze()
    denom = (torch.exp(-0.5 * ((w - c2) / s2) ** 2) + torch.exp(-0.5 * ((w - c1) / s1) ** 2)
             + torch.exp(-0.5 * ((w - c0) ** 2) / (s0 ** 2))).mean()
    return g2, g1, g0, denom


def analytic_hsi(frame01: np.ndarray, wavelengths: np.ndarray | None = None) -> np.ndarray:
    """classic_rgb_to_hsi.py:47-82 on CPU tensors: H x W x B float32 cube.

    Channel 0 always drives the 460 nm lobe, channel 2 the 610 nm lobe, whatever the caller's
    channel order is (HoneyBee passes RGB, so red drives "blue")."""
    wavelengths = default_wavelengths() if wavelengths is None else wavelengths
    t = decode_srgb_torch(torch.as_tensor(frame01, dtype=torch.float32))
    c0, c1, c2 = t[..., 0], t[..., 1], t[..., 2]
    g2, g1, g0, denom = lobe_table(wavelengths)
    spec = g2 * c2.unsqueeze(0) + g1 * c1.unsqueeze(0) + g0 * c0.unsqueeze(0)
    spec = spec / (denom + 1e-8)
    return spec.clamp_min(0.0).permute(1, 2, 0).contiguous().numpy().astype(np.float32)


# ----------------------------------------------------------------------------- illuminant / receptors
def d65_like(lam: np.ndarray) -> np.ndarray:
    """uv_helpers.py:187-192."""
    x = (lam - 560.0) / 50.0
    base = np.exp(-0.5 * x ** 2) + 0.3 * np.exp(-0.5 * ((lam - 450.0) / 35.0) ** 2)
    base /= base.mean()
    return base.astype(np.float32)


def honeybee_curves(lam: np.ndarray):
    """honeybee.py:179-192 + the in-place sum normalisation of :88-93 (float32)."""
    def bump(peak, sigma):
        return np.exp(-0.5 * ((lam - peak) / sigma) ** 2).astype(np.float32)
    out = []
    for v in (bump(350.0, 25.0), bump(440.0, 30.0), bump(540.0, 35.0)):
        s = v.sum()
        if s > 0:
            v /= s
        out.append(v)
    return out


def bandpass_weights(lam: np.ndarray, lo: float, hi: float) -> np.ndarray:
    """uv_helpers.py:125-139: raised cosine on [lo,hi], sum-normalised; UNIFORM 1/B fallback when no
    sample (or no mass) falls inside the band."""
    wl = lam.astype(np.float32)
    w = np.zeros_like(wl, dtype=np.float32)
    inside = (wl >= lo) & (wl <= hi)
    if not np.any(inside):
        return np.ones_like(wl, dtype=np.float32) / float(wl.size)
    x = (wl[inside] - lo) / (hi - lo)
    w[inside] = 0.5 * (1.0 - np.cos(2.0 * np.pi * x))
    s = float(np.sum(w))
    if s > 1e-12:
        w /= s
    else:
        w = np.ones_like(wl, dtype=np.float32) / float(wl.size)
    return w


def integrate_band(hsi: np.ndarray, lam: np.ndarray, lo: float, hi: float) -> np.ndarray:
    """uv_helpers.py:142-146."""
    return np.tensordot(hsi, bandpass_weights(lam, lo, hi), axes=([2], [0])).astype(np.float32)


def safe_norm(x: np.ndarray) -> np.ndarray:
    """uv_helpers.py:47-53."""
    x = x.astype(np.float32)
    mn, mx = float(np.min(x)), float(np.max(x))
    if mx - mn < 1e-9:
        return np.zeros_like(x, dtype=np.float32)
    return (x - mn) / (mx - mn)


# mantis_shrimp.py:49-60: the ten narrow bands (nm) of the "barcode"
MANTIS_BANDS = ((320, 360), (360, 400), (400, 430), (430, 460), (460, 490),
                (490, 520), (520, 550), (550, 580), (580, 610), (610, 680))


def mantis_band_matrix(lam: np.ndarray) -> np.ndarray:
    """10 x B projection matrix built with bandpass_weights for MANTIS_BANDS."""
    return np.stack([bandpass_weights(lam, float(lo), float(hi)) for lo, hi in MANTIS_BANDS])


def project_cube(hsi: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """H x W x B cube times (N x B)^T -> H x W x N float32, one tensordot per receptor as the
    reference does (honeybee.py:133-135, uv_helpers.py:145)."""
    return np.stack([np.tensordot(hsi, w, axes=([2], [0])) for w in weights], axis=2).astype(np.float32)


# ----------------------------------------------------------------------------- adaptation / blur
def von_kries(U, B, G, mode: str | None, eps: float = EPS):
    """uv_helpers.py:195-206 ("white_patch": global max, "gray_world": global mean)."""
    if mode == "white_patch":
        return U / max(U.max(), eps), B / max(B.max(), eps), G / max(G.max(), eps)
    if mode == "gray_world":
        return U / max(U.mean(), eps), B / max(B.mean(), eps), G / max(G.mean(), eps)
    return U, B, G


def uv_gaussian_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """uv_helpers.py:67-73: explicit ksize 2*ceil(3 sigma)+1, REFLECT101; identity for sigma<=0."""
    if sigma <= 0:
        return img
    k = int(2 * np.ceil(3 * sigma) + 1)
    return V.gaussian_blur(img, sigma, (k, k))


# ----------------------------------------------------------------------------- mappers
def hsv_to_rgb(hsv: np.ndarray) -> np.ndarray:
    """uv_mappers.py:14-26."""
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    i = np.floor(h * 6.0).astype(np.int32)
    f = h * 6.0 - i
    p = v * (1.0 - s)
    q = v * (1.0 - f * s)
    t = v * (1.0 - (1.0 - f) * s)
    k = i % 6
    sel = [k == n for n in range(6)]
    r = np.select(sel, [v, q, p, p, t, v], default=0)
    g = np.select(sel, [t, v, v, q, p, p], default=0)
    b = np.select(sel, [p, p, t, v, v, q], default=0)
    return np.stack([r, g, b], axis=2)


def map_opponent(U, B, G, eps: float = EPS) -> np.ndarray:
    """uv_mappers.py:53-64: two GLOBAL 95th percentiles (np.percentile, linear interpolation)."""
    O1, O2 = G - B, B - U
    L = (U + B + G) / 3.0
    hue = (np.arctan2(O2, O1) + np.pi) / (2 * np.pi)
    radius = np.sqrt(O1 * O1 + O2 * O2)
    sat = radius / (np.percentile(radius, 95.0) + eps)
    val = L / (np.percentile(L, 95.0) + eps)
    return hsv_to_rgb(np.stack([hue, np.clip(sat, 0, 1), np.clip(val, 0, 1)], axis=2)).astype(np.float32)


def map_falsecolor(U, B, G, eps: float = EPS) -> np.ndarray:
    """uv_mappers.py:29-43."""
    def n95(x):
        return x / max(float(np.percentile(x, 95.0)), eps)
    Un, Bn, Gn = n95(U), n95(B), n95(G)
    rgb = np.stack([0.85 * Un + 0.10 * Gn, 0.80 * Gn + 0.20 * Bn, 0.70 * Bn + 0.40 * Un], axis=2)
    return np.clip(rgb, 0.0, 1.0).astype(np.float32)


def map_linear_matrix(U, B, G, M: np.ndarray) -> np.ndarray:
    """uv_mappers.py:45-50."""
    H, W = U.shape
    return (np.stack([U, B, G], axis=2).reshape(-1, 3) @ M.T).reshape(H, W, 3).astype(np.float32)


def _s2l(v):
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / (1 + 0.055)) ** 2.4).astype(np.float32)


def map_uv_purple_yellow_soft(U, *, u_gamma=0.90, accent_gamma=0.85, accent_strength=0.05,
                              eps: float = EPS) -> np.ndarray:
    """uv_mappers.py:90-132."""
    denom = max(float(np.percentile(U, 98.0)), eps)
    u = (U.astype(np.float32) / denom).clip(0.0, 1.0) ** float(u_gamma)
    c0 = _s2l(np.array([176, 124, 232], np.float32) / 255.0)
    c1 = _s2l(np.array([255, 211, 138], np.float32) / 255.0)
    u3 = u[..., None]
    rgb = (1.0 - u3) * c0 + u3 * c1
    if accent_strength > 0:
        rgb = rgb + float(accent_strength) * (u ** float(accent_gamma))[..., None] * (c0 - np.array([0.5, 0.5, 0.5], np.float32))
    Y = (0.2126 * rgb[..., 0] + 0.7152 * rgb[..., 1] + 0.0722 * rgb[..., 2]) + eps
    gain = np.clip((np.clip(0.22 + 0.55 * u, 0.0, 1.0) / Y)[..., None], 0.6, 1.6)
    rgb = rgb * gain
    rgb = rgb / (1.0 + 0.6 * rgb)
    return np.clip(rgb, 0.0, 1.0).astype(np.float32)


def map_falsecolor_uv_mixed(U, B, G, alpha: float = 0.35) -> np.ndarray:
    """uv_mappers.py:135-144."""
    a = float(np.clip(alpha, 0.0, 1.0))
    mixed = (1.0 - a) * map_falsecolor(U, B, G) + a * map_uv_purple_yellow_soft(U)
    p99 = float(np.percentile(mixed, 99.0))
    if p99 > EPS:
        mixed = mixed / max(1.0, p99)
    return np.clip(mixed.astype(np.float32), 0.0, 1.0)


# ----------------------------------------------------------------------------- resampling helpers
def resize_preserve_range(x: np.ndarray, out_hw, interp: int) -> np.ndarray:
    """uv_helpers.py:57-64."""
    import cv2
    was_float = np.issubdtype(x.dtype, np.floating)
    y = cv2.resize(x.astype(np.float32, copy=False), (int(out_hw[1]), int(out_hw[0])), interpolation=interp)
    return y.astype(x.dtype, copy=False) if not was_float else y


def analytic_hsi_scaled(rgb01: np.ndarray, wavelengths: np.ndarray, scale: float) -> np.ndarray:
    """uv_helpers.py:155-183 classic_rgb_to_hsi_scaled: INTER_AREA down -> analytic HSI -> INTER_LINEAR up."""
    import cv2
    assert 0.0 < scale <= 1.0
    H, W = rgb01.shape[:2]
    hs, ws = max(1, int(round(H * scale))), max(1, int(round(W * scale)))
    small = resize_preserve_range(rgb01, (hs, ws), cv2.INTER_AREA)
    cube = analytic_hsi(small, wavelengths.astype(np.float32))
    return resize_preserve_range(cube, (H, W), cv2.INTER_LINEAR)


def panorama_warp(img_lin: np.ndarray, scale_x: float) -> np.ndarray:
    """uv_helpers.py:84-99: bicubic horizontal widen, centre crop back to the original width."""
    import cv2
    if abs(scale_x - 1.0) < 1e-3:
        return img_lin
    H, W = img_lin.shape[:2]
    newW = max(2, int(round(W * scale_x)))
    widened = cv2.resize(img_lin, (newW, H), interpolation=cv2.INTER_CUBIC)
    if newW == W:
        return widened
    start = (newW - W) // 2
    return widened[:, start:start + W, :]


# ----------------------------------------------------------------------------- HoneyBee
def honeybee_receptors(image: np.ndarray, lam: np.ndarray | None = None, *, reflectance=True,
                       hsi_downsample: bool = False, hsi_scale: float = 0.1):
    """honeybee.py:105-135: raw (U, B, G) cone catches through the full 31-band cube."""
    lam = default_wavelengths() if lam is None else lam
    if hsi_downsample and 0.05 <= hsi_scale < 1.0:                       # honeybee.py:109-116
        hsi = analytic_hsi_scaled(to_float01(image), lam, hsi_scale)
    else:
        hsi = analytic_hsi(to_float01(image), lam)
    radiance = hsi * d65_like(lam).astype(hsi.dtype)[None, None, :] if reflectance else hsi
    cu, cb, cg = honeybee_curves(lam)
    return (np.tensordot(radiance, cu, axes=([2], [0])),
            np.tensordot(radiance, cb, axes=([2], [0])),
            np.tensordot(radiance, cg, axes=([2], [0])))


def honeybee_visualize(image: np.ndarray, *, adaptation="white_patch", mapping_mode="opponent",
                       blur_sigma_px=0.2, custom_matrix=None, hsi_downsample=False, hsi_scale=0.1):
    """HoneyBee.visualize with constructor defaults (honeybee.py:47-66, :99-175).
    Returns (image, out): the baseline is the input object itself."""
    assert isinstance(image, np.ndarray) and image.ndim == 3 and image.shape[2] == 3
    U, B, G = honeybee_receptors(image, hsi_downsample=hsi_downsample, hsi_scale=hsi_scale)
    U, B, G = von_kries(U, B, G, adaptation)
    sigma = float(blur_sigma_px or 0.0)
    if sigma > 0:
        U, B, G = (uv_gaussian_blur(c, sigma) for c in (U, B, G))
    if mapping_mode == "opponent":
        rgb = map_opponent(U, B, G)
    elif mapping_mode == "falsecolor":
        rgb = map_falsecolor(U, B, G)
    elif mapping_mode == "custom_matrix":
        rgb = map_linear_matrix(U, B, G, custom_matrix)
    elif mapping_mode == "uv_purple_yellow":
        rgb = map_uv_purple_yellow_soft(U)
    elif mapping_mode == "falsecolor_uv_mixed":
        rgb = map_falsecolor_uv_mixed(U, B, G, alpha=0.45)
    else:
        raise ValueError(f"Unknown mapping_mode: {mapping_mode}")
    srgb = encode_srgb_f32(np.clip(rgb, 0.0, 1.0))
    if np.issubdtype(image.dtype, np.integer):
        return image, (srgb * 255.0 + 0.5).astype(image.dtype)
    return image, srgb.astype(image.dtype)
