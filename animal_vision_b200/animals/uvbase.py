"""Shared machinery of the UV species: the steps every one of them takes from uv_helpers.py, on the
device.  A species class keeps the reference's constructor signature and writes its `_render(st)` as NumPy-like
statements over lazy expressions (animal_vision_b200/lazy.py); spatial operators and reductions are the K6 kernels,
every run of element-wise statements is one K7 launch.

Common recipe (e.g. animals/reindeer.py:88-135):
    img01 = to_float01(image); img_lin = srgb_to_linear(img01)                    uv_helpers.py:15-37
    baseline_lin = panorama_warp(img_lin, scale_x)                                 uv_helpers.py:84-99
    baseline_out = from_float01(linear_to_srgb(clip(baseline_lin)), dtype)         uv_helpers.py:26-44
    hsi = classic_rgb_to_hsi_scaled(baseline_lin, wavelengths, scale)              uv_helpers.py:155-183
    band maps = integrate_band(hsi, lambdas, lo, hi) [+ safe_norm]                 uv_helpers.py:47-53, 125-152
The H x W x 81 cube is never built: INTER_AREA down -> analytic spectrum -> band integral -> INTER_LINEAR up is
linear in the (double-linearised, classic_rgb_to_hsi.py:54) small frame, so every band map is a 3-vector applied to
the small frame and up-sampled as ONE plane (672 MB of cube per 1080p frame in the reference).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .. import lazy as L
from .. import tables
from .._abi import AvbError
from ..engine import get_engine
from ..imgops import get_imgops
from .animal import Animal


def band_matrix(lambdas: np.ndarray, bands: Sequence[Tuple[float, float]]) -> np.ndarray:
    """[K,3] float32: band k of the analytic spectrum of a (linearised) pixel c is M[k] . c
    (uv_helpers.py:125-146 weights x classic_rgb_to_hsi.py:60-78 lobes / normaliser, composed in float64)."""
    lam = np.asarray(lambdas, np.float32)
    sens = np.stack([tables.bandpass_weights(lam, float(lo), float(hi)) for lo, hi in bands]).astype(np.float32)
    return tables.uv_collapsed_matrix(lam, sens, None)


def radial_sigmoid(H: int, W: int, softness: float, radius: float) -> np.ndarray:
    """t = 1 / (1 + exp(-softness (r - radius))), r = sqrt(xx^2 + yy^2) on [-1,1]^2 -- the peripheral mask every species
    builds with these very statements (e.g. goldfish.py:166-172); pixel independent, so a host table."""
    yy = (np.linspace(-1.0, 1.0, H, dtype=np.float32))[:, None]
    xx = (np.linspace(-1.0, 1.0, W, dtype=np.float32))[None, :]
    r = np.sqrt(xx * xx + yy * yy)
    return (1.0 / (1.0 + np.exp(-softness * (r - radius)))).astype(np.float32)


def scan_row_gain(H: int, freq: float, soften: float, gain: float) -> np.ndarray:
    """Per-row gain 1 + gain * (rows - 0.5), rows = 0.5 + 0.5 sin(2 pi freq y) blurred with sigma `soften`
    (jumping_spider.py:196-203, mantis_shrimp.py:254-261).  The reference blurs an (H,W) image that is constant along x:
    its horizontal pass is the identity (taps sum to 1), the vertical pass a 1-D REFLECT_101 correlation -- done here."""
    y = np.linspace(0.0, 1.0, H, dtype=np.float32)[:, None]
    rows = (0.5 + 0.5 * np.sin(2.0 * np.pi * freq * y)).astype(np.float32)[:, 0]
    if soften > 0.0:
        taps = tables.uv_blur_taps(float(soften)).astype(np.float64)
        r = taps.size // 2
        idx = np.arange(-r, H + r)
        period = max(1, 2 * H - 2)
        idx = np.abs(((idx % period) + period) % period)
        idx = np.where(idx >= H, period - idx, idx) if H > 1 else np.zeros_like(idx)
        ext = rows.astype(np.float64)[idx]
        rows = np.array([np.dot(taps, ext[i:i + taps.size]) for i in range(H)], np.float32)
    return (1.0 + gain * (rows - 0.5)).astype(np.float32)


class UVStage:
    """One visualize call: batch geometry, the K6 operators and the lazy-expression context."""

    def __init__(self, eng, frames):
        t = eng.torch
        if not (frames.is_cuda and frames.dim() == 4 and frames.shape[3] == 3 and frames.dtype in (t.uint8, t.float32)):
            raise AvbError("UV species: expected a CUDA uint8 / float32 tensor [N,H,W,3]")
        self.eng, self.t, self.ops = eng, t, get_imgops(eng)
        self.n, self.H, self.W = int(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2])
        self.lz = L.Lazy(eng, self.n, self.H, self.W)
        self.img01 = self.ops.to_float01(frames.contiguous())                               # uv_helpers.py:15-23
        self.img_lin = self.lz.eval([L.srgb_to_linear(c) for c in self.lz.channels(self.img01)])   # :33-37
        self.baseline_lin = self.img_lin

    # ---- geometry
    def set_panorama(self, scale_x: float):
        if scale_x and scale_x != 1.0:                                                     # reindeer.py:93-96
            self.baseline_lin = self.ops.panorama(self.img_lin, float(scale_x))
        return self.baseline_lin

    def baseline(self) -> List[L.E]:
        return self.lz.channels(self.baseline_lin)

    # ---- spectral bands
    def bands(self, lambdas, bands: Sequence[Tuple[float, float]], hsi_scale: float):
        """Raw integrate_band maps of classic_rgb_to_hsi[_scaled](baseline_lin) -> float32 tensor [n,H,W,K].
        The spectrum is evaluated per wavelength WITH the reference's clamp at zero (classic_rgb_to_hsi.py:80): the
        bicubic panorama warp overshoots, a negative channel can pull single wavelengths below zero and the band
        integral is then not a 3-vector of the pixel.  That loop runs on the down-sampled frame (1/16 of the pixels at
        the default hsi_scale), three bands per pass of the band-table kernel (avb_uv_catches_f32)."""
        lam = np.asarray(lambdas, np.float32)
        src = self.baseline_lin
        fast = 0.0 < hsi_scale < 1.0                                                        # reindeer.py:102
        if fast:
            hs, ws = tables.scaled_hw(self.H, self.W, hsi_scale)
            src = self.ops.resize(src, (hs, ws), "area")                                    # uv_helpers.py:173
        groups = []
        for k0 in range(0, len(bands), 3):
            grp = list(bands[k0:k0 + 3])
            sens = np.zeros((3, lam.size), np.float32)
            for i, (lo, hi) in enumerate(grp):
                sens[i] = tables.bandpass_weights(lam, float(lo), float(hi))                # uv_helpers.py:125-139
            tab, denom_eps = tables.uv_band_table(lam, sens, None)
            dev = self.eng.cached(("uv_bands", tab.tobytes()), lambda tab=tab: self.eng._dev(tab))
            got = self.ops.uv_catches(src, np.zeros((3, 3), np.float32), dev, denom_eps)    # classic_rgb_to_hsi.py:54-80 + uv_helpers.py:142-146
            groups.append(got[..., :len(grp)])
        out = groups[0].contiguous() if len(groups) == 1 else self.t.cat(groups, dim=3).contiguous()
        if fast:
            out = self.ops.resize(out, (self.H, self.W), "linear")                          # uv_helpers.py:182
        return out

    # ---- reductions (results stay on the device)
    def stats(self, tensor):
        """[n, C*4] per frame: (min, max, mean, 0) of every channel."""
        return self.ops.stats(tensor).view(self.n, -1)

    def safe_norm(self, x: L.E, stats, ch: int) -> L.E:
        """uv_helpers.py:47-53 with the plane's min / max read from `stats` (channel ch)."""
        mn, mx = self.lz.scalar(stats, 4 * ch), self.lz.scalar(stats, 4 * ch + 1)
        rng = mx - mn
        return L.where(rng < 1e-9, 0.0, (x - mn) / rng)

    def normed_bands(self, band_tensor) -> List[L.E]:
        st = self.stats(band_tensor)
        return [self.safe_norm(self.lz.plane(band_tensor, k), st, k) for k in range(band_tensor.shape[3])]

    def percentile(self, tensor, ch: int, q: float) -> L.E:
        return self.lz.scalar(self.ops.percentile_frames(tensor, ch, q), 0)

    # ---- spatial operators on materialised expressions
    def eval(self, exprs: Sequence[L.E]):
        return self.lz.eval(list(exprs))

    def blur(self, tensor, sigma: float):
        return self.ops.gaussian_blur(tensor, float(sigma))                                 # uv_helpers.py:67-73

    def periph_t(self, softness: float, radius: float) -> L.E:
        return self.lz.keyed(("radial_t", float(softness), float(radius)), lambda: radial_sigmoid(self.H, self.W, softness, radius))


def unsharp(st: UVStage, img: Sequence[L.E], sigma: float, amount, *, materialised=None) -> List[L.E]:
    """img + amount * clip(img - blur(img), -1, 1), clipped to [0,1] (the `_unsharp` helper of anchovy.py:122-127 etc.;
    `amount` may be a per-pixel expression: species gate it with a saliency map)."""
    t = materialised if materialised is not None else st.eval(img)
    cur = st.lz.channels(t)
    blurred = st.lz.channels(st.blur(t, sigma))
    return [L.clip(c + amount * L.clip(c - b, -1.0, 1.0), 0.0, 1.0) for c, b in zip(cur, blurred)]


def tone_compress(img: Sequence[L.E], strength: float, knee: float = 0.8) -> List[L.E]:
    """uv_helpers.py:110-121 snow_glare_tone_compress (soft knee above `knee`, linear light)."""
    if strength <= 0.0:
        return list(img)
    out = []
    for c in img:
        x = L.clip(c, 0.0, 1.0)
        tt = (x - knee) / (1.0 - knee)
        out.append(L.where(x <= knee, x, knee + (1.0 - knee) * (tt / (1.0 + strength * tt))))
    return out


def scatter_and_blue_bias(st: UVStage, img: Sequence[L.E], sigma: float, blue_bias: float) -> List[L.E]:
    """uv_helpers.py:101-107 apply_scatter_and_blue_bias: blur when sigma > 0.15, then lift the third channel."""
    cur = list(img)
    if sigma > 0.15:
        cur = st.lz.channels(st.blur(st.eval(cur), sigma))
    cur[2] = L.clip(cur[2] + float(blue_bias), 0.0, 1.0)
    return cur


def periph_mix(st: UVStage, img: Sequence[L.E], sigma: float, softness: float, radius: float) -> List[L.E]:
    """(1 - t) * render + t * gaussian_blur(render, sigma) with the radial sigmoid t (goldfish.py:163-172 and siblings)."""
    t_img = st.eval(img)
    cur, per = st.lz.channels(t_img), st.lz.channels(st.blur(t_img, sigma))
    t = st.periph_t(softness, radius)
    return [(1.0 - t) * c + t * q for c, q in zip(cur, per)]


class UVAnimal(Animal):
    """Base of the panorama / UV species: `visualize` returns a NEW baseline (geometry-warped) and the rendered view.
    Constructor: keyword-only, the reference's parameter names and defaults (class attribute DEFAULTS)."""
    N_OUTPUTS = 2
    DEFAULTS: dict = {}

    def __init__(self, **kw):
        unknown = set(kw) - set(self.DEFAULTS)
        if unknown:
            raise TypeError(f"{type(self).__name__}() got unexpected keyword arguments {sorted(unknown)}")
        for k, v in {**self.DEFAULTS, **kw}.items():
            setattr(self, k, v)
        self.lambdas = self._default_lambdas() if self.lambdas is None else np.asarray(self.lambdas, dtype=np.float32)
        assert self.lambdas.ndim == 1 and self.lambdas.size >= 10, "lambdas must be a 1D vector of wavelengths (nm)."

    def _default_lambdas(self) -> np.ndarray:
        return np.linspace(300.0, 700.0, 81, dtype=np.float32)            # reindeer.py:56 and siblings

    def _render(self, st: UVStage) -> List[L.E]:
        raise NotImplementedError

    def _panorama_scale(self) -> float:
        return float(getattr(self, "panorama_scale", 1.0))

    def _run(self, eng, frames, base_out, out, integer: bool, st: "UVStage" = None):
        st = st if st is not None else UVStage(eng, frames)
        st.set_panorama(self._panorama_scale())
        enc = (lambda e: L.quantize(L.linear_to_srgb(L.clip(e, 0.0, 1.0)))) if integer else (lambda e: L.linear_to_srgb(L.clip(e, 0.0, 1.0)))
        st.lz.run([([enc(c) for c in st.baseline()], base_out)])                            # reindeer.py:98-99
        render = self._render(st)
        st.lz.run([([enc(c) for c in render], out)])                                        # reindeer.py:131-133
        return st

    def visualize_batch(self, frames, out=None):
        """frames: CUDA uint8 (or float32) [N,H,W,3] -> (baseline, view), same dtype and shape.
        out: None, the view tensor, or a (baseline, view) pair of caller-allocated tensors (HostBatchPipeline)."""
        eng = get_engine(frames.device)
        t = eng.torch
        base = None
        if isinstance(out, (tuple, list)):
            base, out = out
        if base is None:
            base = t.empty(tuple(frames.shape), dtype=frames.dtype, device=frames.device)
        if out is None:
            out = t.empty(tuple(frames.shape), dtype=frames.dtype, device=frames.device)
        self._run(eng, frames, base, out, integer=frames.dtype == t.uint8)
        return base, out

    def visualize(self, image: np.ndarray) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        assert isinstance(image, np.ndarray), "Input must be a numpy ndarray."             # reindeer.py:82-83
        assert image.ndim == 3 and image.shape[2] == 3, "Input must be HxWx3 RGB."
        eng = get_engine()
        t = eng.torch
        integer = np.issubdtype(image.dtype, np.integer)
        with t.cuda.device(eng.device):
            if image.dtype == np.uint8:
                dev_in = t.from_numpy(np.ascontiguousarray(image)[None]).pin_memory().to(eng.device, non_blocking=True)
            else:
                dev_in = t.from_numpy(np.ascontiguousarray(image.astype(np.float32))[None]).pin_memory().to(eng.device, non_blocking=True)
            base = t.empty_like(dev_in)
            out = t.empty_like(dev_in)
            self._run(eng, dev_in, base, out, integer)
            res = t.stack([base[0], out[0]]).cpu().numpy()
        return res[0].astype(image.dtype), res[1].astype(image.dtype)
