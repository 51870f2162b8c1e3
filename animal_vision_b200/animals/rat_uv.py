"""RatUV -- drop-in for reference animals/rat_uv.py (constructor :57-97, visualize :131-214)."""
from typing import Optional, Tuple

import numpy as np

from .. import lazy as L
from ..engine import get_engine
from .uvbase import UVAnimal, UVStage, scatter_and_blue_bias, tone_compress


class RatUV(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.55, panorama_scale=1.45, uv_boost_alpha=0.55, day_blur_sigma=0.8, night_blur_sigma=1.25,
                    blue_bias_day=0.03, blue_bias_night=0.05, tone_knee=0.82, tone_strength=0.65, ground_vignette_day=0.1,
                    ground_vignette_night=0.14)
    UV_BAND, B_BAND, G_BAND = (330.0, 400.0), (400.0, 500.0), (500.0, 600.0)             # :51-53

    def __init__(self, **kw):
        lam = kw.get("lambdas")
        super().__init__(**{**kw, "lambdas": None})
        if lam is not None:                                                                # :77-80: snapped to a uniform grid
            wl = np.asarray(lam, dtype=np.float64).ravel()
            self.lambdas = np.linspace(float(wl[0]), float(wl[-1]), wl.size, dtype=np.float64).astype(np.float32)
        self.uv_boost_alpha = float(np.clip(self.uv_boost_alpha, 0.0, 1.0))
        self._mode = "auto"

    def _default_lambdas(self):
        return np.linspace(320.0, 700.0, 129, dtype=np.float64).astype(np.float32)        # :48 (cast to float32 where it is used)

    def _modes(self, st: UVStage, mode: str):
        """:99-104 `_choose_mode` per frame: night when the median Rec.709 luma of the sRGB frame is below 0.12."""
        if mode != "auto":
            return [mode] * st.n
        y = st.eval([L.luma(st.lz.channels(st.img01))])
        med = st.ops.percentile_frames(y, 0, 50.0).view(-1).cpu().numpy()                  # the one host decision of this species
        return ["night" if float(m) < 0.12 else "day" for m in med]

    def _run(self, eng, frames, base_out, out, integer):
        st0 = UVStage(eng, frames)
        modes = self._modes(st0, self._mode)
        if len(set(modes)) > 1:                       # mixed batch: the blur radius differs, run the two groups separately
            t = eng.torch
            for m in ("day", "night"):
                idx = t.tensor([i for i, mm in enumerate(modes) if mm == m], device=frames.device)
                sub_b, sub_o = t.empty_like(frames[idx]), t.empty_like(frames[idx])
                saved, self._mode = self._mode, m
                try:
                    self._run(eng, frames[idx].contiguous(), sub_b, sub_o, integer)
                finally:
                    self._mode = saved
                base_out[idx], out[idx] = sub_b, sub_o
            return st0
        self._night = modes[0] == "night"
        return super()._run(eng, frames, base_out, out, integer, st=st0)        # the stage that made the decision is reused

    def _render(self, st):
        lz = st.lz
        bt = st.bands(self.lambdas, [self.UV_BAND, self.B_BAND, self.G_BAND], self.hsi_scale)   # :113-127, :163-166
        stats = st.stats(bt)
        U = st.eval([st.safe_norm(lz.plane(bt, 0), stats, 0)])                             # integrate_uv: normalised
        norm95 = lambda tensor, ch: lz.plane(tensor, ch) / L.maximum(1e-8, st.percentile(tensor, ch, 95.0))   # noqa: E731 (:171-172)
        U_n, B_n, G_n = norm95(U, 0), norm95(bt, 1), norm95(bt, 2)
        false = [L.clip(0.85 * U_n + 0.10 * G_n, 0.0, 1.0), L.clip(0.80 * G_n + 0.20 * B_n, 0.0, 1.0),
                 L.clip(0.70 * B_n + 0.40 * U_n, 0.0, 1.0)]                                 # :174-181
        a = self.uv_boost_alpha
        render = [L.clip((1.0 - a) * c + a * f, 0.0, 1.0) for c, f in zip(st.baseline(), false)]   # :184-185
        night = self._night
        render = scatter_and_blue_bias(st, render, self.night_blur_sigma if night else self.day_blur_sigma,
                                       self.blue_bias_night if night else self.blue_bias_day)   # :193
        if not night:
            render = tone_compress(render, self.tone_strength, self.tone_knee)             # :197
        else:
            Y = L.luma(render)                                                             # :199-202
            gain = (Y + 0.18) / (Y + 1e-6)
            render = [L.clip(c * gain, 0.0, 1.0) for c in render]
        yy = np.linspace(0.0, 1.0, st.H, dtype=np.float32)[:, None]                       # :106-111 ground-focus vignette (per-row table)
        amount = self.ground_vignette_night if night else self.ground_vignette_day
        gain_row = lz.row(1.0 - amount * (1.0 - np.clip(1.0 - yy, 0.0, 1.0)))
        return [L.clip(c * gain_row, 0.0, 1.0) for c in render]

    def visualize(self, image: np.ndarray, *, mode: str = "auto") -> Optional[Tuple[np.ndarray, np.ndarray]]:
        self._mode = mode
        try:
            return super().visualize(image)
        finally:
            self._mode = "auto"
