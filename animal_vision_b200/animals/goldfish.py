"""Goldfish -- drop-in for reference animals/goldfish.py (constructor :35-82, visualize :84-180)."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal, periph_mix


class Goldfish(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
                    red_band=(600.0, 680.0), uv_boost=3.0, panorama_scale=1.45, haze_strength=0.12, haze_tint=(0.78, 0.92, 1.0),
                    red_kill=0.55, green_lift=0.12, blue_lift=0.06, base_blur_sigma=0.8, periph_blur_sigma=1.8, periph_radius=0.65,
                    periph_softness=6.0)

    def _render(self, st):
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band, self.green_band, self.red_band], self.hsi_scale)
        U, Bv, Gv, Rv = st.normed_bands(bt)                                                # :124-127
        ratio = st.eval([U / (1e-6 + 0.45 * Gv + 0.35 * Bv + 0.15 * Rv)])                  # :130
        sal = st.safe_norm(st.lz.plane(ratio, 0), st.stats(ratio), 0)
        r, g, b = st.baseline()
        r = L.clip(r * (1.0 - self.red_kill), 0.0, 1.0)                                    # :136-138
        g = L.clip(g + self.green_lift, 0.0, 1.0)
        b = L.clip(b + self.blue_lift, 0.0, 1.0)
        render = [r, g, b]
        if self.haze_strength > 0.0:                                                       # :141-143
            a = float(np.clip(self.haze_strength, 0.0, 1.0))
            veil = a * np.array(self.haze_tint, np.float32)
            render = [(1.0 - a) * c + float(veil[i]) for i, c in enumerate(render)]
        if self.base_blur_sigma > 0.0:                                                     # :146-147
            render = st.lz.channels(st.blur(st.eval(render), self.base_blur_sigma))
        r, g, b = render
        r = L.clip(r + self.uv_boost * 0.42 * sal, 0.0, 1.0)                               # :152-154
        b = L.clip(b + self.uv_boost * 0.35 * sal, 0.0, 1.0)
        g = L.clip(g + self.uv_boost * 0.12 * sal, 0.0, 1.0)
        b = L.clip(b + 0.22 * Bv, 0.0, 1.0)                                                # :157-158
        g = L.clip(g + 0.30 * Gv, 0.0, 1.0)
        render = [r, g, b]
        if self.periph_blur_sigma > 0.0:                                                   # :161-172
            render = periph_mix(st, render, self.periph_blur_sigma, self.periph_softness, self.periph_radius)
        return render
