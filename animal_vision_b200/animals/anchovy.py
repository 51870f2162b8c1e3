"""Anchovy -- drop-in for reference animals/anchovy.py (constructor :38-108, visualize :130-253)."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal, periph_mix


class Anchovy(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), green_band=(500.0, 570.0),
                    red_band=(600.0, 680.0), panorama_scale=1.2, red_kill=0.25, base_soft_sigma=0.3, unsharp_sigma=1.0,
                    unsharp_amount=0.35, haze_strength=0.04, haze_tint=(0.9, 0.97, 1.0), evec_angle_deg=0.0, pol_strength=0.55,
                    pol_gamma=1.2, orientation_mix=0.35, uv_gloss_gain=0.28, blue_chroma_gain=0.18, green_chroma_gain=0.1,
                    periph_blur_sigma=0.6, periph_radius=0.78, periph_softness=7.0)

    def __init__(self, **kw):
        super().__init__(**kw)
        self.evec_angle = np.deg2rad(float(self.evec_angle_deg))                            # :96
        self.orientation_mix = float(np.clip(self.orientation_mix, 0.0, 1.0))              # :99

    def _render(self, st):
        lz, ops = st.lz, st.ops
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band, self.green_band], self.hsi_scale)   # :163-171
        Un, Bn, Gn = st.normed_bands(bt)
        gx_t, gy_t = ops.sobel(st.eval([Un]))                                              # :176
        gx, gy = lz.plane(gx_t, 0), lz.plane(gy_t, 0)
        theta = L.arctan2(gy, gx)                                                          # :177
        mix = self.orientation_mix
        align = (1.0 - mix) * float(np.cos(2.0 * self.evec_angle)) + mix * L.cos(2.0 * theta)   # :180-189
        align01 = L.clip(0.5 * (align + 1.0), 0.0, 1.0) ** float(self.pol_gamma)           # :191
        mag_t = st.eval([L.sqrt(gx * gx + gy * gy)])                                       # :194-196
        mag = L.clip(lz.plane(mag_t, 0) / (st.percentile(mag_t, 0, 95.0) + 1e-8), 0.0, 1.0)
        pol_gain = 1.0 + self.pol_strength * (align01 * Un * mag)                          # :199
        r, g, b = st.baseline()
        render = [L.clip(r * (1.0 - self.red_kill), 0.0, 1.0), g, b]                       # :204
        if self.haze_strength > 0.0:                                                       # :207-209
            a = float(np.clip(self.haze_strength, 0.0, 1.0))
            veil = a * np.array(self.haze_tint, np.float32)
            render = [(1.0 - a) * c + float(veil[i]) for i, c in enumerate(render)]
        if self.base_soft_sigma > 0.0:                                                     # :212-213
            render = lz.channels(st.blur(st.eval(render), self.base_soft_sigma))
        if self.unsharp_sigma > 0.0 and self.unsharp_amount > 0.0:                         # :216-219
            t_img = st.eval(render)
            cur, blurred = lz.channels(t_img), lz.channels(st.blur(t_img, self.unsharp_sigma))
            k = self.unsharp_amount * pol_gain
            render = [L.clip(c + k * L.clip(c - q, -1.0, 1.0), 0.0, 1.0) for c, q in zip(cur, blurred)]
        gloss = self.uv_gloss_gain * (align01 * Un)                                        # :222-224
        r, g, b = render
        b = L.clip(b + 0.70 * gloss, 0.0, 1.0)
        g = L.clip(g + 0.30 * gloss, 0.0, 1.0)
        b = L.clip(b + self.blue_chroma_gain * (Bn * Un), 0.0, 1.0)                        # :227-228
        g = L.clip(g + self.green_chroma_gain * (Gn * Un), 0.0, 1.0)
        render = [r, g, b]
        if self.periph_blur_sigma > 0.0:                                                   # :231-240
            render = periph_mix(st, render, self.periph_blur_sigma, self.periph_softness, self.periph_radius)
        return render
