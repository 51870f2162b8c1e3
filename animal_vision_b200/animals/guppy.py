"""Guppy -- drop-in for reference animals/guppy.py (constructor :41-100, visualize :132-235)."""
import numpy as np

from .. import lazy as L
from .heliconius import sat_apply
from .uvbase import UVAnimal, radial_sigmoid


class Guppy(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
                    red_band=(600.0, 680.0), panorama_scale=1.22, red_kill=0.28, haze_strength=0.06, haze_tint=(0.92, 0.98, 1.0),
                    warm_tint=(1.03, 1.01, 0.99), base_soft_sigma=0.35, unsharp_sigma=0.9, unsharp_amount=0.28, dog_small_sigma=0.8,
                    dog_large_sigma=2.4, dog_gain=0.85, uv_chroma_boost=0.4, uv_blue_gain=0.55, uv_green_gain=0.35, uv_red_gain=0.12,
                    background_desat=0.18, vignette_strength=0.12, vignette_radius=0.78, vignette_softness=7.0)

    def _render(self, st):
        lz = st.lz
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band, self.green_band], self.hsi_scale)   # :163-171
        Un, Bn, Gn = st.normed_bands(bt)
        r, g, b = st.baseline()
        render = [L.clip(r * (1.0 - self.red_kill), 0.0, 1.0), g, b]                       # :175
        if self.haze_strength > 0.0:                                                       # :177-179
            a = float(np.clip(self.haze_strength, 0.0, 1.0))
            veil = a * np.array(self.haze_tint, np.float32)
            render = [(1.0 - a) * c + float(veil[i]) for i, c in enumerate(render)]
        warm = np.array(self.warm_tint, np.float32)                                        # :181
        render = [L.clip(c * float(warm[i]), 0.0, 1.0) for i, c in enumerate(render)]
        if self.base_soft_sigma > 0.0:                                                     # :183-184
            render = lz.channels(st.blur(st.eval(render), self.base_soft_sigma))
        un_t = st.eval([Un])                                                               # :187-191 DoG spot saliency
        dog_t = st.eval([L.clip(lz.plane(st.blur(un_t, self.dog_small_sigma), 0) - lz.plane(st.blur(un_t, self.dog_large_sigma), 0), 0.0, 1.0)])
        spot = L.clip(lz.plane(dog_t, 0) / (st.percentile(dog_t, 0, 95.0) + 1e-8), 0.0, 1.0)
        if self.unsharp_sigma > 0.0 and self.unsharp_amount > 0.0:                         # :194-197
            t_img = st.eval(render)
            cur, blurred = lz.channels(t_img), lz.channels(st.blur(t_img, self.unsharp_sigma))
            k = self.unsharp_amount * spot
            render = [L.clip(c + k * L.clip(c - q, -1.0, 1.0), 0.0, 1.0) for c, q in zip(cur, blurred)]
        lift = self.uv_chroma_boost * spot                                                 # :200-203
        r, g, b = render
        b = L.clip(b + self.uv_blue_gain * lift * Bn, 0.0, 1.0)
        g = L.clip(g + self.uv_green_gain * lift * Gn, 0.0, 1.0)
        r = L.clip(r + self.uv_red_gain * lift * Un, 0.0, 1.0)
        t_img = st.eval([r, g, b])                                                         # :106-109 `_saturation`: mean |lin - Y| / P95
        cur = lz.channels(t_img)
        Y = L.luma(cur)
        mc_t = st.eval([(L.absolute(cur[0] - Y) + L.absolute(cur[1] - Y) + L.absolute(cur[2] - Y)) / 3.0])
        sat = lz.plane(mc_t, 0) / (st.percentile(mc_t, 0, 95.0) + 1e-8)
        render = sat_apply(cur, 1.0 - self.background_desat * (1.0 - Un) * (1.0 - sat))    # :207-208
        if self.vignette_strength > 0.0:                                                   # :211-218
            vign = lz.keyed(("guppy_vign", self.vignette_softness, self.vignette_radius, self.vignette_strength),
                            lambda: 1.0 - self.vignette_strength * radial_sigmoid(st.H, st.W, self.vignette_softness, self.vignette_radius))
            render = [L.clip(c * vign, 0.0, 1.0) for c in render]
        return render
