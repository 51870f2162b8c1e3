"""Anableps -- drop-in for reference animals/anableps.py (constructor :39-114, visualize :124-255)."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal, periph_mix, unsharp


class Anableps(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
                    red_band=(600.0, 680.0), panorama_scale=1.2, horizon_y=0.44, seam_softness_px=8.0, ripple_amp_px=6.0,
                    ripple_waves=2.5, refract_push_px=3.0, air_warmth=(1.06, 1.03, 0.99), air_clarity_unsharp=0.35, air_unsharp_sigma=1.0,
                    red_kill=0.55, blue_lift=0.08, green_lift=0.12, haze_strength=0.1, haze_tint=(0.8, 0.92, 1.0),
                    base_blur_sigma_water=0.7, uv_boost=3.4, uv_R_gain=0.36, uv_G_gain=0.18, uv_B_gain=0.42, periph_blur_sigma=1.2,
                    periph_radius=0.7, periph_softness=6.0)

    def _geometry(self, H, W):
        """anableps.py:171-186, :224-231: rippled horizon, air weight and the refraction map -- pixel independent tables
        built with the reference's own NumPy statements."""
        y0 = int(np.clip(self.horizon_y * H, 0, H - 1))
        if self.ripple_amp_px > 0.0:
            x = np.linspace(0, 2.0 * np.pi * self.ripple_waves, W, dtype=np.float32)
            ripple = (self.ripple_amp_px * np.sin(x)).astype(np.float32)
        else:
            ripple = np.zeros((W,), np.float32)
        yy = np.arange(H, dtype=np.float32)[:, None]
        seam_soft = max(1.0, float(self.seam_softness_px))
        horizon = y0 + ripple[None, :]
        air_w = 1.0 / (1.0 + np.exp(+(yy - horizon) / seam_soft))
        maps = None
        if self.refract_push_px > 0.0:
            yi = np.repeat(np.arange(H, dtype=np.float32)[:, None], W, axis=1)
            xi = np.repeat(np.arange(W, dtype=np.float32)[None, :], H, axis=0)
            push = self.refract_push_px * np.exp(-np.maximum(yi - horizon, 0.0) / (2.5 * self.seam_softness_px))
            maps = (xi.astype(np.float32), np.clip(yi + push, 0, H - 1).astype(np.float32))
        return air_w.astype(np.float32), maps

    def _render(self, st):
        lz = st.lz
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band, self.green_band], self.hsi_scale)   # :158-164
        Un, Bv, Gv = st.normed_bands(bt)
        gkey = ("anableps_geo", self.horizon_y, self.seam_softness_px, self.ripple_amp_px, self.ripple_waves, self.refract_push_px, st.H, st.W)
        geo = st.eng.cached(gkey, lambda: self._geometry(st.H, st.W))                     # host tables built once per geometry
        air_np, maps = geo
        air_w = lz.keyed(gkey, lambda: air_np)
        warm = np.array(self.air_warmth, np.float32)
        air = [L.clip(c * float(warm[i]), 0.0, 1.0) for i, c in enumerate(st.baseline())]   # :190-191
        if self.air_unsharp_sigma > 0.0 and self.air_clarity_unsharp > 0.0:                 # :192, :116-122
            air = unsharp(st, air, self.air_unsharp_sigma, self.air_clarity_unsharp)
        air_t = st.eval(air)
        r, g, b = st.baseline()                                                            # :195-198
        water = [L.clip(r * (1.0 - self.red_kill), 0.0, 1.0), L.clip(g + self.green_lift, 0.0, 1.0), L.clip(b + self.blue_lift, 0.0, 1.0)]
        if self.haze_strength > 0.0:                                                       # :200-202
            a = float(np.clip(self.haze_strength, 0.0, 1.0))
            veil = a * np.array(self.haze_tint, np.float32)
            water = [(1.0 - a) * c + float(veil[i]) for i, c in enumerate(water)]
        if self.base_blur_sigma_water > 0.0:                                               # :204-205
            water = lz.channels(st.blur(st.eval(water), self.base_blur_sigma_water))
        r, g, b = water
        r = L.clip(r + self.uv_boost * self.uv_R_gain * Un, 0.0, 1.0)                      # :208-211
        g = L.clip(g + self.uv_boost * self.uv_G_gain * Un, 0.0, 1.0)
        b = L.clip(b + self.uv_boost * self.uv_B_gain * Un, 0.0, 1.0)
        b = L.clip(b + 0.20 * Bv, 0.0, 1.0)                                                # :214-215
        g = L.clip(g + 0.26 * Gv, 0.0, 1.0)
        water_t = st.eval([r, g, b])
        if maps is not None:                                                               # :218-237 refraction remap
            water_t = st.ops.remap(water_t, maps[0], maps[1], key=gkey)
        water_w = 1.0 - air_w
        render = [a_ * air_w + w_ * water_w for a_, w_ in zip(lz.channels(air_t), lz.channels(water_t))]   # :240
        if self.periph_blur_sigma > 0.0:                                                   # :243-252
            render = periph_mix(st, render, self.periph_blur_sigma, self.periph_softness, self.periph_radius)
        return render
