"""Kestrel -- drop-in for reference animals/kestrel.py (constructor :40-98, ridge measure :113-136, visualize :138-234)."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal, periph_mix


class Kestrel(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), green_band=(500.0, 570.0),
                    red_band=(600.0, 680.0), panorama_scale=1.1, sky_cool_tint=(0.95, 0.98, 1.03), sky_haze=0.1,
                    ground_warm_tint=(1.02, 1.01, 0.99), ground_contrast=0.08, uv_overlay_strength=0.55, uv_magenta=(0.6, 0.12, 0.7),
                    ridge_sigma=3, ridge_gain=1.0, unsharp_sigma=1.0, unsharp_amount=0.3, periph_blur_sigma=0.7, periph_radius=0.82,
                    periph_softness=7.0)

    def _ridge(self, st, u_t, sigma):
        """Structure-tensor coherence x energy of the UV map (kestrel.py:113-136)."""
        lz = st.lz
        gx_t, gy_t = st.ops.sobel(u_t)
        gx, gy = lz.plane(gx_t, 0), lz.plane(gy_t, 0)
        tens = st.blur(st.eval([gx * gx, gy * gy, gx * gy]), sigma)                       # :121-123
        gxx, gyy, gxy = lz.channels(tens)
        trace, diff = gxx + gyy, gxx - gyy
        root = L.sqrt(L.maximum((0.5 * diff) ** 2 + gxy * gxy, 0.0))                      # :126
        lam1, lam2 = 0.5 * trace + root, 0.5 * trace - root
        coh = (lam1 - lam2) / (lam1 + lam2 + 1e-8)                                         # :130
        en_t = st.eval([L.clip(trace, 0.0, None)])                                         # :131-132
        energy = lz.plane(en_t, 0) / (st.percentile(en_t, 0, 95.0) + 1e-8)
        return L.clip(coh * energy, 0.0, 1.0)

    def _render(self, st):
        lz = st.lz
        cool, warm, magenta = (np.array(v, np.float32) for v in (self.sky_cool_tint, self.ground_warm_tint, self.uv_magenta))
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band, self.green_band], self.hsi_scale)   # :156-159
        U, Bv, Gv = st.normed_bands(bt)
        prior = lz.row(np.linspace(1.0, 0.0, st.H, dtype=np.float32))                      # :162
        blue_dom = L.clip(Bv - 0.6 * Gv, 0.0, 1.0)
        sky_t = st.blur(st.eval([0.6 * prior + 0.4 * blue_dom]), 3.0)                      # :164-165
        sky = L.clip(lz.plane(sky_t, 0) / (st.percentile(sky_t, 0, 98.0) + 1e-8), 0.0, 1.0)
        sky_w = 1.0 / (1.0 + L.exp(-6.0 * (sky - 0.45)))                                   # :168
        u_t = st.eval([U])
        ridge = self._ridge(st, u_t, float(self.ridge_sigma))
        w_t = st.eval([sky_w, L.clip(float(self.ridge_gain) * ridge * (1.0 - sky_w), 0.0, 1.0)])   # (sky_w, trailness) :174-175
        sky_w, trail = lz.plane(w_t, 0), lz.plane(w_t, 1)
        ground_w = 1.0 - sky_w
        render = st.baseline()
        tinted = [L.clip(c * float(cool[i]), 0.0, 1.0) for i, c in enumerate(render)]
        if self.sky_haze > 0.0:                                                            # :180-185
            a = float(np.clip(self.sky_haze, 0.0, 1.0))
            veil = a * np.array([0.90, 0.97, 1.00], np.float32)
            render = [sky_w * ((1.0 - a) * tc + float(veil[i])) + ground_w * c for i, (tc, c) in enumerate(zip(tinted, render))]
        else:
            render = [sky_w * tc + ground_w * c for tc, c in zip(tinted, render)]
        r_t = st.eval(render)
        render = lz.channels(r_t)
        ground = [L.clip(c * float(warm[i]), 0.0, 1.0) for i, c in enumerate(render)]      # :188-189
        if self.ground_contrast > 0.0:                                                     # :190-192
            g_t = st.eval(ground)
            cur, blurred = lz.channels(g_t), lz.channels(st.blur(g_t, 1.2))
            ground = [L.clip(c + self.ground_contrast * (c - q), 0.0, 1.0) for c, q in zip(cur, blurred)]
        render = [sky_w * c + ground_w * gp for c, gp in zip(render, ground)]              # :195
        U95 = L.clip(lz.plane(u_t, 0) / (st.percentile(u_t, 0, 95.0) + 1e-8), 0.0, 1.0)    # :198-200
        k = self.uv_overlay_strength * ground_w
        render = [L.clip((1.0 - k) * c + k * (U95 * float(magenta[i])), 0.0, 1.0) for i, c in enumerate(render)]   # :201-203
        if self.unsharp_sigma > 0.0 and self.unsharp_amount > 0.0:                         # :206-209
            t_img = st.eval(render)
            cur, blurred = lz.channels(t_img), lz.channels(st.blur(t_img, self.unsharp_sigma))
            amt = self.unsharp_amount * trail
            render = [L.clip(c + amt * L.clip(c - q, -1.0, 1.0), 0.0, 1.0) for c, q in zip(cur, blurred)]
        if self.periph_blur_sigma > 0.0:                                                   # :212-218
            render = periph_mix(st, render, self.periph_blur_sigma, self.periph_softness, self.periph_radius)
        return render
