"""Dragonfly -- drop-in for reference animals/dragonfly.py (constructor :40-117, visualize :146-251)."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal, periph_mix, unsharp


def soft_knee(img, knee, amount):
    """dragonfly.py:133-144 `_soft_knee` (like uv_helpers' tone compression, with 1e-8 in the knee span)."""
    if amount <= 0.0:
        return list(img)
    out = []
    for c in img:
        x = L.clip(c, 0.0, 1.0)
        t = (x - knee) / (1.0 - knee + 1e-8)
        out.append(L.where(x <= knee, x, knee + (1.0 - knee) * (t / (1.0 + amount * t))))
    return out


class Dragonfly(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), green_band=(500.0, 570.0),
                    red_band=(600.0, 680.0), panorama_scale=1.15, sky_prior_strength=0.6, sky_blue_weight=0.4, sky_sigmoid_mid=0.46,
                    sky_sigmoid_steepness=6.0, sky_pol_strength=0.65, sky_pol_gamma=1.3, water_pol_strength=0.55, water_pol_gamma=1.2,
                    sky_evec_base_deg=90.0, sky_evec_sweep_deg=-45.0, red_kill=0.22, sky_uv_blue_gain=(0.25, 0.2),
                    water_uv_blue_gain=(0.3, 0.24), ventral_green_gain=0.12, base_soft_sigma=0.3, unsharp_sigma=1.0, unsharp_amount=0.3,
                    highlight_knee=0.85, highlight_strength=0.35, periph_blur_sigma=0.7, periph_radius=0.8, periph_softness=7.0)

    def __init__(self, **kw):
        super().__init__(**kw)
        self.sky_evec_base = np.deg2rad(float(self.sky_evec_base_deg))                      # :104-105
        self.sky_evec_sweep = np.deg2rad(float(self.sky_evec_sweep_deg))
        self.sky_uv_blue_gain = tuple(map(float, self.sky_uv_blue_gain))
        self.water_uv_blue_gain = tuple(map(float, self.water_uv_blue_gain))

    def _render(self, st):
        lz, ops = st.lz, st.ops
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band, self.green_band], self.hsi_scale)   # :172-175
        U, Bv, Gv = st.normed_bands(bt)
        prior = lz.row(np.linspace(1.0, 0.0, st.H, dtype=np.float32))                      # :179
        blue_dom = L.clip(Bv - 0.6 * Gv, 0.0, 1.0)
        sc_t = st.blur(st.eval([self.sky_prior_strength * prior + self.sky_blue_weight * blue_dom]), 2.5)   # :181-182
        score = lz.plane(sc_t, 0) / (st.percentile(sc_t, 0, 98.0) + 1e-8)                  # :183 (not clipped)
        sky_w = 1.0 / (1.0 + L.exp(-self.sky_sigmoid_steepness * (score - self.sky_sigmoid_mid)))   # :184
        gx_t, gy_t = ops.sobel(st.eval([0.6 * Bv + 0.4 * U]))                              # :190-192
        theta = L.arctan2(lz.plane(gy_t, 0), lz.plane(gx_t, 0))
        y_norm = np.linspace(0.0, 1.0, st.H, dtype=np.float32)[:, None]                    # :194-195: E-vector per image row
        sky_evec = self.sky_evec_base + self.sky_evec_sweep * y_norm
        cos2_sky, sin2_sky = lz.row(np.cos(2.0 * sky_evec)), lz.row(np.sin(2.0 * sky_evec))
        c2, s2 = L.cos(2.0 * theta), L.sin(2.0 * theta)                                    # :197-198
        align_sky01 = L.clip(0.5 * ((c2 * cos2_sky + s2 * sin2_sky) + 1.0), 0.0, 1.0) ** self.sky_pol_gamma     # :202-203
        align_water01 = L.clip(0.5 * ((c2 * 1.0 + s2 * 0.0) + 1.0), 0.0, 1.0) ** self.water_pol_gamma          # :206-208
        m_t = st.eval([sky_w, align_sky01, align_water01])                                 # three masks, materialised once
        sky_w, align_sky01, align_water01 = lz.channels(m_t)
        ground_w = 1.0 - sky_w
        r, g, b = st.baseline()
        render = [L.clip(r * (1.0 - self.red_kill), 0.0, 1.0), g, b]                       # :212
        if self.base_soft_sigma > 0.0:                                                     # :213-214
            render = lz.channels(st.blur(st.eval(render), self.base_soft_sigma))
        sky_gain = 1.0 + self.sky_pol_strength * (align_sky01 * sky_w)                     # :217
        render = [L.clip(c * (0.95 + 0.05 * sky_w), 0.0, 1.0) for c in render]             # :218
        r, g, b = render
        b = L.clip(b + self.sky_uv_blue_gain[1] * (Bv * sky_w * align_sky01), 0.0, 1.0)    # :219-220
        g = L.clip(g + 0.10 * (U * sky_w * align_sky01), 0.0, 1.0)
        render = [L.clip(c * sky_gain, 0.0, 1.0) for c in (r, g, b)]                       # :221
        water_gain = 1.0 + self.water_pol_strength * (align_water01 * ground_w)            # :224
        r, g, b = render
        b = L.clip(b + self.water_uv_blue_gain[1] * (Bv * ground_w * align_water01), 0.0, 1.0)   # :225-228
        b = L.clip(b + self.water_uv_blue_gain[0] * (U * ground_w * align_water01), 0.0, 1.0)
        g = L.clip(g + self.ventral_green_gain * (Gv * ground_w), 0.0, 1.0)                # :229
        render = [L.clip(c * water_gain, 0.0, 1.0) for c in (r, g, b)]                     # :230
        if self.unsharp_sigma > 0.0 and self.unsharp_amount > 0.0:                         # :233-236
            render = unsharp(st, render, self.unsharp_sigma, self.unsharp_amount)
        render = soft_knee(render, self.highlight_knee, self.highlight_strength)           # :237
        if self.periph_blur_sigma > 0.0:                                                   # :240-246
            render = periph_mix(st, render, self.periph_blur_sigma, self.periph_softness, self.periph_radius)
        return render
