"""JumpingSpider -- drop-in for reference animals/jumping_spider.py (constructor :41-104, visualize :135-236)."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal, radial_sigmoid, scan_row_gain, unsharp


def attention_spots(H, W, spots, spot_sigma):
    """jumping_spider.py:122-133: Gaussian attention spots normalised by their own 95th percentile (pixel independent)."""
    yy = np.linspace(0.0, 1.0, H, dtype=np.float32)[:, None]
    xx = np.linspace(0.0, 1.0, W, dtype=np.float32)[None, :]
    mask = np.zeros((H, W), np.float32)
    s2 = max(spot_sigma, 1e-4) ** 2
    for yc, xc in spots:
        mask += np.exp(-((yy - yc) ** 2 + (xx - xc) ** 2) / (2.0 * s2))
    m95 = max(1e-8, float(np.percentile(mask, 95.0)))
    return np.clip(mask / m95, 0.0, 1.0).astype(np.float32)


class JumpingSpider(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), green_band=(500.0, 570.0), red_band=(600.0, 680.0),
                    blue_band=(430.0, 500.0), panorama_scale=1.02, dog_small_sigma=0.9, dog_large_sigma=2.2, uv_patch_gain=0.95,
                    opponent_gain=0.3, red_kill=0.25, base_soft_sigma=0.25, clarity_sigma=0.9, clarity_amount=0.24, fovea_radius=0.38,
                    fovea_softness=10.0, periph_blur_sigma=2.2, periph_vignette_strength=0.22, scan_row_freq=22.0, scan_row_gain=0.08,
                    scan_soften=0.9, spots=((0.5, 0.52), (0.57, 0.48)), spot_sigma=0.08, spot_gain=0.2)

    def __init__(self, **kw):
        super().__init__(**kw)
        self.spots = tuple((float(y), float(x)) for (y, x) in self.spots)                  # :104

    def _render(self, st):
        lz = st.lz
        bt = st.bands(self.lambdas, [self.uv_band, self.green_band, self.blue_band], self.hsi_scale)   # :159-162 (red unused)
        U, Gv, Bv = st.normed_bands(bt)
        r, g, b = st.baseline()
        render = [L.clip(r * (1.0 - self.red_kill), 0.0, 1.0), g, b]                       # :165
        if self.base_soft_sigma > 0.0:                                                     # :166-167
            render = lz.channels(st.blur(st.eval(render), self.base_soft_sigma))
        u_t = st.eval([U])                                                                 # :170-174 UV patch saliency (DoG)
        dog_t = st.eval([L.clip(lz.plane(st.blur(u_t, self.dog_small_sigma), 0) - lz.plane(st.blur(u_t, self.dog_large_sigma), 0), 0.0, 1.0)])
        patch = L.clip(lz.plane(dog_t, 0) / (st.percentile(dog_t, 0, 95.0) + 1e-8), 0.0, 1.0)
        opp_t = st.eval([Gv - U, L.absolute(Gv - U)])                                      # :177-179
        opp = L.clip(lz.plane(opp_t, 0) / (st.percentile(opp_t, 1, 95.0) + 1e-8), -1.0, 1.0)
        g_boost = L.clip(opp, 0.0, 1.0) * self.opponent_gain                               # :181-182
        u_boost = L.clip(-opp, 0.0, 1.0) * self.opponent_gain
        r, g, b = render
        g = L.clip(g + 0.40 * g_boost, 0.0, 1.0)                                           # :183-185
        b = L.clip(b + 0.30 * u_boost * Bv, 0.0, 1.0)
        r = L.clip(r + 0.12 * u_boost * U, 0.0, 1.0)
        render = [r, g, b]
        if self.clarity_sigma > 0.0 and self.clarity_amount > 0.0:                         # :188-191
            render = unsharp(st, render, self.clarity_sigma, self.clarity_amount * self.uv_patch_gain * patch)
        if self.scan_row_gain != 0.0:                                                      # :194-203
            rg = lz.keyed(("scan_rows", self.scan_row_freq, self.scan_soften, self.scan_row_gain),
                          lambda: scan_row_gain(st.H, self.scan_row_freq, self.scan_soften, self.scan_row_gain), "row")
            render = [L.clip(c * rg, 0.0, 1.0) for c in render]
        if self.spot_gain > 0.0:                                                           # :206-211
            sm = lz.keyed(("spots", self.spots, self.spot_sigma), lambda: attention_spots(st.H, st.W, self.spots, self.spot_sigma))
            lifted = st.eval([L.clip(c + self.spot_gain * sm, 0.0, 1.0) for c in render])
            sharp = unsharp(st, None, 0.8, 0.25, materialised=lifted)
            render = [L.clip((1.0 - 0.6 * sm) * c + (0.6 * sm) * s, 0.0, 1.0) for c, s in zip(lz.channels(lifted), sharp)]
        if self.periph_blur_sigma > 0.0 or self.periph_vignette_strength > 0.0:            # :214-226
            edge = st.periph_t(self.fovea_softness, self.fovea_radius)
            if self.periph_blur_sigma > 0.0:
                t_img = st.eval(render)
                render = [(1.0 - edge) * c + edge * q for c, q in zip(lz.channels(t_img), lz.channels(st.blur(t_img, self.periph_blur_sigma)))]
            if self.periph_vignette_strength > 0.0:
                vign = lz.keyed(("spider_vign", self.fovea_softness, self.fovea_radius, self.periph_vignette_strength),
                                lambda: 1.0 - self.periph_vignette_strength * radial_sigmoid(st.H, st.W, self.fovea_softness, self.fovea_radius))
                render = [L.clip(c * vign, 0.0, 1.0) for c in render]
        return render
