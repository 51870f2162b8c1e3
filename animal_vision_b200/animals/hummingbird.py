"""Hummingbird -- drop-in for reference animals/hummingbird.py (constructor :39-102, visualize :128-227)."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal, periph_mix


def _s2l(rgb):
    """hummingbird.py:96-99: sRGB 0..255 target -> linear float32."""
    v = np.array(rgb, np.float32) / 255.0
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / (1 + 0.055)) ** 2.4).astype(np.float32)


class Hummingbird(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
                    red_band=(600.0, 680.0), panorama_scale=1.05, red_kill=0.1, base_soft_sigma=0.25, unsharp_sigma=0.9,
                    unsharp_amount=0.24, combo_opacity=0.55, combo_saturation=0.45, combo_sheen=0.28, tgt_uvb_srgb=(120, 150, 255),
                    tgt_uvg_srgb=(110, 255, 170), tgt_uvr_srgb=(255, 110, 210), guide_sigma=1.0, guide_gain=0.25, periph_blur_sigma=0.6,
                    periph_radius=0.82, periph_softness=7.0)

    def __init__(self, **kw):
        super().__init__(**kw)
        self.tgt_lin = [_s2l(self.tgt_uvb_srgb), _s2l(self.tgt_uvg_srgb), _s2l(self.tgt_uvr_srgb)]   # :100-102

    def _render(self, st):
        lz = st.lz
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band, self.green_band, self.red_band], self.hsi_scale)   # :150-153
        U, Bv, Gv, Rv = st.normed_bands(bt)
        prod = st.eval([U * Bv, U * Gv, U * Rv])                                           # :156-158
        pst = st.stats(prod)
        combos = st.eval([st.safe_norm(lz.plane(prod, k), pst, k) for k in range(3)])
        small, large = st.blur(combos, 0.8), st.blur(combos, 2.0)                          # :160-165 `bandpass`
        d_t = st.eval([L.clip(lz.plane(small, k) - lz.plane(large, k), 0.0, 1.0) for k in range(3)])
        bp_t = st.eval([L.clip(lz.plane(d_t, k) / (st.percentile(d_t, k, 95.0) + 1e-8), 0.0, 1.0) for k in range(3)])
        cb, cg, cr = lz.channels(bp_t)
        r, g, b = st.baseline()
        render = [L.clip(r * (1.0 - self.red_kill), 0.0, 1.0), g, b]                       # :172
        if self.base_soft_sigma > 0.0:                                                     # :173-174
            render = lz.channels(st.blur(st.eval(render), self.base_soft_sigma))
        cmax = L.maximum(L.maximum(cb, cg), cr)                                            # :177
        if self.unsharp_sigma > 0.0 and self.unsharp_amount > 0.0:                         # :178-181
            t_img = st.eval(render)
            cur, blurred = lz.channels(t_img), lz.channels(st.blur(t_img, self.unsharp_sigma))
            amt = self.unsharp_amount * cmax
            render = [L.clip(c + amt * L.clip(c - q, -1.0, 1.0), 0.0, 1.0) for c, q in zip(cur, blurred)]
        if self.combo_sheen > 0.0:                                                         # :184-186
            sheen = 0.55 * cb + 0.65 * cg + 0.75 * cr
            render = [L.clip(c + self.combo_sheen * sheen, 0.0, 1.0) for c in render]
        w_sum = cb + cg + cr + 1e-8                                                        # :189-196
        wB, wG, wR = cb / w_sum, cg / w_sum, cr / w_sum
        tb, tg, tr = self.tgt_lin
        tint = [wB * float(tb[i]) + wG * float(tg[i]) + wR * float(tr[i]) for i in range(3)]
        s = 1.0 + self.combo_saturation                                                    # :197, :121-126 `_apply_saturation`
        if s != 1.0:
            Y = L.luma(tint)
            tint = [L.clip(Y + (c - Y) * s, 0.0, 1.0) for c in tint]
        render = [L.clip((1.0 - self.combo_opacity) * c + self.combo_opacity * tc, 0.0, 1.0) for c, tc in zip(render, tint)]   # :198
        if self.guide_gain > 0.0:                                                          # :201-205
            us_t = st.blur(st.eval([U]), self.guide_sigma)
            Us = L.clip(lz.plane(us_t, 0) / (st.percentile(us_t, 0, 95.0) + 1e-8), 0.0, 1.0)
            arr = np.array([0.20, 0.25, 0.10], np.float32)
            render = [L.clip(c + self.guide_gain * Us * float(arr[i]), 0.0, 1.0) for i, c in enumerate(render)]
        if self.periph_blur_sigma > 0.0:                                                   # :208-215
            render = periph_mix(st, render, self.periph_blur_sigma, self.periph_softness, self.periph_radius)
        return render
