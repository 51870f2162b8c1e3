"""Heliconius -- drop-in for reference animals/heliconius.py (constructor :33-58, visualize :66-135)."""
from .. import lazy as L
from .uvbase import UVAnimal


def sat_apply(lin, scale):
    """heliconius.py:62-64 `_sat_apply`: clip(Y + (lin - Y) * scale) around the Rec.709 luma."""
    Y = L.luma(lin)
    return [L.clip(Y + (c - Y) * scale, 0.0, 1.0) for c in lin]


class Heliconius(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), red_band=(600.0, 680.0), green_band=(500.0, 570.0),
                    panorama_scale=1.05, conj_sigma_small=0.8, conj_sigma_large=2.2, conj_gain=1.0, sat_boost=0.45, red_gain=0.4,
                    bg_desat=0.2, bg_cool=0.04, base_soft_sigma=0.3, unsharp_sigma=1.0, unsharp_amount=0.25)

    def _render(self, st):
        lz = st.lz
        bt = st.bands(self.lambdas, [self.uv_band, self.red_band], self.hsi_scale)         # :88-90 (the green map is unused)
        maps = st.eval(st.normed_bands(bt))                                                # (U, Rb)
        small, large = st.blur(maps, self.conj_sigma_small), st.blur(maps, self.conj_sigma_large)   # :92-95
        dog = [L.clip(lz.plane(small, c) - lz.plane(large, c), 0.0, 1.0) for c in range(2)]
        conj_t = st.eval([dog[0] * dog[1]])                                                # :98
        conj = L.clip(lz.plane(conj_t, 0) / (st.percentile(conj_t, 0, 95.0) + 1e-8), 0.0, 1.0)   # :99-100
        render = st.baseline()
        if self.base_soft_sigma > 0.0:                                                     # :103-104
            render = lz.channels(st.blur(st.baseline_lin, self.base_soft_sigma))
        bg_w = 1.0 - conj
        render[2] = L.clip(render[2] + self.bg_cool * bg_w, 0.0, 1.0)                      # :107
        render = sat_apply(render, 1.0 - self.bg_desat * bg_w)                             # :108-109
        if self.unsharp_sigma > 0.0 and self.unsharp_amount > 0.0:                         # :111-113
            t_img = st.eval(render)
            cur, blurred = lz.channels(t_img), lz.channels(st.blur(t_img, self.unsharp_sigma))
            k = self.unsharp_amount * conj
            render = [L.clip(c + k * (c - q), 0.0, 1.0) for c, q in zip(cur, blurred)]
        render[0] = L.clip(render[0] + self.red_gain * conj, 0.0, 1.0)                     # :115
        return sat_apply(render, 1.0 + self.sat_boost * conj)                              # :116
