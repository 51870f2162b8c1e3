"""Pieris -- drop-in for reference animals/pieris.py (constructor :31-67, visualize :69-124)."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal, radial_sigmoid


class Pieris(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(430.0, 500.0), green_band=(500.0, 570.0),
                    panorama_scale=1.05, guide_sigma=1.2, guide_gain=0.75, foliage_opponent_gain=0.25, petal_warmth=0.08,
                    clarity_unsharp_sigma=0.8, clarity_amount=0.22, center_bias=0.12, bias_radius=0.8, bias_softness=7.0)

    def _render(self, st):
        lz = st.lz
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band, self.green_band], self.hsi_scale)
        U, Bv, Gv = st.normed_bands(bt)                                                    # :91-93
        Us_t = st.blur(st.eval([U]), self.guide_sigma)                                     # :96-98
        Us = L.clip(lz.plane(Us_t, 0) / (st.percentile(Us_t, 0, 95.0) + 1e-8), 0.0, 1.0)
        guide_w = self.guide_gain * Us
        tint = np.array([0.35, 0.35 + self.petal_warmth, 0.25], np.float32)                # :100
        render = [L.clip(c + guide_w * float(tint[i]), 0.0, 1.0) for i, c in enumerate(st.baseline())]
        foliage = L.clip(Gv - 0.5 * (U + Bv), 0.0, 1.0)                                    # :102-103
        render[1] = L.clip(render[1] + self.foliage_opponent_gain * foliage, 0.0, 1.0)
        if self.clarity_unsharp_sigma > 0.0 and self.clarity_amount > 0.0:                 # :105-107
            t_img = st.eval(render)
            cur, blurred = lz.channels(t_img), lz.channels(st.blur(t_img, self.clarity_unsharp_sigma))
            render = [L.clip(c + self.clarity_amount * (c - q), 0.0, 1.0) for c, q in zip(cur, blurred)]
        att = lz.keyed(("pieris_att", self.bias_softness, self.bias_radius, self.center_bias),                  # :109-115
                       lambda: 1.0 + self.center_bias * (1.0 - radial_sigmoid(st.H, st.W, self.bias_softness, self.bias_radius)))
        return [L.clip(c * att, 0.0, 1.0) for c in render]
