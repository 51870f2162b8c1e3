"""Morpho -- drop-in for reference animals/morpho.py (constructor :33-63, visualize :95-154)."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal


class Morpho(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), green_band=(500.0, 570.0),
                    panorama_scale=1.05, sheen_strength=0.55, hue_shift_strength=0.45, gloss_sigma=1.0, mosaic_downscale=0.35,
                    center_clarity=0.25, vignette_softness=7.0, vignette_radius=0.82)

    def __init__(self, **kw):
        super().__init__(**kw)
        self.mosaic_downscale = float(np.clip(self.mosaic_downscale, 0.15, 1.0))           # :59

    def _render(self, st):
        lz, ops = st.lz, st.ops
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band], self.hsi_scale)        # :117-119 (the green map is unused)
        maps = st.eval(st.normed_bands(bt))                                                # (U, Bv)
        gx_t, gy_t = ops.sobel(maps[..., 1:2].contiguous())                                # :65-74 `_grad`
        ori = L.arctan2(lz.plane(gy_t, 0), lz.plane(gx_t, 0))                              # :124
        align = 0.5 * (1.0 + L.cos(2.0 * ori))                                             # :125
        gl_t = st.blur(maps[..., 0:1].contiguous(), self.gloss_sigma)                      # :127-129
        gloss = L.clip(lz.plane(gl_t, 0) / (st.percentile(gl_t, 0, 95.0) + 1e-8), 0.0, 1.0)
        cyan = self.hue_shift_strength * align                                             # :131-132
        deep = self.hue_shift_strength * (1.0 - align)
        r, g, b = st.baseline()
        b = L.clip(b + 0.40 * deep + 0.25 * cyan, 0.0, 1.0)                                # :133-134
        g = L.clip(g + 0.35 * cyan, 0.0, 1.0)
        sheen = self.sheen_strength * gloss                                                # :136
        tint = np.array([0.10, 0.25, 0.45], np.float32)
        render = [L.clip(c + sheen * float(tint[i]), 0.0, 1.0) for i, c in enumerate((r, g, b))]
        t_img = st.eval(render)
        if self.mosaic_downscale < 0.999:                                                  # :85-93 `_mosaic`: INTER_AREA down, INTER_NEAREST up
            h = max(1, int(round(st.H * self.mosaic_downscale)))
            w = max(1, int(round(st.W * self.mosaic_downscale)))
            t_img = ops.resize(ops.resize(t_img, (h, w), "area"), (st.H, st.W), "nearest")
        cur, blurred = lz.channels(t_img), lz.channels(st.blur(t_img, 1.0))               # :146-148
        t = st.periph_t(self.vignette_softness, self.vignette_radius)
        return [L.clip((1.0 - t) * (c + 0.22 * (c - q)) + t * c, 0.0, 1.0) for c, q in zip(cur, blurred)]
