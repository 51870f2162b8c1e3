"""Reindeer -- drop-in for reference animals/reindeer.py (constructor :41-66, visualize :70-135)."""
from .. import lazy as L
from .uvbase import UVAnimal, scatter_and_blue_bias, tone_compress


class Reindeer(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(300.0, 410.0), uv_boost=3.5, snow_glare_compression=0.55,
                    winter_mode=True, scatter_sigma=1.2, blue_bias=0.08, panorama_scale=1.3, return_uv_heatmap=True)

    def _render(self, st):
        bt = st.bands(self.lambdas, [self.uv_band, (420.0, 680.0)], self.hsi_scale)       # :101-113
        uv_map, vis_map = st.normed_bands(bt)
        ratio = st.eval([uv_map / (1e-6 + 0.6 * vis_map)])                                # :116
        sal = st.safe_norm(st.lz.plane(ratio, 0), st.stats(ratio), 0)
        r, g, b = st.baseline()
        b = L.clip(b + self.uv_boost * 0.35 * sal, 0.0, 1.0)                              # :121-122
        g = L.clip(g + self.uv_boost * 0.15 * sal, 0.0, 1.0)
        render = tone_compress([r, g, b], self.snow_glare_compression)                    # :125
        if self.winter_mode:
            render = scatter_and_blue_bias(st, render, self.scatter_sigma, self.blue_bias)   # :128-129
        return render
