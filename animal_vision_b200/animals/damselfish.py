"""Damselfish -- drop-in for reference animals/damselfish.py (constructor :39-85, visualize :87-181)."""
from .. import lazy as L
from .uvbase import UVAnimal, periph_mix


class Damselfish(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, uv_band=(320.0, 400.0), blue_band=(440.0, 500.0), yellow_band=(560.0, 600.0),
                    red_band=(600.0, 680.0), uv_edge_boost=0.45, uv_gloss_boost=0.3, blue_chroma_gain=0.22, yellow_chroma_gain=0.28,
                    red_kill=0.35, base_blur_sigma=0.35, unsharp_sigma=1.2, panorama_scale=1.25, periph_radius=0.7,
                    periph_softness=7.0, periph_extra_blur=0.8)

    def _render(self, st):
        bt = st.bands(self.lambdas, [self.uv_band, self.blue_band, self.yellow_band], self.hsi_scale)   # :118-127 (the red map is unused)
        Un, Bn, Yn = st.normed_bands(bt)          # safe_norm(integrate_uv(.)) normalises twice: the second pass is the identity
        r, g, b = st.baseline()
        render = [L.clip(r * (1.0 - self.red_kill), 0.0, 1.0), g, b]                       # :130
        if self.base_blur_sigma > 0.0:                                                     # :133-134
            render = st.lz.channels(st.blur(st.eval(render), self.base_blur_sigma))
        if self.unsharp_sigma > 0.0 and self.uv_edge_boost > 0.0:                          # :137-141
            t_img = st.eval(render)
            cur, blurred = st.lz.channels(t_img), st.lz.channels(st.blur(t_img, self.unsharp_sigma))
            gain = 1.0 + self.uv_edge_boost * Un
            render = [L.clip(c + gain * L.clip(c - q, -1.0, 1.0), 0.0, 1.0) for c, q in zip(cur, blurred)]
        r, g, b = render
        if self.uv_gloss_boost > 0.0:                                                      # :144-148
            lift = self.uv_gloss_boost * Un
            b = L.clip(b + 0.60 * lift, 0.0, 1.0)
            g = L.clip(g + 0.30 * lift, 0.0, 1.0)
            r = L.clip(r + 0.15 * lift, 0.0, 1.0)
        b = L.clip(b + self.blue_chroma_gain * Bn, 0.0, 1.0)                               # :151
        y_boost = self.yellow_chroma_gain * Yn                                             # :152-154
        g = L.clip(g + 0.65 * y_boost, 0.0, 1.0)
        r = L.clip(r + 0.35 * y_boost, 0.0, 1.0)
        render = [r, g, b]
        if self.periph_extra_blur > 0.0:                                                   # :157-166
            render = periph_mix(st, render, self.periph_extra_blur, self.periph_softness, self.periph_radius)
        return render
