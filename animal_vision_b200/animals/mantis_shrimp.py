"""MantisShrimp -- drop-in for reference animals/mantis_shrimp.py (constructor :39-120, visualize :143-279): the
classic-HSI route the reference actually runs (ten narrow bands of the analytic spectrum), SURVEY.md 8f-2."""
import numpy as np

from .. import lazy as L
from .uvbase import UVAnimal, periph_mix, scan_row_gain


def barcode_lut(N: int) -> np.ndarray:
    """mantis_shrimp.py:172-196: N hues around the wheel at s = 0.95, v = 1 -> [N,3] float32."""
    h = (np.arange(N, dtype=np.float32) / max(N, 1)).astype(np.float32)
    s, v = np.full_like(h, 0.95, np.float32), np.ones_like(h, np.float32)
    i = np.floor(h * 6.0).astype(np.int32)
    f = h * 6.0 - i
    p, q, t = v * (1.0 - s), v * (1.0 - f * s), v * (1.0 - (1.0 - f) * s)
    i = i % 6
    conds = [i == 0, i == 1, i == 2, i == 3, i == 4, i == 5]
    return np.stack([np.select(conds, [v, q, p, p, t, v], default=v), np.select(conds, [t, v, v, q, p, p], default=v),
                     np.select(conds, [p, p, t, v, v, q], default=v)], axis=-1).astype(np.float32)


def _sum_numpy_order(vals):
    """np.sum over a short last axis: eight running lanes combined pairwise, the tail added in order (NumPy's pairwise_sum)."""
    if len(vals) < 8:
        tot = vals[0]
        for v in vals[1:]:
            tot = tot + v
        return tot
    r = list(vals[:8])
    i = 8
    while i + 8 <= len(vals):
        r = [a + b for a, b in zip(r, vals[i:i + 8])]
        i += 8
    tot = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    for v in vals[i:]:
        tot = tot + v
    return tot


class MantisShrimp(UVAnimal):
    DEFAULTS = dict(lambdas=None, hsi_scale=0.25, panorama_scale=1.12,
                    bands=((320.0, 360.0), (360.0, 400.0), (400.0, 430.0), (430.0, 460.0), (460.0, 490.0), (490.0, 520.0), (520.0, 550.0),
                           (550.0, 580.0), (580.0, 610.0), (610.0, 680.0)),
                    red_kill=0.18, haze_strength=0.03, haze_tint=(0.92, 0.98, 1.0), pre_soft_sigma=0.25, unsharp_sigma=1.0,
                    unsharp_amount=0.32, evec_angle_deg=30.0, pol_linear_strength=0.55, pol_linear_gamma=1.2, pol_circular_strength=0.35,
                    orientation_mix=0.5, barcode_saturation=0.4, barcode_opacity=0.55, winner_take_most=0.35, scan_row_freq=26.0,
                    scan_row_gain=0.08, scan_soften=0.8, periph_blur_sigma=0.7, periph_radius=0.8, periph_softness=7.0)

    def __init__(self, **kw):
        super().__init__(**kw)
        self.bands = tuple((float(a), float(b)) for (a, b) in self.bands)                  # :108
        self.evec_angle = np.deg2rad(float(self.evec_angle_deg))                            # :113

    def _render(self, st):
        lz, ops = st.lz, st.ops
        N = len(self.bands)
        bt = st.bands(self.lambdas, self.bands, self.hsi_scale)                            # :163-171: ten band maps
        s_t = st.eval(st.normed_bands(bt))                                                 # S (H,W,N)
        p95 = lz.scalar(ops.percentile_frames(s_t, 0, 95.0, joint=True), 0)               # :199 percentile over the whole stack
        Sn = [L.clip(lz.plane(s_t, k) / (p95 + 1e-8), 0.0, 1.0) for k in range(N)]         # :199-200
        lut = barcode_lut(N)
        tot = _sum_numpy_order(Sn)                                                         # :206
        w = [s / (tot + 1e-8) for s in Sn]
        soft = []
        for c in range(3):                                                                 # :207 weights @ lut
            acc = w[0] * float(lut[0, c])
            for k in range(1, N):
                acc = acc + w[k] * float(lut[k, c])
            soft.append(acc)
        best, hard = Sn[0], [L._e(float(lut[0, c])) for c in range(3)]                      # :203-204, :208 argmax -> lut[max_idx]
        for k in range(1, N):
            gt = Sn[k] > best
            hard = [L.where(gt, float(lut[k, c]), hard[c]) for c in range(3)]
            best = L.maximum(best, Sn[k])
        wtm = self.winner_take_most
        bar = [(1.0 - wtm) * s_ + wtm * h_ for s_, h_ in zip(soft, hard)]                  # :210
        Yb = L.luma(bar)                                                                   # :213-214
        bar = [L.clip(Yb + (c - Yb) * (1.0 + self.barcode_saturation), 0.0, 1.0) for c in bar]
        broad = _sum_numpy_order(Sn) / float(N)                                            # :228 np.mean over the band axis
        bb_t = st.eval(bar + [broad])                                                      # (barcode rgb, broad)
        gx_t, gy_t = ops.sobel(bb_t[..., 3:4].contiguous())                                # :229-230
        theta = L.arctan2(lz.plane(gy_t, 0), lz.plane(gx_t, 0))
        mix = self.orientation_mix                                                         # :233-244
        c_mix = (1.0 - mix) * float(np.cos(2.0 * self.evec_angle)) + mix * L.cos(2.0 * theta)
        s_mix = (1.0 - mix) * float(np.sin(2.0 * self.evec_angle)) + mix * L.sin(2.0 * theta)
        align01 = L.clip(0.5 * (c_mix + 1.0), 0.0, 1.0) ** self.pol_linear_gamma
        align_circ = L.clip(0.5 * (s_mix + 1.0), 0.0, 1.0)
        pol_gain = 1.0 + self.pol_linear_strength * align01 + self.pol_circular_strength * align_circ
        r, g, b = st.baseline()
        render = [L.clip(r * (1.0 - self.red_kill), 0.0, 1.0), g, b]                       # :218
        if self.haze_strength > 0.0:                                                       # :219-221
            a = float(np.clip(self.haze_strength, 0.0, 1.0))
            veil = a * np.array(self.haze_tint, np.float32)
            render = [(1.0 - a) * c + float(veil[i]) for i, c in enumerate(render)]
        if self.pre_soft_sigma > 0.0:                                                      # :222-223
            render = lz.channels(st.blur(st.eval(render), self.pre_soft_sigma))
        if self.unsharp_sigma > 0.0 and self.unsharp_amount > 0.0:                         # :247-250
            t_img = st.eval(render)
            cur, blurred = lz.channels(t_img), lz.channels(st.blur(t_img, self.unsharp_sigma))
            amt = self.unsharp_amount * pol_gain
            render = [L.clip(c + amt * L.clip(c - q, -1.0, 1.0), 0.0, 1.0) for c, q in zip(cur, blurred)]
        op = self.barcode_opacity                                                          # :253
        render = [L.clip((1.0 - op) * c + op * lz.plane(bb_t, i), 0.0, 1.0) for i, c in enumerate(render)]
        if self.scan_row_gain != 0.0:                                                      # :256-263
            rg = lz.keyed(("scan_rows", self.scan_row_freq, self.scan_soften, self.scan_row_gain),
                          lambda: scan_row_gain(st.H, self.scan_row_freq, self.scan_soften, self.scan_row_gain), "row")
            render = [L.clip(c * rg, 0.0, 1.0) for c in render]
        if self.periph_blur_sigma > 0.0:                                                   # :266-275
            render = periph_mix(st, render, self.periph_blur_sigma, self.periph_softness, self.periph_radius)
        return render
