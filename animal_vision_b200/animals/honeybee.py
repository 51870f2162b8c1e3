"""HoneyBee -- drop-in for reference animals/honeybee.py (constructor signature :47-66,
visualize :99-175).  The RGB->HSI step is the analytic 3-lobe model of
ml/classic_rgb_to_hsi/classic_rgb_to_hsi.py:47-82; the cube is never materialised."""
from __future__ import annotations

import hashlib
from typing import Callable, Literal, Optional, Tuple

import numpy as np

from .. import tables
from ..engine import get_engine
from .animal import Animal, run_single, run_single_float

_ADAPT = {None: 0, "white_patch": 1, "gray_world": 2}
_MAP = {"opponent": 0, "falsecolor": 1, "custom_matrix": 2, "uv_purple_yellow": 3, "falsecolor_uv_mixed": 4}   # AVB_MAP_*


class HoneyBee(Animal):
    def __init__(
        self,
        onnx_path: str = "./ml/MST_plus_plus/export/mst_plus_plus.onnx",
        hsi_band_centers_nm: Optional[np.ndarray] = None,
        illuminant: Optional[Callable[[np.ndarray], np.ndarray]] = None,
        adaptation: Optional[Literal["white_patch", "gray_world"]] = "white_patch",
        mapping_mode: Literal["falsecolor", "custom_matrix", "opponent", "uv_purple_yellow", "falsecolor_uv_mixed"] = "opponent",
        custom_matrix: Optional[np.ndarray] = None,
        blur_sigma_px: Optional[float] = 0.2,
        assume_hsi_is_reflectance: bool = True,
        *,
        hsi_downsample: bool = False,
        hsi_scale: float = 0.1,
        spectral_mode: Literal["collapsed", "bands"] = "collapsed",
    ):
        self.onnx_path = onnx_path
        self.adaptation = adaptation
        self.mapping_mode = mapping_mode
        self.custom_matrix = custom_matrix
        self.blur_sigma_px = float(blur_sigma_px or 0.0)
        self.assume_hsi_is_reflectance = assume_hsi_is_reflectance
        self.hsi_downsample = bool(hsi_downsample)
        self.hsi_scale = float(hsi_scale)
        self.spectral_mode = spectral_mode
        self.lambdas = (np.linspace(400.0, 700.0, 31, dtype=np.float32) if hsi_band_centers_nm is None
                        else np.asarray(hsi_band_centers_nm, dtype=np.float32))
        self.E = illuminant if illuminant is not None else tables.d65_like
        self.UV_curve, self.Blue_curve, self.Green_curve = tables.honeybee_curves(self.lambdas)
        if adaptation not in _ADAPT:
            raise ValueError(f"Unknown adaptation: {adaptation}")
        if mapping_mode not in _MAP:
            raise ValueError(f"Unknown mapping_mode: {mapping_mode}")                      # honeybee.py:163-164
        if mapping_mode == "custom_matrix":
            assert custom_matrix is not None and np.shape(custom_matrix) == (3, 3), \
                "Provide custom_matrix as 3\u00d73 for 'custom_matrix' mode."                    # honeybee.py:153-156
        # pixel-independent spectral tables, built once with the reference's own expressions
        sens = np.stack([self.UV_curve, self.Blue_curve, self.Green_curve])
        E = self.E(self.lambdas).astype(np.float32) if assume_hsi_is_reflectance else None
        self._band_tab, self._denom_eps = tables.uv_band_table(self.lambdas, sens, E)
        self._band_key = hashlib.sha1(np.ascontiguousarray(self._band_tab).tobytes()).hexdigest()   # cache by CONTENT, not id()
        self._M3 = tables.uv_collapsed_matrix(self.lambdas, sens, E)
        self._taps = tables.uv_blur_taps(self.blur_sigma_px)
        self._map_params = tables.uv_map_params(custom_matrix)

    def _bands_dev(self, eng):
        if self.spectral_mode != "bands":
            return None
        return eng.cached(("bee_bands", self._band_key), lambda: eng._dev(self._band_tab))

    def _downsampled(self) -> bool:
        return self.hsi_downsample and 0.05 <= self.hsi_scale < 1.0                       # honeybee.py:109

    def _fused_ok(self, frames) -> bool:
        """The fused uint8 kernel (K3) covers uint8 frames at full resolution with a blur of at most 5 taps."""
        return frames.dtype == get_engine(frames.device).torch.uint8 and not self._downsampled() and self._taps.size <= 5

    def _run(self, eng, frames, out, dbg=None, quantize=False):
        if self._fused_ok(frames) and out.dtype == frames.dtype:
            eng.uv_map(frames, out, self._M3, self._bands_dev(eng), self._denom_eps, _ADAPT[self.adaptation], self._taps,
                       _MAP[self.mapping_mode], self._map_params, 0.45, dbg)       # honeybee.py:161: alpha=0.45
            return
        self._run_planes(eng, frames, out, quantize)

    def _run_planes(self, eng, frames, out, quantize):
        """The float32 plane route (csrc/k6_imgops.cu + the f32 kernels of k3_uv.cu): float / wide-integer frames,
        hsi_downsample (uv_helpers.py:155-183: INTER_AREA down -> analytic HSI -> INTER_LINEAR up; the band
        projection is linear, so the three catch planes are up-sampled instead of the 31-band cube) and blur sigmas
        beyond five taps.  Step order as honeybee.py:105-173."""
        from .._abi import AVB_STAT_MAX, AVB_STAT_MEAN
        from ..imgops import get_imgops
        ops = get_imgops(eng)
        img01 = ops.to_float01(frames.contiguous())                                        # :106
        n, H, W, _ = img01.shape
        if self._downsampled():                                                            # :109-116
            hs, ws = tables.scaled_hw(H, W, self.hsi_scale)
            small = ops.resize(img01, (hs, ws), "area")
            ubg = ops.resize(ops.uv_catches(small, self._M3, self._bands_dev(eng), self._denom_eps), (H, W), "linear")
        else:
            ubg = ops.uv_catches(img01, self._M3, self._bands_dev(eng), self._denom_eps)   # :121-135
        if self.adaptation is not None:                                                    # :138-141
            ubg = ops.divide_channels(ubg, ops.stats(ubg), AVB_STAT_MAX if self.adaptation == "white_patch" else AVB_STAT_MEAN)
        if self._taps.size:                                                                # :144-147
            ubg = ops.blur_taps(ubg, self._taps)
        ops.uv_map(ubg, out, quantize, _MAP[self.mapping_mode], self._map_params, 0.45)    # :150-173

    def visualize_batch(self, frames, out=None):
        eng = get_engine(frames.device)
        if out is None:
            out = eng.torch.empty_like(frames)
        self._run(eng, frames, out)
        return frames, out

    def receptor_catches(self, frames):
        """Raw (U,B,G) catches as a float32 tensor [N,H,W,3] (honeybee.py:133-135) -- test hook."""
        eng = get_engine(frames.device)
        n, h, w, _ = frames.shape
        dbg = eng.torch.empty((n, h, w, 3), dtype=eng.torch.float32, device=eng.device)
        self._run(eng, frames, eng.torch.empty_like(frames), dbg)
        return dbg

    def visualize(self, image: np.ndarray) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        assert isinstance(image, np.ndarray), "Input must be a numpy ndarray."          # honeybee.py:102-103
        assert image.ndim == 3 and image.shape[2] == 3, "Input must be HxWx3 RGB."
        eng = get_engine()
        if image.dtype != np.uint8:         # float in => float out, wider integers => x*255+0.5 truncated (honeybee.py:170-173)
            out = run_single_float(eng, image, lambda d_in, d_out, d_tmp, q: self._run_planes(eng, d_in, d_out, q))
            return image, out
        (out,) = run_single(eng, image, lambda d_in, d_out: self._run(eng, d_in, d_out[0]))
        return image, out
