"""HoneyBee -- drop-in for reference animals/honeybee.py (constructor signature :47-66,
visualize :99-175).  The RGB->HSI step is the analytic 3-lobe model of
ml/classic_rgb_to_hsi/classic_rgb_to_hsi.py:47-82; the cube is never materialised."""
from __future__ import annotations

import hashlib
from typing import Callable, Literal, Optional, Tuple

import numpy as np

from .. import tables
from ..engine import get_engine
from .animal import Animal, run_single

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
        if self.hsi_downsample:
            raise NotImplementedError("hsi_downsample=True is not implemented on the GPU path (the default is False)")
        # pixel-independent spectral tables, built once with the reference's own expressions
        sens = np.stack([self.UV_curve, self.Blue_curve, self.Green_curve])
        E = self.E(self.lambdas).astype(np.float32) if assume_hsi_is_reflectance else None
        self._band_tab, self._denom_eps = tables.uv_band_table(self.lambdas, sens, E)
        self._band_key = hashlib.sha1(np.ascontiguousarray(self._band_tab).tobytes()).hexdigest()   # cache by CONTENT, not id()
        self._M3 = tables.uv_collapsed_matrix(self.lambdas, sens, E)
        self._taps = tables.uv_blur_taps(self.blur_sigma_px)
        self._map_params = tables.uv_map_params(custom_matrix)
        if self._taps.size > 5:
            raise NotImplementedError("blur_sigma_px > 2/3 (ksize > 5) is not implemented on the GPU path")

    def _run(self, eng, frames, out, dbg=None):
        bands = None
        if self.spectral_mode == "bands":
            bands = eng.cached(("bee_bands", self._band_key), lambda: eng._dev(self._band_tab))
        eng.uv_map(frames, out, self._M3, bands, self._denom_eps, _ADAPT[self.adaptation], self._taps,
                   _MAP[self.mapping_mode], self._map_params, 0.45, dbg)       # honeybee.py:161: alpha=0.45

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
        (out,) = run_single(eng, image, lambda d_in, d_out: self._run(eng, d_in, d_out[0]))
        return image, out
