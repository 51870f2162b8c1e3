"""Cat -- drop-in for reference animals/cat.py (the runnable `Tina-animals` side of its unresolved
merge conflict, cat.py:73-114): returns (human_zoomed, cat_view), both new arrays."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import tables
from ..engine import get_engine
from .animal import Animal, run_single


class Cat(Animal):
    N_OUTPUTS = 2
    # geometry parameters, cat.py:17-21
    CAMERA_HFOV_DEG = 100.0
    CAT_PER_EYE_HALF_FOV_DEG = 105.0
    CAT_OVERLAP_DEG = 40.0
    CAT_TO_HUMAN_RATIO = 1.30
    ENABLE_FOV_WARP = True
    LM_ALPHA = 0.5          # cat.py:98
    SIGMA = 1.0             # cat.py:102

    def _geometry_key(self):
        return (self.CAMERA_HFOV_DEG, self.CAT_PER_EYE_HALF_FOV_DEG, self.CAT_OVERLAP_DEG, self.CAT_TO_HUMAN_RATIO)

    def _run(self, eng, frames, out_human, out_cat):
        _, H, W, _ = frames.shape
        warp = None                                          # ENABLE_FOV_WARP = False: cat.py:84 skips the warp
        if self.ENABLE_FOV_WARP:
            warp = eng.cached(("cat_warp", W) + self._geometry_key(), lambda: eng._dev(
                tables.cat_warp_device_table(W, self.CAMERA_HFOV_DEG, self.CAT_PER_EYE_HALF_FOV_DEG, self.CAT_OVERLAP_DEG)))
        scale = tables.cat_zoom_scale(self.CAMERA_HFOV_DEG, self.CAT_PER_EYE_HALF_FOV_DEG, self.CAT_TO_HUMAN_RATIO)
        zoom = eng.cached(("cat_zoom", W, H) + self._geometry_key(),
                          lambda: eng._dev(tables.center_zoom_tables(W, H, scale)))
        taps = tables.gaussian_taps(tables.gaussian_ksize(self.SIGMA), self.SIGMA)
        eng.cat(frames, out_human, out_cat, tables.cat_matrix(self.LM_ALPHA), taps, warp, zoom)

    def visualize_batch(self, frames, out=None):
        """frames: CUDA uint8 [N,H,W,3] -> (human_zoomed, cat_view) tensors; `out` = optional pair."""
        eng = get_engine(frames.device)
        human, cat = out if out is not None else (eng.torch.empty_like(frames), eng.torch.empty_like(frames))
        self._run(eng, frames, human, cat)
        return human, cat

    def visualize(self, image: np.ndarray) -> Optional[tuple[np.ndarray, np.ndarray]]:
        assert isinstance(image, np.ndarray) and image.ndim == 3 and image.shape[2] == 3, "HxWx3 RGB"   # cat.py:24
        eng = get_engine()
        human, cat = run_single(eng, image, lambda d_in, d_out: self._run(eng, d_in, d_out[0], d_out[1]), n_out=2)
        return human, cat
