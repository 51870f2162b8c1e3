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

    def _run_f32(self, eng, frames, quantize: bool):
        """Float / wide-integer frames (cat.py:24 accepts them): frames = CUDA float32 [N,H,W,3] holding
        image.astype(float32).  Returns (human_zoomed, cat_view) float32 tensors.
          human: center_zoom on the RAW values (cat_widevision_utils.py:11-29: crop + cv2.resize INTER_LINEAR, float
                 arithmetic for float frames -- no normalisation, no clipping)
          cat:   get_normalized_image -> binocular warp -> decode -> collapsed L/M-merge 3x3 -> Gaussian sigma 1 ->
                 clip -> encode -> clip (cat.py:82-103; the reference runs the tail in float64, here float32)."""
        from .._abi import AVB_F32_GAUSS, AVB_IMG_NORM_MAMMAL, check
        from ..imgops import get_imgops
        ops = get_imgops(eng)
        t = eng.torch
        n, H, W, _ = frames.shape
        scale = tables.cat_zoom_scale(self.CAMERA_HFOV_DEG, self.CAT_PER_EYE_HALF_FOV_DEG, self.CAT_TO_HUMAN_RATIO)
        human = frames if scale <= 1.0 else ops.resize(frames, (H, W), "linear", crop=tables.center_zoom_box(W, H, scale))
        srgb01 = ops.to_float01(frames, AVB_IMG_NORM_MAMMAL)
        if self.ENABLE_FOV_WARP:
            warp = eng.cached(("cat_warp", W) + self._geometry_key(), lambda: eng._dev(
                tables.cat_warp_device_table(W, self.CAMERA_HFOV_DEG, self.CAT_PER_EYE_HALF_FOV_DEG, self.CAT_OVERLAP_DEG)))
            warped = t.empty_like(srgb01)
            with t.cuda.device(eng.device):
                check(eng.lib.avb_cat_warp_f32(srgb01.data_ptr(), warped.data_ptr(), n, H, W, warp.data_ptr(), eng.stream_ptr()),
                      "avb_cat_warp_f32")
            eng.launches += 1
            srgb01 = warped
        cat, tmp = t.empty_like(srgb01), t.empty_like(srgb01)
        taps = tables.gaussian_taps(tables.gaussian_ksize(self.SIGMA), self.SIGMA)
        # values are already in [0,1] (max <= 1): avb_dichromat_f32's own normalisation is the identity
        eng.dichromat_f32(srgb01, cat, tmp, tables.cat_matrix(self.LM_ALPHA), AVB_F32_GAUSS, taps=taps, quantize=quantize)
        return human, cat

    def visualize(self, image: np.ndarray) -> Optional[tuple[np.ndarray, np.ndarray]]:
        assert isinstance(image, np.ndarray) and image.ndim == 3 and image.shape[2] == 3, "HxWx3 RGB"   # cat.py:24
        eng = get_engine()
        if image.dtype != np.uint8:
            return self._visualize_float(eng, image)
        human, cat = run_single(eng, image, lambda d_in, d_out: self._run(eng, d_in, d_out[0], d_out[1]), n_out=2)
        return human, cat

    def _visualize_float(self, eng, image: np.ndarray):
        """cat.py:105-112: integer dtypes get the zoomed frame as cv2.resize made it (rounded to the dtype) and
        cat*255+0.5 truncated; float dtypes a plain cast of both."""
        torch = eng.torch
        integer = np.issubdtype(image.dtype, np.integer)
        with torch.cuda.device(eng.device):
            pin = torch.empty((1,) + image.shape, dtype=torch.float32).pin_memory()
            pin[0].numpy()[...] = image                      # image.astype(float32): animal_utils.py:45
            dev = pin.to(eng.device, non_blocking=True)
            human, cat = self._run_f32(eng, dev, integer)
            h_host, c_host = human.cpu()[0].numpy(), cat.cpu()[0].numpy()
        if integer:
            h_host = np.rint(h_host)                         # cv2.resize on an integer Mat ends in saturate_cast: round half to even
        return h_host.astype(image.dtype), c_host.astype(image.dtype)
