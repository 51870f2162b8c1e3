"""Species plugin base class -- same contract as reference animals/animal.py:5-7.

`visualize(image: np.ndarray HxWx3) -> Optional[(baseline, out)]` is the reference's entry point
(called by main.py:41-48, :60-71, :82-93 and utils.py:141-149).  `visualize_batch` is the
device-resident path added by this implementation: uint8 CUDA tensor [N,H,W,3] in, tensors out.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


class Animal:
    N_OUTPUTS = 1          # device outputs of visualize_batch (Cat: 2, the panorama species: baseline + view)

    def visualize(self, image: np.ndarray) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        pass

    def visualize_batch(self, frames, out=None):
        """frames: torch.uint8 CUDA tensor [N,H,W,3].  Returns (baseline, out) tensors."""
        raise NotImplementedError


def is_frame(image) -> bool:
    """animals/animal_utils.py:21-39 check_input_image."""
    return (isinstance(image, np.ndarray) and image.ndim == 3 and image.shape[2] == 3
            and np.issubdtype(image.dtype, np.number))


def run_single_float(engine, image: np.ndarray, fn):
    """Float (or non-uint8 integer) frame through the float32 device path: the reference's own first step
    is `image.astype(np.float32)` (animals/animal_utils.py:45), done here on the way into pinned memory;
    `fn(dev_in, dev_out, dev_tmp, quantize)` runs on the current stream.  Returns one HxWx3 array of the
    caller's dtype (dog.py:56-59: integer dtypes get x*255+0.5 truncated, floats a plain cast)."""
    torch = engine.torch
    integer = np.issubdtype(image.dtype, np.integer)
    with torch.cuda.device(engine.device):
        pin = torch.empty((1,) + image.shape, dtype=torch.float32).pin_memory()
        pin[0].numpy()[...] = image                      # float32 conversion, as astype(np.float32)
        dev_in = pin.to(engine.device, non_blocking=True)
        dev_out, dev_tmp = torch.empty_like(dev_in), torch.empty_like(dev_in)
        fn(dev_in, dev_out, dev_tmp, integer)
        pin.copy_(dev_out, non_blocking=True)
        torch.cuda.current_stream(engine.device).synchronize()
        return pin[0].numpy().astype(image.dtype)


def run_single(engine, image: np.ndarray, fn, n_out: int = 1):
    """NumPy-in / NumPy-out shim: pinned H2D copy, `fn(dev_in, dev_outs)` on the current stream,
    pinned D2H copy, one synchronise.  Returns a list of fresh HxWx3 uint8 arrays."""
    torch = engine.torch
    if image.dtype != np.uint8:
        raise NotImplementedError(
            "animal_vision_b200 runs the uint8 frame path on the GPU; float frames are not implemented "
            "(there is deliberately no CPU fallback)")
    with torch.cuda.device(engine.device):
        pin_in, dev_in, dev_out, pin_out = engine.staging(image.shape, n_out)
        pin_in[0].numpy()[...] = image
        dev_in.copy_(pin_in, non_blocking=True)
        fn(dev_in, dev_out)
        for d, p in zip(dev_out, pin_out):
            p.copy_(d, non_blocking=True)
        torch.cuda.current_stream(engine.device).synchronize()
        return [p[0].numpy().copy() for p in pin_out]
