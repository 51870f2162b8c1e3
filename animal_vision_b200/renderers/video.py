"""Batched frame feed that sits beside the reference's `VideoRenderer.get_image()`
(renderers/video.py:82-96).  Decode / encode / label drawing stay on the CPU in the reference's
own renderer; this only gathers its frames into pinned batches for `visualize_batch`."""
from __future__ import annotations

from typing import Optional

import numpy as np


class BatchFeed:
    """Wraps any object with `get_image() -> Optional[np.ndarray HxWx3 uint8]` (the reference's
    VideoRenderer / WebcamRenderer / ImageRenderer all have it)."""

    def __init__(self, source, batch: int = 8, pinned: bool = True):
        self.source = source
        self.batch = int(batch)
        self.pinned = pinned
        self._buf = None

    def get_batch(self, n: Optional[int] = None):
        """Up to n frames as a uint8 array/tensor [k,H,W,3] (k <= n; None at end of stream)."""
        n = n or self.batch
        first = self.source.get_image()
        if first is None:
            return None
        assert first.ndim == 3 and first.shape[2] == 3 and first.dtype == np.uint8
        if self._buf is None or tuple(self._buf.shape[1:]) != first.shape or self._buf.shape[0] < n:
            if self.pinned:
                import torch
                self._buf = torch.empty((n,) + first.shape, dtype=torch.uint8).pin_memory()
            else:
                self._buf = np.empty((n,) + first.shape, np.uint8)
        view = self._buf.numpy() if self.pinned else self._buf
        view[0] = first
        k = 1
        while k < n:
            f = self.source.get_image()
            if f is None:
                break
            view[k] = f
            k += 1
        return self._buf[:k]
