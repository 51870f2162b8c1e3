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


def split_compare_batch(original, modified, out=None, *, draw_seam: bool = True):
    """Device-side bulk of `make_split_frame` (reference renderers/video.py:198-245) for a whole batch: left half of
    `original`, right half of `modified`, optional 1-px white seam at W // 2.  Tensors are uint8 [N,H,W,3] on the same
    device (the frames never visit the host between `visualize_batch` and the encoder's D2H); strided row copies only.
    The corner labels are a few hundred pixels of cv2.putText: `draw_split_labels` adds them on the host copy, exactly
    as the reference draws them.  (The reference's INTER_AREA resize of a mismatching `modified` is not needed: every
    species returns frames of the input's shape.)"""
    assert original.shape == modified.shape and original.dim() == 4 and original.shape[3] == 3, "frames must be [N,H,W,3] of one shape"
    assert original.dtype == modified.dtype and original.device == modified.device
    if out is None:
        out = original.new_empty(original.shape)
    mid = original.shape[2] // 2
    out[:, :, :mid].copy_(original[:, :, :mid])
    out[:, :, mid:].copy_(modified[:, :, mid:])
    if draw_seam:
        out[:, :, mid:mid + 1] = 255
    return out


def draw_split_labels(frame: np.ndarray, left_label: str = "Original", right_label: str = "Transformed") -> np.ndarray:
    """The label step of `make_split_frame` (renderers/video.py:160-196 `_draw_label`, :241-244) on one host frame, in place."""
    import cv2
    H, W = frame.shape[:2]
    font = cv2.FONT_HERSHEY_SIMPLEX

    def label(text, org):
        scale, thickness, pad = max(0.5, min(1.2, H / 900.0)), 2, 8
        (tw, th), baseline = cv2.getTextSize(text, font, scale, thickness)
        x, y = org
        if x + tw + pad > W:                                   # keep inside the frame
            x = W - tw - pad
        if y - th - baseline - pad < 0:
            y = th + baseline + pad
        x0, y0 = max(x - pad, 0), max(y - th - baseline - pad, 0)
        x1, y1 = min(x + tw + pad, W - 1), min(y + baseline + pad, H - 1)
        overlay = frame.copy()
        cv2.rectangle(overlay, (x0, y0), (x1, y1), (0, 0, 0), thickness=-1)
        cv2.addWeighted(overlay, 0.6, frame, 0.4, 0, frame)
        cv2.putText(frame, text, (x, y), font, scale, (0, 0, 0), thickness + 2, cv2.LINE_AA)
        cv2.putText(frame, text, (x, y), font, scale, (255, 255, 255), thickness, cv2.LINE_AA)
    label(left_label, (10, 24))
    (tw, _), _ = cv2.getTextSize(right_label, font, max(0.45, min(1.2, H / 900.0)), 1)      # :242: its own scale / thickness
    label(right_label, (max(W - tw - 10, 10), 24))
    return frame
