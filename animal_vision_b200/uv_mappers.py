"""Device version of the one mapper of reference uv_mappers.py that the fused K3 kernel does not cover:
`map_uv_purple_yellow` (uv_mappers.py:67-87; HoneyBee never selects it, external callers may).  The five mappers
HoneyBee does select (`falsecolor`, `custom_matrix`, `opponent`, `uv_purple_yellow` = the *_soft variant, and
`falsecolor_uv_mixed`) run inside `avb_uv_map_u8` / `avb_uv_map_f32` together with the encode tail."""
from __future__ import annotations

import numpy as np

from . import lazy as L
from .engine import get_engine
from .imgops import get_imgops


def _s2l(v):
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / (1 + 0.055)) ** 2.4).astype(np.float32)


def map_uv_purple_yellow(U, eps: float = 1e-8):
    """UV-only visualisation between purple and yellow, linear RGB in [0,1].
    U: numpy HxW / HxWx1 (returns numpy HxWx3 float32) or CUDA float32 [N,H,W,1] (returns CUDA [N,H,W,3])."""
    is_np = isinstance(U, np.ndarray)
    if is_np:
        if U.ndim == 3 and U.shape[2] == 1:
            U = U[..., 0]
        elif U.ndim != 2:
            raise ValueError(f"U must be HxW or HxWx1, got {U.shape}")           # uv_mappers.py:69-72
        eng = get_engine()
        t = eng.torch
        dev = t.from_numpy(np.ascontiguousarray(U.astype(np.float32))[None, :, :, None]).to(eng.device)
    else:
        dev = U.contiguous()
        eng = get_engine(dev.device)
    n, H, W, _ = dev.shape
    ops, lz = get_imgops(eng), L.Lazy(eng, n, H, W)
    p99 = lz.scalar(ops.percentile_frames(dev, 0, 99.0), 0)
    u = L.clip(lz.plane(dev, 0) / L.maximum(p99, eps), 0.0, 1.0) ** 0.85            # :73-74
    c0 = _s2l(np.array([128, 0, 150], np.float32) / 255.0)                         # :76-84
    c1 = _s2l(np.array([255, 225, 60], np.float32) / 255.0)
    out = lz.eval([L.clip((1.0 - u) * float(c0[i]) + u * float(c1[i]), 0.0, 1.0) for i in range(3)])   # :85-87
    return out[0].cpu().numpy() if is_np else out
