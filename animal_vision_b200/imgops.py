"""Torch-tensor wrappers of the generic float32 image operators (csrc/k6_imgops.cu, include/avb200.h "K6").

Every function takes / returns CUDA float32 tensors packed [n, H, W, C] and enqueues on the current stream of
the engine's device; the arithmetic happens in libavb200.so (torch only owns the memory).  These are the
building blocks of the float-frame routes of Cat and HoneyBee and of the UV species (uv_helpers.py steps).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import tables
from ._abi import AVB_IMG_NORM_UV, AVB_STAT_MAX, AvbError, check
from .engine import Engine, _fptr


class ImgOps:
    def __init__(self, eng: Engine):
        self.eng = eng
        self.t = eng.torch
        self.lib = eng.lib

    # ------------------------------------------------------------------ helpers
    def _check(self, x, name="image"):
        t = self.t
        if not (isinstance(x, t.Tensor) and x.is_cuda and x.dtype == t.float32 and x.dim() == 4 and x.is_contiguous()):
            raise AvbError(f"{name}: expected a contiguous CUDA float32 tensor [N,H,W,C]")
        if x.device != self.eng.device:
            raise AvbError(f"{name} lives on {x.device}, engine on {self.eng.device}")
        return tuple(x.shape)

    def _scratch(self, nbytes: int):
        key = ("img_scratch", self.eng.stream_ptr())
        buf = self.eng._cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = self.eng._cache[key] = self.t.empty(max(int(nbytes), 1 << 16), dtype=self.t.uint8, device=self.eng.device)
        return buf

    def _taps_dev(self, key, build):
        return self.eng.cached(key, lambda: tuple(self.eng._dev(a) for a in build()))

    # ------------------------------------------------------------------ operators
    def to_float01(self, frames, mode: int = AVB_IMG_NORM_UV):
        """uint8 or float32 CUDA [N,H,W,3] -> float32 (uv_helpers.py:15-23 / animal_utils.py:41-50 by `mode`)."""
        t = self.t
        if not (frames.is_cuda and frames.dim() == 4 and frames.dtype in (t.uint8, t.float32) and frames.is_contiguous()):
            raise AvbError("to_float01: expected a contiguous CUDA uint8 / float32 tensor [N,H,W,C]")
        n = frames.shape[0]
        out = t.empty(tuple(frames.shape), dtype=t.float32, device=self.eng.device)
        with t.cuda.device(self.eng.device):
            rc = self.lib.avb_img_to_float01(frames.data_ptr(), int(frames.dtype == t.uint8), out.data_ptr(), n,
                                             frames[0].numel(), int(mode), self._scratch(4 * n).data_ptr(), self.eng.stream_ptr())
        check(rc, "avb_img_to_float01")
        self.eng.launches += 2
        return out

    def _resample_axis(self, x, out_len: int, axis: int, interp: str, src_len: int, src_off: int = 0):
        """One axis of cv2.resize; `x` may be read as a crop [src_off, src_off + src_len) along the axis."""
        t = self.t
        n, H, W, Cn = x.shape
        idx, w = self._taps_dev(("resize", interp, src_len, out_len, axis),
                                lambda: tables.resize_taps(src_len, out_len, interp, vertical=(axis == 1)))
        Hout, Wout = (H, out_len) if axis == 0 else (out_len, W)
        out = t.empty((n, Hout, Wout, Cn), dtype=t.float32, device=self.eng.device)
        base = x.data_ptr() + 4 * (src_off * Cn if axis == 0 else src_off * W * Cn)
        with t.cuda.device(self.eng.device):
            rc = self.lib.avb_img_resample(base, out.data_ptr(), n, Hout, Wout, Cn, axis, H * W * Cn, W * Cn,
                                           idx.data_ptr(), w.data_ptr(), int(idx.shape[1]), self.eng.stream_ptr())
        check(rc, "avb_img_resample")
        self.eng.launches += 1
        return out

    def resize(self, x, out_hw: Tuple[int, int], interp: str, crop=None):
        """cv2.resize(x, (W_out, H_out), interpolation) on float32 frames, horizontal pass first.
        crop = (x0, y0, cw, ch): resize that sub-rectangle (cat_widevision_utils.py:19-26 center_zoom)."""
        n, H, W, Cn = self._check(x)
        x0, y0, cw, ch = crop if crop is not None else (0, 0, W, H)
        Ho, Wo = int(out_hw[0]), int(out_hw[1])
        y = x
        if not (cw == W and x0 == 0 and Wo == W):                        # an axis whose length is unchanged is the identity
            y = self._resample_axis(y, Wo, 0, interp, cw, x0)          # [n, H, Wo, C]: all rows, the crop's columns
        if not (ch == H and y0 == 0 and Ho == H):
            y = self._resample_axis(y, Ho, 1, interp, ch, y0)
        return y

    def panorama(self, x, scale_x: float):
        """uv_helpers.py:84-99 panorama_warp: cv2.resize(INTER_CUBIC) to round(W*scale_x) columns, centre crop back to W.
        Only the W surviving columns are computed (the tap tables are sliced, not the image)."""
        n, H, W, Cn = self._check(x)
        geo = tables.panorama_geometry(W, float(scale_x))
        if geo is None:
            return x
        newW, start = geo

        def build():
            idx, w = tables.resize_taps(W, newW, "cubic")
            return np.ascontiguousarray(idx[start:start + W]), np.ascontiguousarray(w[start:start + W])
        idx, w = self._taps_dev(("panorama", W, newW, start), build)
        out = self.t.empty_like(x)
        with self.t.cuda.device(self.eng.device):
            rc = self.lib.avb_img_resample(x.data_ptr(), out.data_ptr(), n, H, W, Cn, 0, H * W * Cn, W * Cn,
                                           idx.data_ptr(), w.data_ptr(), int(idx.shape[1]), self.eng.stream_ptr())
        check(rc, "avb_img_resample")
        self.eng.launches += 1
        return out

    def sobel(self, x):
        """cv2.Sobel(x, CV_32F, 1, 0, ksize=3) and (0, 1) with BORDER_REFLECT_101 (mantis_shrimp.py:124-125 and siblings):
        the separable pairs ([-1,0,1] along x, [1,2,1] along y) and ([1,2,1], [-1,0,1])."""
        d = np.array([-1.0, 0.0, 1.0], np.float32)
        sm = np.array([1.0, 2.0, 1.0], np.float32)
        return self.blur_taps(x, d, sm), self.blur_taps(x, sm, d)

    def remap(self, x, map_x: np.ndarray, map_y: np.ndarray, key=None):
        """cv2.remap(x, map_x, map_y, INTER_LINEAR, BORDER_REFLECT101) with host-built float32 maps (anableps.py:224-237)."""
        n, H, W, Cn = self._check(x)
        mx = np.ascontiguousarray(map_x, np.float32)
        my = np.ascontiguousarray(map_y, np.float32)
        assert mx.shape == (H, W) and my.shape == (H, W)
        import hashlib
        kx = ("remap_x", key) if key is not None else ("remap", hashlib.sha1(mx.tobytes()).hexdigest())
        ky = ("remap_y", key) if key is not None else ("remap", hashlib.sha1(my.tobytes()).hexdigest())
        dx = self.eng.cached(kx, lambda: self.eng._dev(mx))
        dy = self.eng.cached(ky, lambda: self.eng._dev(my))
        out = self.t.empty_like(x)
        with self.t.cuda.device(self.eng.device):
            rc = self.lib.avb_img_remap(x.data_ptr(), out.data_ptr(), n, H, W, Cn, dx.data_ptr(), dy.data_ptr(), self.eng.stream_ptr())
        check(rc, "avb_img_remap")
        self.eng.launches += 1
        return out

    def percentile_frames(self, x, channel: int, q: float, joint: bool = False):
        """numpy.percentile(plane, q) of one channel of every frame -> float32 tensor [n, 1] on the device.
        joint = True: the percentile over ALL channels of a frame together (np.percentile of an (H,W,N) stack)."""
        n, H, W, Cn = self._check(x)
        if joint:
            flat = x.view(n, 1, H * W * Cn, 1)
            return self.percentile(flat, [(f, 0, q) for f in range(n)]).view(n, 1)
        return self.percentile(x, [(f, channel, q) for f in range(n)]).view(n, 1)

    def blur_taps(self, x, taps_x: np.ndarray, taps_y: np.ndarray = None):
        """Separable correlation, BORDER_REFLECT_101, rows first (cv2.GaussianBlur / sepFilter2D)."""
        t = self.t
        n, H, W, Cn = self._check(x)
        taps_y = taps_x if taps_y is None else taps_y
        tx = np.ascontiguousarray(taps_x, np.float32)
        ty = np.ascontiguousarray(taps_y, np.float32)
        dx = self.eng.cached(("taps", tx.tobytes()), lambda: self.eng._dev(tx))
        dy = self.eng.cached(("taps", ty.tobytes()), lambda: self.eng._dev(ty))
        out, tmp = t.empty_like(x), t.empty_like(x)
        with t.cuda.device(self.eng.device):
            rc = self.lib.avb_img_blur(x.data_ptr(), out.data_ptr(), tmp.data_ptr(), n, H, W, Cn, dx.data_ptr(), int(tx.size),
                                       dy.data_ptr(), int(ty.size), self.eng.stream_ptr())
        check(rc, "avb_img_blur")
        self.eng.launches += 2
        return out

    def gaussian_blur(self, x, sigma: float):
        """uv_helpers.py:67-73 gaussian_blur: k = 2*ceil(3*sigma)+1 taps, identity for sigma <= 0."""
        if sigma <= 0:
            return x
        return self.blur_taps(x, tables.uv_blur_taps(float(sigma)))

    def stats(self, x):
        """[n, C, 4] = per frame and channel (min, max, mean, 0)."""
        t = self.t
        n, H, W, Cn = self._check(x)
        out = t.empty((n, Cn, 4), dtype=t.float32, device=self.eng.device)
        with t.cuda.device(self.eng.device):
            rc = self.lib.avb_img_stats(x.data_ptr(), n, H * W, Cn, out.data_ptr(), self._scratch(16 * n * Cn).data_ptr(), self.eng.stream_ptr())
        check(rc, "avb_img_stats")
        self.eng.launches += 3
        return out

    def divide_channels(self, x, stats, which: int = AVB_STAT_MAX, eps: float = 1e-8):
        t = self.t
        n, H, W, Cn = self._check(x)
        out = t.empty_like(x)
        with t.cuda.device(self.eng.device):
            rc = self.lib.avb_img_divide_channels(x.data_ptr(), out.data_ptr(), n, H * W, Cn, stats.data_ptr(), int(which), float(eps),
                                                  self.eng.stream_ptr())
        check(rc, "avb_img_divide_channels")
        self.eng.launches += 1
        return out

    def percentile(self, x, requests: Sequence[Tuple[int, int, float]]):
        """numpy.percentile of single channels: requests = [(frame, channel, q), ...] -> float32 tensor [len(requests)]."""
        t = self.t
        n, H, W, Cn = self._check(x)
        nreq = len(requests)
        offs = np.array([f * H * W * Cn + c for f, c, _ in requests], np.int64)
        qs = np.array([q for _, _, q in requests], np.float64)
        out = t.empty(nreq, dtype=t.float32, device=self.eng.device)
        need = int(self.lib.avb_img_percentile_scratch_bytes(min(nreq, 16)))
        with t.cuda.device(self.eng.device):
            rc = self.lib.avb_img_percentile(x.data_ptr(), H * W, Cn, offs.ctypes.data_as(C.c_void_p), qs.ctypes.data_as(C.c_void_p),
                                             nreq, out.data_ptr(), self._scratch(need).data_ptr(), self.eng.stream_ptr())
        check(rc, "avb_img_percentile")
        self.eng.launches += 9 * ((nreq + 15) // 16)
        return out

    # ------------------------------------------------------------------ UV plane route
    def uv_catches(self, img01, M3: np.ndarray, bands_dev, denom_eps: float):
        t = self.t
        n, H, W, Cn = self._check(img01)
        assert Cn == 3
        out = t.empty_like(img01)
        M3 = np.ascontiguousarray(M3, np.float32)
        with t.cuda.device(self.eng.device):
            rc = self.lib.avb_uv_catches_f32(img01.data_ptr(), out.data_ptr(), n * H * W, _fptr(M3),
                                             None if bands_dev is None else bands_dev.data_ptr(),
                                             0 if bands_dev is None else int(bands_dev.shape[0]), float(denom_eps), self.eng.stream_ptr())
        check(rc, "avb_uv_catches_f32")
        self.eng.launches += 1
        return out

    def uv_map(self, ubg, out, quantize: bool, map_mode: int, map_params, mix_alpha: float):
        """(U,B,G) planes -> mapper -> encode into `out`: uint8 [N,H,W,3] (strided rows allowed) or float32 packed."""
        t = self.t
        n, H, W, Cn = self._check(ubg)
        assert Cn == 3
        out_f32 = out.dtype == t.float32
        if out_f32:
            assert out.is_contiguous() and tuple(out.shape) == (n, H, W, 3)
            ofs = ors = 0
        else:
            _, _, _, ofs, ors = self.eng.check_frames(out, "out")
        need = int(self.lib.avb_uv_workspace_bytes(n, H, W, int(map_mode)))
        ws_key = ("uv_ws", self.eng.stream_ptr())
        ws = self.eng._cache.get(ws_key)
        if ws is None or ws.numel() < need:
            self.eng._cache[ws_key] = None
            ws = self.eng._cache[ws_key] = t.empty(need, dtype=t.uint8, device=self.eng.device)
        mp = None if map_params is None else np.ascontiguousarray(map_params, np.float32)
        with t.cuda.device(self.eng.device):
            rc = self.lib.avb_uv_map_f32(ubg.data_ptr(), out.data_ptr(), int(out_f32), int(bool(quantize)), n, H, W, ofs, ors,
                                         self.eng.enc.data_ptr(), int(map_mode), None if mp is None else _fptr(mp), float(mix_alpha),
                                         ws.data_ptr(), self.eng.stream_ptr())
        check(rc, "avb_uv_map_f32")
        self.eng.launches += 1 + (6 if map_mode != 2 else 0)
        return out


def get_imgops(eng: Engine) -> ImgOps:
    ops = getattr(eng, "_imgops", None)
    if ops is None:
        ops = eng._imgops = ImgOps(eng)
    return ops
