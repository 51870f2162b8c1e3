"""MST++ RGB -> hyperspectral inference on the GPU -- host side of K4.

Mirrors the reference's two entry points:
  * `MSTPlusPlus.forward(x)`            <-> MST_Plus_Plus.forward (architecture/MST_Plus_Plus.py:279-293):
                                            NCHW float32 in [0,1] -> NCHW float32, 31 bands
  * `MSTPlusPlus.predict_rgb_to_hsi(im)` <-> predict_rgb_to_hsi_torch (predict_torch.py:249-310):
                                            HWC uint8/float frame -> HWC float32 cube (centred reflect pad
                                            to a multiple of 16, crop), without fp16 autocast / OOM tiling
Weights are handed over as the reference's own state_dict (the names of MST_Plus_Plus().state_dict());
the reference ships none (model_zoo is git-ignored), so callers load a checkpoint themselves exactly
as architecture/__init__.py:36-40 does and pass the dict.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Mapping

import numpy as np

from . import tables
from ._abi import AvbError, check
from .engine import get_engine

N_FEAT = 31
N_PARAMS = 1619625


def param_order() -> "OrderedDict[str, tuple]":
    """Names and shapes of MST_Plus_Plus().state_dict() in registration order (227 tensors)."""
    sh: "OrderedDict[str, tuple]" = OrderedDict()

    def msab(prefix, dim, heads):
        p = f"{prefix}.blocks.0."
        sh[p + "0.rescale"] = (heads, 1, 1)
        for n in ("to_q", "to_k", "to_v"):
            sh[p + f"0.{n}.weight"] = (N_FEAT * heads, dim)
        sh[p + "0.proj.weight"] = (dim, N_FEAT * heads)
        sh[p + "0.proj.bias"] = (dim,)
        sh[p + "0.pos_emb.0.weight"] = (dim, 1, 3, 3)
        sh[p + "0.pos_emb.2.weight"] = (dim, 1, 3, 3)
        sh[p + "1.fn.net.0.weight"] = (dim * 4, dim, 1, 1)
        sh[p + "1.fn.net.2.weight"] = (dim * 4, 1, 3, 3)
        sh[p + "1.fn.net.4.weight"] = (dim, dim * 4, 1, 1)
        sh[p + "1.norm.weight"] = (dim,)
        sh[p + "1.norm.bias"] = (dim,)

    sh["conv_in.weight"] = (N_FEAT, 3, 3, 3)
    for s in range(3):
        b = f"body.{s}."
        sh[b + "embedding.weight"] = (N_FEAT, N_FEAT, 3, 3)
        dim = N_FEAT
        for i in range(2):
            msab(b + f"encoder_layers.{i}.0", dim, dim // N_FEAT)
            sh[b + f"encoder_layers.{i}.1.weight"] = (dim * 2, dim, 4, 4)
            dim *= 2
        msab(b + "bottleneck", dim, dim // N_FEAT)
        for i in range(2):
            sh[b + f"decoder_layers.{i}.0.weight"] = (dim, dim // 2, 2, 2)
            sh[b + f"decoder_layers.{i}.0.bias"] = (dim // 2,)
            sh[b + f"decoder_layers.{i}.1.weight"] = (dim // 2, dim, 1, 1)
            msab(b + f"decoder_layers.{i}.2", dim // 2, (dim // 2) // N_FEAT)
            dim //= 2
        sh[b + "mapping.weight"] = (N_FEAT, N_FEAT, 3, 3)
    sh["conv_out.weight"] = (N_FEAT, N_FEAT, 3, 3)
    return sh


def flatten_state_dict(state_dict: Mapping[str, object]) -> np.ndarray:
    """float32 vector of every parameter in registration order; tolerates the `module.` prefix of
    DataParallel checkpoints like the reference loader (architecture/__init__.py:38-39)."""
    sd = {k.replace("module.", "", 1) if k.startswith("module.") else k: v for k, v in state_dict.items()}
    parts = []
    for name, shape in param_order().items():
        if name not in sd:
            raise KeyError(f"state_dict is missing {name}")
        t = sd[name]
        a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(a.shape)}")
        parts.append(np.ascontiguousarray(a, np.float32).ravel())
    flat = np.concatenate(parts)
    assert flat.size == N_PARAMS
    return flat


def synthetic_state_dict(seed: int = 0) -> "OrderedDict[str, object]":
    """Deterministic synthetic weights in the reference's state_dict naming -- the reference ships no
    checkpoint (`model_zoo` is git-ignored), so benchmarks and smoke tests run on these.  Fan-in scaled
    normals for matrices / kernels; LayerNorm / rescale / bias values moved OFF their init values (1 / 1 / 0)
    so that every parameter matters.  Same generator and draw order as the test oracle's make_weights
    (tests/test_mstpp_host.py holds the two equal)."""
    import torch
    g = torch.Generator().manual_seed(int(seed))
    sd: "OrderedDict[str, object]" = OrderedDict()
    for name, shape in param_order().items():
        if name.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("rescale"):
            t = 1.0 + 0.25 * torch.rand(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if ".decoder_layers." in name and name.endswith(".0.weight"):
                fan_in = shape[0]                       # ConvTranspose2d k=2,s=2: one tap per output
            t = torch.randn(shape, generator=g) * (0.7 / fan_in ** 0.5)   # gain 0.7: |y| stays O(1)
        sd[name] = t.float()
    return sd


class MSTPlusPlus:
    def __init__(self, state_dict: Mapping[str, object], device=None):
        self.eng = get_engine(device)
        flat = flatten_state_dict(state_dict)
        h = C.c_void_p()
        with self.eng.torch.cuda.device(self.eng.device):
            rc = self.eng.lib.avb_mstpp_create(flat.ctypes.data_as(C.c_void_p), flat.size, C.byref(h))
        check(rc, "avb_mstpp_create")
        self._h = h
        self._ws = {}              # workspace per CUDA stream: forwards on different streams may overlap
        self._streams = []

    def close(self):
        if getattr(self, "_h", None):
            self.eng.lib.avb_mstpp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run(self, frames, pad_multiple: int, centred: bool, out=None):
        """frames: CUDA tensor [N,H,W,3] float32 (in [0,1]) or uint8 -> float32 [N,H,W,31]."""
        t = self.eng.torch
        if not (isinstance(frames, t.Tensor) and frames.is_cuda and frames.dim() == 4 and frames.shape[3] == 3
                and frames.dtype in (t.float32, t.uint8)):
            raise AvbError("MSTPlusPlus: expected a CUDA tensor [N,H,W,3] of float32 or uint8")
        if frames.device != self.eng.device:
            raise AvbError(f"MSTPlusPlus: frames live on {frames.device}, the model on {self.eng.device}")
        with t.cuda.device(self.eng.device):           # launches act on the CURRENT device
            return self._run_on_device(frames, pad_multiple, centred, out)

    def _run_on_device(self, frames, pad_multiple: int, centred: bool, out=None):
        t = self.eng.torch
        frames = frames.contiguous()
        n, h, w, _ = frames.shape
        need = int(self.eng.lib.avb_mstpp_workspace_bytes(n, h, w, pad_multiple, int(centred)))
        if need <= 0:
            raise AvbError("avb_mstpp_workspace_bytes: bad geometry")
        key = self.eng.stream_ptr()
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            self._ws[key] = None
            ws = self._ws[key] = t.empty(need, dtype=t.uint8, device=self.eng.device)
        if out is None:
            out = t.empty((n, h, w, N_FEAT), dtype=t.float32, device=self.eng.device)
        assert out.is_contiguous() and tuple(out.shape) == (n, h, w, N_FEAT) and out.dtype == t.float32
        rc = self.eng.lib.avb_mstpp_forward(self._h, frames.data_ptr(), int(frames.dtype == t.uint8), out.data_ptr(),
                                            n, h, w, pad_multiple, int(centred), ws.data_ptr(), self.eng.stream_ptr())
        check(rc, "avb_mstpp_forward")
        return out

    def forward_nhwc_streams(self, frames, streams: int = 2):
        """forward_nhwc with the batch cut into `streams` parts that run concurrently on their own CUDA
        streams (fork / join on the current stream).  Patches are independent -- the attention statistics
        are per patch -- and the coarse levels of the network are latency bound, so overlapping forwards
        raises the throughput of a batch."""
        t = self.eng.torch
        n = frames.shape[0]
        parts = max(1, min(int(streams), n))
        if parts == 1:
            return self._run(frames, 8, False)
        frames = frames.contiguous()
        while len(self._streams) < parts:
            self._streams.append(t.cuda.Stream(device=self.eng.device))
        out = t.empty((n,) + tuple(frames.shape[1:3]) + (N_FEAT,), dtype=t.float32, device=self.eng.device)
        main = t.cuda.current_stream(self.eng.device)
        fork = t.cuda.Event()
        fork.record(main)
        for k in range(parts):
            a, b = n * k // parts, n * (k + 1) // parts
            st = self._streams[k]
            st.wait_event(fork)
            with t.cuda.stream(st):
                self._run(frames[a:b], 8, False, out=out[a:b])
                done = t.cuda.Event()
                done.record(st)
            main.wait_event(done)
        return out

    def forward_nhwc(self, frames):
        """Model-direct semantics on channels-last frames: [N,H,W,3] -> [N,H,W,31]."""
        return self._run(frames, 8, False)

    def forward_bands(self, frames, weights: np.ndarray, *, want_cube: bool = False, pad_multiple: int = 8, centred: bool = False):
        """Forward with the band projection fused on the network output (avb_mstpp_forward_bands): frames [N,H,W,3]
        float32 / uint8 CUDA, weights [R,31] (e.g. tables.mantis_band_matrix) -> bands [N,H,W,R]; with want_cube also
        the 31-band cube, otherwise it is never written.  BASELINE config 4: the mantis-shrimp multi-receptor projection."""
        t = self.eng.torch
        w = np.ascontiguousarray(weights, np.float32)
        assert w.ndim == 2 and w.shape[1] == N_FEAT
        if not (frames.is_cuda and frames.dim() == 4 and frames.shape[3] == 3 and frames.dtype in (t.float32, t.uint8)):
            raise AvbError("forward_bands: expected a CUDA float32 / uint8 tensor [N,H,W,3]")
        frames = frames.contiguous()
        n, h, wd_, _ = frames.shape
        with t.cuda.device(self.eng.device):
            need = int(self.eng.lib.avb_mstpp_workspace_bytes(n, h, wd_, pad_multiple, int(centred)))
            if need <= 0:
                raise AvbError("avb_mstpp_workspace_bytes: bad geometry")
            key = self.eng.stream_ptr()
            ws = self._ws.get(key)
            if ws is None or ws.numel() < need:
                self._ws[key] = None
                ws = self._ws[key] = t.empty(need, dtype=t.uint8, device=self.eng.device)
            import hashlib
            wdev = self.eng.cached(("band_w", hashlib.sha1(w.tobytes()).hexdigest()), lambda: self.eng._dev(w))
            bands = t.empty((n, h, wd_, w.shape[0]), dtype=t.float32, device=self.eng.device)
            cube = t.empty((n, h, wd_, N_FEAT), dtype=t.float32, device=self.eng.device) if want_cube else None
            rc = self.eng.lib.avb_mstpp_forward_bands(self._h, frames.data_ptr(), int(frames.dtype == t.uint8),
                                                      None if cube is None else cube.data_ptr(), bands.data_ptr(), wdev.data_ptr(),
                                                      int(w.shape[0]), n, h, wd_, pad_multiple, int(centred), ws.data_ptr(), self.eng.stream_ptr())
        check(rc, "avb_mstpp_forward_bands")
        return (bands, cube) if want_cube else bands

    def forward(self, x):
        """MST_Plus_Plus.forward: NCHW float32 CUDA tensor [b,3,h,w] -> [b,31,h,w]."""
        return self._run(x.permute(0, 2, 3, 1), 8, False).permute(0, 3, 1, 2)

    __call__ = forward

    def predict_rgb_to_hsi(self, image: np.ndarray) -> np.ndarray:
        """predict_rgb_to_hsi_torch: HWC uint8 / float frame -> HWC float32 cube."""
        t = self.eng.torch
        a = np.asarray(image)
        assert a.ndim == 3 and a.shape[2] == 3, "Input must be HxWx3"
        if a.dtype == np.uint8:                                  # the uint8 path divides by 255 on the device
            dev = t.from_numpy(np.ascontiguousarray(a)).to(self.eng.device)
        elif np.issubdtype(a.dtype, np.integer):                 # predict_torch.py:12-15: any integer dtype is /255, never wrapped
            dev = t.from_numpy(np.ascontiguousarray(a.astype(np.float32) / np.float32(255.0))).to(self.eng.device)
        else:                                                    # predict_torch.py:16-19
            a = a.astype(np.float32)
            if a.max() > 1.001:
                a = np.clip(a / 255.0, 0.0, 1.0)
            dev = t.from_numpy(np.ascontiguousarray(a)).to(self.eng.device)
        return self._run(dev[None], 16, True)[0].cpu().numpy()

    def project_bands(self, cube, weights: np.ndarray):
        """cube: CUDA float32 [..., B]; weights [R, B] (e.g. tables.mantis_band_matrix) -> [..., R]
        (np.tensordot over the band axis: uv_helpers.py:142-146)."""
        t = self.eng.torch
        w = np.ascontiguousarray(weights, np.float32)
        assert cube.is_cuda and cube.dtype == t.float32 and cube.shape[-1] == w.shape[1]
        cube = cube.contiguous()
        wd = t.from_numpy(w).to(self.eng.device)
        out = t.empty(tuple(cube.shape[:-1]) + (w.shape[0],), dtype=t.float32, device=self.eng.device)
        npx = cube.numel() // cube.shape[-1]
        with t.cuda.device(self.eng.device):
            rc = self.eng.lib.avb_band_project_f32(cube.data_ptr(), wd.data_ptr(), out.data_ptr(), npx, w.shape[1], w.shape[0],
                                                   self.eng.stream_ptr())
        check(rc, "avb_band_project_f32")
        return out


def safe_norm_maps(maps):
    """uv_helpers.py:47-53 safe_norm applied independently to every map of a CUDA float32 tensor
    [..., R] (R interleaved maps, e.g. the band projections)."""
    eng = get_engine(maps.device)
    t = eng.torch
    assert maps.is_cuda and maps.dtype == t.float32
    maps = maps.contiguous()
    r = maps.shape[-1]
    out = t.empty_like(maps)
    scratch = t.empty(2 * r, dtype=t.int32, device=maps.device)
    with t.cuda.device(eng.device):
        rc = eng.lib.avb_safe_norm_f32(maps.data_ptr(), out.data_ptr(), maps.numel() // r, r, r, scratch.data_ptr(), eng.stream_ptr())
    check(rc, "avb_safe_norm_f32")
    return out


def mantis_bands(cube, model: MSTPlusPlus, wavelengths=None):
    """The ten mantis-shrimp bands of animals/mantis_shrimp.py:49-60 integrated over an MST++ cube."""
    lam = np.linspace(400.0, 700.0, N_FEAT, dtype=np.float32) if wavelengths is None else np.asarray(wavelengths, np.float32)
    return model.project_bands(cube, tables.mantis_band_matrix(lam))
