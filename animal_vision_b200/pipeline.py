"""Batched, overlapped host feed for video: the B200-side counterpart of the reference's
one-frame-at-a-time loop (main.py:60-71: get_image -> visualize -> render).

Frames are independent, so a batch is cut into chunks that flow through three CUDA streams:
  H2D copy of chunk i+1  ||  kernels of chunk i  ||  D2H copy of chunk i-1
from / to pinned host memory.  No collective, no host synchronisation inside the loop.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .engine import get_engine


class HostBatchPipeline:
    def __init__(self, device=None, chunk_frames: int = 4, depth: int = 3):
        self.eng = get_engine(device)
        t = self.eng.torch
        self.chunk = int(chunk_frames)
        self.depth = int(depth)
        with t.cuda.device(self.eng.device):
            self.s_in, self.s_run, self.s_out = (t.cuda.Stream() for _ in range(3))
        self._slots = {}

    def _slot_bufs(self, shape, n_out):
        """`depth` rotating device buffers: one input + n_out outputs of `shape` ([chunk,H,W,3])."""
        key = (tuple(shape), n_out)
        if key not in self._slots:
            t = self.eng.torch
            self._slots[key] = [
                (t.empty(shape, dtype=t.uint8, device=self.eng.device),
                 [t.empty(shape, dtype=t.uint8, device=self.eng.device) for _ in range(n_out)],
                 [None, None])       # [event: outputs copied out (slot free), unused]
                for _ in range(self.depth)]
        return self._slots[key]

    @staticmethod
    def n_outputs(species) -> int:
        return int(getattr(species, "N_OUTPUTS", 1))         # class attribute: Cat (and its subclasses) produce two frames

    def pinned_like(self, frames, n_out: int):
        t = self.eng.torch
        return [t.empty(tuple(frames.shape), dtype=t.uint8).pin_memory() for _ in range(n_out)]

    def run(self, jobs: Sequence[Tuple[object, object, List[object]]]):
        """jobs: (species, frames_host, outs_host) with frames_host / outs_host pinned uint8
        tensors [N,H,W,3] (outs_host: one tensor per species output, see n_outputs).
        Enqueues everything, then waits once.  Returns (h2d_bytes, d2h_bytes)."""
        t = self.eng.torch
        h2d = d2h = 0
        slot_i = 0
        with t.cuda.device(self.eng.device):
            for species, src, outs in jobs:
                n = src.shape[0]
                n_out = len(outs)
                for a in range(0, n, self.chunk):
                    b = min(n, a + self.chunk)
                    shape = (self.chunk,) + tuple(src.shape[1:])
                    slots = self._slot_bufs(shape, n_out)
                    d_in, d_outs, ev = slots[slot_i % self.depth]
                    slot_i += 1
                    if ev[0] is not None:                    # slot reused: its last D2H must be done
                        self.s_in.wait_event(ev[0])
                        self.s_run.wait_event(ev[0])
                    with t.cuda.stream(self.s_in):
                        d_in[: b - a].copy_(src[a:b], non_blocking=True)
                        e_in = t.cuda.Event()
                        e_in.record(self.s_in)
                    with t.cuda.stream(self.s_run):
                        self.s_run.wait_event(e_in)
                        if n_out == 2:
                            species.visualize_batch(d_in[: b - a], out=(d_outs[0][: b - a], d_outs[1][: b - a]))
                        else:
                            species.visualize_batch(d_in[: b - a], out=d_outs[0][: b - a])
                        e_run = t.cuda.Event()
                        e_run.record(self.s_run)
                    with t.cuda.stream(self.s_out):
                        self.s_out.wait_event(e_run)
                        for d, o in zip(d_outs, outs):
                            o[a:b].copy_(d[: b - a], non_blocking=True)
                        e_out = t.cuda.Event()
                        e_out.record(self.s_out)
                    ev[0] = e_out
                    h2d += (b - a) * int(np.prod(src.shape[1:]))
                    d2h += n_out * (b - a) * int(np.prod(src.shape[1:]))
            self.s_out.synchronize()
            self.s_run.synchronize()
        return h2d, d2h
