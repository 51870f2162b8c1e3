"""Server-side batching beside the reference's `processimage` (utils.py:133-199) and its socket loop
(server/server.py:40-68), SURVEY.md 8f-4.

The reference handles one socket's frame at a time: JPEG bytes -> temp.jpg -> cv2.imread -> `filter.visualize(img)[1]` ->
tempexport.jpg -> base64 data URI, and the asyncio loop blocks on it.  Here every frame that is waiting -- from any
number of sockets -- is grouped by (species key, frame shape) and each group runs as ONE device batch through
`visualize_batch`; JPEG decode / encode stay on the host and in memory (no temp files).  `FrameBatcher.submit` is thread
safe and returns a `concurrent.futures.Future`; `process_image` is the drop-in for `processimage` (same arguments, same
data-URI result, same string keys: "human", "cat", "cow", ... as utils.py:145-191, plus keys for the UV species)."""
from __future__ import annotations

import base64
import threading
import time
from collections import OrderedDict, deque
from concurrent.futures import Future
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from . import animals as A

# the reference's lower-case server keys (utils.py:145-191) and, beyond them, one key per remaining registry entry
SERVER_KEYS: Dict[str, str] = {
    "cat": "Cat", "cow": "Cow", "goat": "Goat", "pig": "Pig", "sheep": "Sheep", "dog": "Dog", "rat": "Rat", "horse": "Horse",
    "rabbit": "Rabbit", "panda": "Panda", "squirrel": "Squirrel", "elephant": "Elephant", "lion": "Lion", "wolf": "Wolf", "fox": "Fox",
    "bear": "Bear", "raccoon": "Raccoon", "deer": "Deer", "kangaroo": "Kangaroo", "tiger": "Tiger", "honeybee": "HoneyBee",
    "reindeer": "Reindeer", "ratuv": "RatUV", "goldfish": "Goldfish", "damselfish": "Damselfish", "anableps": "Anableps",
    "anchovy": "Anchovy", "guppy": "Guppy", "morpho": "Morpho", "heliconius": "Heliconius", "pieris": "Pieris",
    "mantisshrimp": "MantisShrimp", "kestrel": "Kestrel", "jumpingspider": "JumpingSpider", "dragonfly": "Dragonfly",
    "hummingbird": "Hummingbird",
}


def _device_runner(device=None) -> Callable[[str, np.ndarray], np.ndarray]:
    """run(key, frames[k,H,W,3] uint8) -> the species view of every frame, [k,H,W,3] uint8 (index [1] of `visualize`)."""
    from .engine import get_engine
    species: Dict[str, object] = {}

    def run(key: str, frames: np.ndarray) -> np.ndarray:
        eng = get_engine(device)
        t = eng.torch
        sp = species.get(key)
        if sp is None:
            sp = species[key] = getattr(A, SERVER_KEYS[key])()
        with t.cuda.device(eng.device):
            pin = t.from_numpy(frames).pin_memory()
            dev = pin.to(eng.device, non_blocking=True)
            res = sp.visualize_batch(dev)
            out = t.empty(tuple(frames.shape), dtype=t.uint8).pin_memory()
            out.copy_(res[1], non_blocking=True)
            t.cuda.current_stream(eng.device).synchronize()
            return out.numpy()
    return run


class FrameBatcher:
    """Collects (frame, species key) requests from any number of producers and runs them as device batches.

    A worker thread takes everything that is pending (waiting up to `max_delay_ms` for stragglers once a request is in,
    never beyond `max_batch` frames per launch), groups it by (key, shape) in arrival order and resolves every request's
    Future with its own output frame.  `run_batch` is injectable (tests run the host logic without a GPU)."""

    def __init__(self, run_batch: Optional[Callable[[str, np.ndarray], np.ndarray]] = None, *, max_batch: int = 16,
                 max_delay_ms: float = 2.0, device=None):
        self._run = run_batch if run_batch is not None else _device_runner(device)
        self.max_batch, self.max_delay = int(max_batch), float(max_delay_ms) * 1e-3
        self._q: deque = deque()
        self._cv = threading.Condition()
        self._stop = False
        self.batches: List[Tuple[str, int]] = []            # (key, frames) of every launch, for inspection
        self._worker = threading.Thread(target=self._loop, name="avb-batcher", daemon=True)
        self._worker.start()

    def submit(self, frame: np.ndarray, animal: str) -> Future:
        fut: Future = Future()
        if animal != "human" and animal not in SERVER_KEYS:
            fut.set_exception(KeyError(f"no case implemented for {animal!r}"))             # utils.py:192-193 prints and fails later
            return fut
        if not (isinstance(frame, np.ndarray) and frame.ndim == 3 and frame.shape[2] == 3 and frame.dtype == np.uint8):
            fut.set_exception(AssertionError("frame must be a HxWx3 uint8 array"))
            return fut
        if animal == "human":                                                               # utils.py:146-147
            fut.set_result(frame)
            return fut
        with self._cv:
            if self._stop:
                raise RuntimeError("FrameBatcher is closed")
            self._q.append((animal, frame, fut))
            self._cv.notify()
        return fut

    def _loop(self):
        while True:
            with self._cv:
                while not self._q and not self._stop:
                    self._cv.wait()
                if self._stop and not self._q:
                    return
                deadline = time.monotonic() + self.max_delay
                while len(self._q) < self.max_batch and not self._stop:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        break
                    self._cv.wait(left)
                pending = list(self._q)
                self._q.clear()
            groups: "OrderedDict[Tuple[str, Tuple[int, ...]], list]" = OrderedDict()
            for animal, frame, fut in pending:
                groups.setdefault((animal, frame.shape), []).append((frame, fut))
            for (animal, _shape), items in groups.items():
                for a in range(0, len(items), self.max_batch):
                    part = items[a:a + self.max_batch]
                    try:
                        out = self._run(animal, np.stack([f for f, _ in part]))
                        self.batches.append((animal, len(part)))
                        for i, (_, fut) in enumerate(part):
                            fut.set_result(np.array(out[i], copy=True))
                    except Exception as e:          # noqa: BLE001  (a failed batch fails its own requests, the loop lives on)
                        for _, fut in part:
                            if not fut.done():
                                fut.set_exception(e)

    def close(self):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        self._worker.join(timeout=10)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


_default: Optional[FrameBatcher] = None
_default_lock = threading.Lock()


def default_batcher() -> FrameBatcher:
    global _default
    with _default_lock:
        if _default is None:
            _default = FrameBatcher()
        return _default


def process_image(imagedata: bytes, animal: str, batcher: Optional[FrameBatcher] = None) -> str:
    """Drop-in for utils.py:133-199 `processimage`: JPEG bytes in, JPEG data URI out; decoded BGR as cv2.imread gives it
    (the species index channels 0/1/2, never "R/G/B" -- the reference feeds BGR here too, utils.py:141-142)."""
    import cv2
    img = cv2.imdecode(np.frombuffer(imagedata, np.uint8), cv2.IMREAD_COLOR)
    if img is None:
        raise ValueError("could not decode the image")
    out = (batcher or default_batcher()).submit(img, animal).result()
    ok, enc = cv2.imencode(".jpg", out)
    if not ok:
        raise ValueError("could not encode the result")
    return "data:image/jpeg;base64," + base64.b64encode(enc.tobytes()).decode("utf-8")
