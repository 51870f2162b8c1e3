"""Synthetic frames shared by the golden generator, the oracle tests and the GPU parity tests.

Noise alone is a poor parity input: a 29-tap blur collapses it to flat grey and hides errors
(SURVEY.md 8d), so every parity case also runs the structured set below.
"""
from __future__ import annotations

import numpy as np


def noise(h: int, w: int, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def ramps(h: int, w: int) -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w]
    r = xx * 255 // max(1, w - 1)
    g = yy * 255 // max(1, h - 1)
    b = (xx + yy) * 255 // max(1, w + h - 2)
    return np.stack([r, g, b], 2).astype(np.uint8)


def bars(h: int, w: int) -> np.ndarray:
    """Eight vertical colour bars (white, yellow, cyan, green, magenta, red, blue, black)."""
    cols = np.array([[255, 255, 255], [255, 255, 0], [0, 255, 255], [0, 255, 0],
                     [255, 0, 255], [255, 0, 0], [0, 0, 255], [0, 0, 0]], np.uint8)
    idx = (np.arange(w) * 8 // w).clip(0, 7)
    return np.repeat(cols[idx][None], h, 0).copy()


def checker(h: int, w: int, cell: int = 7) -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w]
    m = ((yy // cell + xx // cell) % 2).astype(np.uint8)
    return np.stack([m * 255, (1 - m) * 200 + 20, m * 90 + 60], 2).astype(np.uint8)


def impulses(h: int, w: int) -> np.ndarray:
    f = np.zeros((h, w, 3), np.uint8)
    for (y, x, c) in [(0, 0, 0), (h - 1, w - 1, 1), (h // 2, w // 2, 2), (1, w - 2, 0), (h - 2, 1, 2),
                      (h // 3, 2 * w // 3, 1)]:
        f[min(max(y, 0), h - 1), min(max(x, 0), w - 1), c] = 255
    f[h // 4, w // 4] = 255
    return f


def constant(h: int, w: int, v: int) -> np.ndarray:
    return np.full((h, w, 3), v, np.uint8)


def le1(h: int, w: int, seed: int = 3) -> np.ndarray:
    """All bytes in {0,1}: the frame max is <= 1, so get_normalized_image does NOT divide by 255."""
    return np.random.default_rng(seed).integers(0, 2, (h, w, 3), dtype=np.uint8)


def natural(h: int, w: int, seed: int = 5) -> np.ndarray:
    """Smooth blobs + edges + mild noise: closer to a photograph than uniform noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for _ in range(12):
        cy, cx = rng.random() * h, rng.random() * w
        s = (0.05 + 0.25 * rng.random()) * max(h, w)
        col = rng.random(3).astype(np.float32)
        img += col * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))[..., None]
    img /= max(1e-6, img.max())
    img[:, w // 2:] = 1.0 - 0.8 * img[:, w // 2:]
    img[h // 3: h // 3 + max(2, h // 20)] *= 0.25
    img += 0.02 * rng.standard_normal((h, w, 3)).astype(np.float32)
    return (np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8)


STRUCTURED = {
    "ramps": ramps,
    "bars": bars,
    "checker": checker,
    "impulses": impulses,
    "zeros": lambda h, w: constant(h, w, 0),
    "full": lambda h, w: constant(h, w, 255),
    "le1": le1,
    "natural": natural,
}


def parity_set(h: int, w: int):
    """(name, frame) pairs: seeded noise + the structured set."""
    yield "noise0", noise(h, w, 0)
    for name, fn in STRUCTURED.items():
        yield name, fn(h, w)


def float_set(h: int, w: int):
    """Non-uint8 frames for the "float in => float out" contract (dog.py:56-59) and the data-dependent
    normalisation branch (animal_utils.py:41-50): unit-range float32 (not divided), 0..255 float32
    (divided), unit-range float64, and an integer dtype other than uint8."""
    return [
        ("f32_unit", (natural(h, w).astype(np.float32) / np.float32(255.0))),
        ("f32_255", noise(h, w, 7).astype(np.float32)),
        ("f64_unit", np.random.default_rng(11).random((h, w, 3))),
        ("u16_255", noise(h, w, 9).astype(np.uint16)),
    ]
