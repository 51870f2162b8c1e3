#!/usr/bin/env python
"""Generate tests/golden/mammals_float.npz by running the UNMODIFIED reference in place (container
only): the mammals' `visualize()` on float32 / float64 / uint16 frames (tests/frames.py float_set).

    python tools/make_golden_float.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import frames  # noqa: E402
import ref_loader as R  # noqa: E402

FLOAT_HW = (48, 64)
SPECIES = ["dog", "squirrel", "rat", "cow", "panda", "pig"]


def main():
    h, w = FLOAT_HW
    store = {}
    for sp in SPECIES:
        cls = R.species(sp, sp.capitalize())
        for name, f in frames.float_set(h, w):
            base, out = cls().visualize(f.copy())
            assert base.dtype == f.dtype and out.dtype == f.dtype and np.array_equal(base, f)
            store[f"{sp}/{name}"] = out
    path = os.path.join(ROOT, "tests", "golden", "mammals_float.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()
