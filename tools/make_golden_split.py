#!/usr/bin/env python
"""tests/golden/split_compare.npz: the reference's VideoRenderer.make_split_frame (renderers/video.py:198-245) on two
small frames, run in place from /root/reference (container only).   python tools/make_golden_split.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import frames  # noqa: E402
import ref_loader as R  # noqa: E402

import importlib, types  # noqa: E402
pkg = types.ModuleType("renderers")             # an unrelated installed package is called `renderers` too: point the name at the reference
pkg.__path__ = [os.path.join(R.REF, "renderers")]
sys.modules["renderers"] = pkg
R._install_package_stub()
mod = importlib.import_module("renderers.video")
cls = mod.VideoRenderer
r = cls.__new__(cls)                      # make_split_frame / _draw_label use no instance state
store = {}
for name, (h, w) in (("small", (120, 200)), ("tall", (300, 161))):
    a, b = frames.natural(h, w, 1), frames.noise(h, w, 2)
    store[f"{name}/out"] = r.make_split_frame(a, b)
    store[f"{name}/out_noseam_labels"] = r.make_split_frame(a, b, left_label="A", right_label="Dog view", draw_seam=False)
path = os.path.join(ROOT, "tests", "golden", "split_compare.npz")
np.savez_compressed(path, **store)
print("wrote", path, os.path.getsize(path))
