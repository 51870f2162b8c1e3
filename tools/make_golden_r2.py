#!/usr/bin/env python
"""Generate tests/golden/boundary_r2.npz by running the UNMODIFIED reference in place (container only):
the input-contract cases closed in round 2 --
  * HoneyBee on float32 / float64 / uint16 frames (honeybee.py:106, :166-173; uv_helpers.py:15-23),
  * HoneyBee(hsi_downsample=True) (uv_helpers.py:155-183), HoneyBee(blur_sigma_px > 2/3) (uv_helpers.py:67-73),
  * Cat on float32 / float64 / uint16 frames (cat.py:24, :80-112).

    python tools/make_golden_r2.py
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

HW = (54, 76)          # neither side divisible by 4 or 10: general INTER_AREA tables
BEE_VARIANTS = {
    "default": {},
    "down10": dict(hsi_downsample=True, hsi_scale=0.1),
    "down25_falsecolor": dict(hsi_downsample=True, hsi_scale=0.25, mapping_mode="falsecolor"),
    "sigma1p5": dict(blur_sigma_px=1.5),
    "sigma2p2_gray_mixed": dict(blur_sigma_px=2.2, adaptation="gray_world", mapping_mode="falsecolor_uv_mixed"),
}


def main():
    h, w = HW
    store = {}
    Bee = R.species("honeybee", "HoneyBee")
    Cat = R.species("cat", "Cat")
    u8 = [("natural", frames.natural(h, w)), ("bars", frames.bars(h, w)), ("noise", frames.noise(h, w, 2))]
    for vname, kw in BEE_VARIANTS.items():
        bee = Bee(**kw)
        cases = list(u8) if vname != "default" else []
        cases += frames.float_set(h, w)
        for name, f in cases:
            base, out = bee.visualize(f.copy())
            assert base.dtype == f.dtype and out.dtype == f.dtype
            store[f"bee/{vname}/{name}"] = out
    for name, f in frames.float_set(h, w):
        try:
            human, cat = Cat().visualize(f.copy())
        except Exception as e:      # noqa: BLE001  (cv2.resize rejects some integer depths: recorded, not hidden)
            print(f"cat/{name}: reference raises {type(e).__name__}: {str(e)[:100]}")
            continue
        assert human.dtype == f.dtype and cat.dtype == f.dtype
        store[f"cat/{name}/human"] = human
        store[f"cat/{name}/cat"] = cat
    path = os.path.join(ROOT, "tests", "golden", "boundary_r2.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes,", len(store), "arrays")
    for k in sorted(store):
        print(" ", k, store[k].dtype, store[k].shape)


if __name__ == "__main__":
    main()
