#!/usr/bin/env python
"""Generate tests/golden/uv_species.npz by running the UNMODIFIED reference UV species in place (container only).

For every species: (baseline, view) of the reference's `visualize()` on small uint8 frames (structured + noise) and
one float32 frame, with the analytic branch of classic_rgb_to_hsi on CPU tensors (tools/ref_loader.py shim 3).  Also
checks that the default constructor arguments restated in oracle/uv_species.py equal the reference's signature.

    python tools/make_golden_uv.py [species ...]
"""
from __future__ import annotations

import importlib
import inspect
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import frames  # noqa: E402
import ref_loader as R  # noqa: E402

HW = (72, 104)
# module, class, oracle function / defaults name
SPECIES = [("reindeer", "Reindeer"), ("goldfish", "Goldfish"), ("damselfish", "Damselfish"), ("rat_uv", "RatUV"),
           ("anableps", "Anableps"), ("anchovy", "Anchovy"), ("guppy", "Guppy"), ("morpho", "Morpho"),
           ("heliconius", "Heliconius"), ("pieris", "Pieris"), ("kestrel", "Kestrel"), ("jumping_spider", "JumpingSpider"),
           ("dragonfly", "Dragonfly"), ("hummingbird", "Hummingbird"), ("mantis_shrimp", "MantisShrimp")]


def ref_class(module, cls):
    R._install_package_stub()
    R._patch_classic_hsi()
    mod = importlib.import_module(f"animals.{module}")
    chsi = importlib.import_module("ml.classic_rgb_to_hsi.classic_rgb_to_hsi")
    if hasattr(mod, "classic_rgb_to_hsi"):
        mod.classic_rgb_to_hsi = chsi.classic_rgb_to_hsi          # the module bound the unpatched function at import
    return getattr(mod, cls)


def cases(h, w):
    out = [("natural", frames.natural(h, w)), ("bars", frames.bars(h, w)), ("noise", frames.noise(h, w, 2)),
           ("dark", (frames.natural(h, w, 9) // 6).astype(np.uint8))]
    out.append(("f32_unit", frames.natural(h, w, 7).astype(np.float32) / np.float32(255.0)))
    return out


def check_defaults(module, cls, Ref):
    from oracle import uv_species as O
    name = module.upper()
    if not hasattr(O, name):
        return
    mine = getattr(O, name)
    sig = inspect.signature(Ref.__init__)
    ref = {k: v.default for k, v in sig.parameters.items() if k != "self" and v.default is not inspect.Parameter.empty}
    assert set(ref) == set(mine), (module, set(ref) ^ set(mine))
    for k, v in ref.items():
        a, b = mine[k], v
        if isinstance(b, tuple):
            assert np.allclose(np.array(a, float), np.array(b, float)), (module, k, a, b)
        else:
            assert a == b or (a is None and b is None), (module, k, a, b)


def main():
    want = sys.argv[1:]
    path = os.path.join(ROOT, "tests", "golden", "uv_species.npz")
    store = dict(np.load(path)) if (want and os.path.exists(path)) else {}
    store = {k: v for k, v in store.items() if not k.endswith("/base") or k.split("/")[1] == "natural"}
    h, w = HW
    for module, cls in SPECIES:
        if want and module not in want:
            continue
        Ref = ref_class(module, cls)
        check_defaults(module, cls, Ref)
        sp = Ref()
        for name, f in cases(h, w):
            base, out = sp.visualize(f.copy())
            assert base.dtype == f.dtype and out.dtype == f.dtype and base.shape == f.shape
            if name == "natural":                  # the baseline path (panorama warp + encode) is shared: one case per species
                store[f"{module}/{name}/base"] = base
            store[f"{module}/{name}/out"] = out
        if module == "rat_uv":
            for name, f in cases(h, w)[:2]:
                base, out = sp.visualize(f.copy(), mode="night")
                store[f"{module}/{name}_night/out"] = out
        print(module, "ok")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()


def mapper_golden():
    """tests/golden/uv_mapper_py.npz: reference uv_mappers.map_uv_purple_yellow on a seeded plane."""
    um = R.module("uv_mappers")
    U = (frames.natural(60, 84, 4)[..., 0].astype(np.float32) / 255.0) ** 2
    path = os.path.join(ROOT, "tests", "golden", "uv_mapper_py.npz")
    np.savez_compressed(path, U=U, out=um.map_uv_purple_yellow(U))
    print("wrote", path)
