"""GPU parity of the UV species against golden vectors generated from the unmodified reference
(tools/make_golden_uv.py -> tests/golden/uv_species.npz) and against the oracle at other shapes.
Tolerances (BASELINE.json north_star): <= 1 LSB on uint8 outputs; <= 1e-5 relative on float32 outputs, measured against
the output range [0, 1] (the species compute in float32 throughout; the reference mixes in float64 steps)."""
import os

import numpy as np
import pytest

import frames
from oracle import uv_species as O

pytestmark = pytest.mark.gpu

SPECIES = {   # golden / oracle name -> (module, class)
    "reindeer": ("reindeer", "Reindeer"), "goldfish": ("goldfish", "Goldfish"), "damselfish": ("damselfish", "Damselfish"),
    "rat_uv": ("rat_uv", "RatUV"), "anableps": ("anableps", "Anableps"), "anchovy": ("anchovy", "Anchovy"),
    "guppy": ("guppy", "Guppy"), "morpho": ("morpho", "Morpho"), "heliconius": ("heliconius", "Heliconius"),
    "pieris": ("pieris", "Pieris"), "kestrel": ("kestrel", "Kestrel"), "jumping_spider": ("jumping_spider", "JumpingSpider"),
    "dragonfly": ("dragonfly", "Dragonfly"), "hummingbird": ("hummingbird", "Hummingbird"),
    "mantis_shrimp": ("mantis_shrimp", "MantisShrimp"),
}
HW = (72, 104)          # tools/make_golden_uv.py


def _cls(name):
    import importlib
    mod, cls = SPECIES[name]
    return getattr(importlib.import_module(f"animal_vision_b200.animals.{mod}"), cls)


def _inputs(h, w):
    return {"natural": frames.natural(h, w), "bars": frames.bars(h, w), "noise": frames.noise(h, w, 2),
            "dark": (frames.natural(h, w, 9) // 6).astype(np.uint8),
            "f32_unit": frames.natural(h, w, 7).astype(np.float32) / np.float32(255.0)}


def _cmp(got, ref, what, max_frac=0.02, ftol=2e-5):
    assert got.shape == ref.shape and got.dtype == ref.dtype, what
    if ref.dtype == np.uint8:
        d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
        assert d.max() <= 1, f"{what}: max diff {d.max()} LSB"
        assert (d > 0).mean() <= max_frac, f"{what}: {(d > 0).mean():.4f} of bytes differ by 1 LSB"
    else:
        assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= ftol, f"{what}: {np.abs(got - ref).max():.3e}"


def _available():
    return [n for n in SPECIES if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "animal_vision_b200", "animals", SPECIES[n][0] + ".py"))]


@pytest.mark.parametrize("name", _available())
def test_against_reference_golden(name, golden):
    g = golden("uv_species")
    sp = _cls(name)()
    ins = _inputs(*HW)
    seen = 0
    for key, ref in g.items():
        s, case, which = key.split("/")
        if s != name or case.endswith("_night"):
            continue
        base, out = sp.visualize(ins[case])
        _cmp(base if which == "base" else out, ref, key)
        seen += 1
    assert seen >= 8, f"no golden vectors for {name}"


def test_rat_uv_night_mode(golden):
    g = golden("uv_species")
    sp = _cls("rat_uv")()
    ins = _inputs(*HW)
    for case in ("natural", "bars"):
        _, out = sp.visualize(ins[case], mode="night")
        _cmp(out, g[f"rat_uv/{case}_night/out"], f"rat_uv night {case}")


@pytest.mark.parametrize("name", _available())
def test_against_oracle_other_shape_and_batch(name):
    import torch
    h, w = 135, 241
    f0, f1 = frames.natural(h, w, 3), frames.checker(h, w, 11)
    sp = _cls(name)()
    fn = getattr(O, name)
    refs = [fn(f) for f in (f0, f1)]
    batch = torch.from_numpy(np.stack([f0, f1])).cuda()
    base, out = sp.visualize_batch(batch)
    for i in range(2):
        _cmp(base[i].cpu().numpy(), refs[i][0], f"{name} batch base {i}")
        _cmp(out[i].cpu().numpy(), refs[i][1], f"{name} batch out {i}", max_frac=0.03)
