"""The UV-species oracle (oracle/uv_species.py) against golden vectors produced by the unmodified reference
(tools/make_golden_uv.py): bit-exact for uint8 frames, exact for float32 frames (same NumPy / OpenCV / torch build)."""
import numpy as np
import pytest

import frames
from oracle import uv_species as O

HW = (72, 104)


def _inputs(h, w):
    return {"natural": frames.natural(h, w), "bars": frames.bars(h, w), "noise": frames.noise(h, w, 2),
            "dark": (frames.natural(h, w, 9) // 6).astype(np.uint8),
            "f32_unit": frames.natural(h, w, 7).astype(np.float32) / np.float32(255.0)}


def _names(golden):
    return sorted({k.split("/")[0] for k in golden("uv_species")})


def test_every_golden_species_has_an_oracle(golden):
    for name in _names(golden):
        assert hasattr(O, name), name


@pytest.mark.parametrize("case", ["natural", "bars", "noise", "dark", "f32_unit"])
def test_oracle_equals_reference(case, golden):
    g = golden("uv_species")
    f = _inputs(*HW)[case]
    for name in _names(golden):
        base, out = getattr(O, name)(f.copy())
        if f"{name}/{case}/base" in g:
            assert np.array_equal(base, g[f"{name}/{case}/base"]), f"{name}/{case}/base"
        assert np.array_equal(out, g[f"{name}/{case}/out"]), f"{name}/{case}/out"
    if case in ("natural", "bars"):
        _, out = O.rat_uv(f.copy(), mode="night")
        assert np.array_equal(out, g[f"rat_uv/{case}_night/out"])
