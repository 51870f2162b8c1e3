"""GPU parity for Cat: centre zoom must be bit-exact (integer arithmetic), the wide-FOV view
<= 1 LSB (float32 tail instead of the reference's float64 tail)."""
import numpy as np
import pytest

import frames
from oracle import mammals as M

pytestmark = pytest.mark.gpu


def _check(human, cat, ref_h, ref_c, what):
    assert human.dtype == cat.dtype == np.uint8 and human.shape == ref_h.shape and cat.shape == ref_c.shape
    assert np.array_equal(human, ref_h), f"{what}: centre zoom differs ({(human != ref_h).mean():.5f} of bytes)"
    d = np.abs(cat.astype(np.int16) - ref_c.astype(np.int16))
    assert d.max() <= 1, f"{what}: cat view max diff {d.max()} LSB"
    assert (d > 0).mean() <= 0.02, f"{what}: {(d > 0).mean():.4f} of bytes differ"


def test_cat_against_golden(golden, golden_meta):
    from animal_vision_b200.animals import Cat
    h, w = golden_meta["small_hw"]
    g = golden("cat")
    for name, f in frames.parity_set(h, w):
        human, cat = Cat().visualize(f)
        assert human is not f and cat is not f
        _check(human, cat, g[f"human/{name}"], g[f"cat/{name}"], name)


@pytest.mark.parametrize("hw", [(61, 67), (270, 480), (33, 300), (1080, 1920)])
def test_cat_against_oracle(hw):
    from animal_vision_b200.animals import Cat
    h, w = hw
    cases = list(frames.parity_set(h, w)) if h < 1000 else [("noise0", frames.noise(h, w, 0)), ("natural", frames.natural(h, w))]
    for name, f in cases:
        ref_h, ref_c = M.cat_visualize(f)
        human, cat = Cat().visualize(f)
        _check(human, cat, ref_h, ref_c, f"{name}/{h}x{w}")


def test_cat_batch():
    import torch
    from animal_vision_b200.animals import Cat
    fs = [frames.noise(120, 200, s) for s in range(3)] + [frames.le1(120, 200)]
    batch = torch.from_numpy(np.stack(fs)).cuda()
    human, cat = Cat().visualize_batch(batch)
    for i, f in enumerate(fs):
        ref_h, ref_c = M.cat_visualize(f)
        _check(human[i].cpu().numpy(), cat[i].cpu().numpy(), ref_h, ref_c, f"batch[{i}]")


@pytest.mark.parametrize("where", [(0, 0), (5, 17), (63, 100), (64, 3), (199, 299), (130, 0)])
def test_cat_single_witness_byte_decides_the_normalisation(where):
    """get_normalized_image divides by 255 iff the frame maximum exceeds 1 (animal_utils.py:41-50).  The flag pass looks
    for a witness in every 64th row first and scans the whole frame only when the sample holds none: a frame of zeros and
    ones with ONE byte of 2 -- in a sampled row, in an unsampled one, in the last row, in the blind strip's columns -- must
    take the /255 branch, and the same frame without it the other branch."""
    import torch
    from animal_vision_b200.animals import Cat
    f = frames.le1(200, 300).copy()
    assert f.max() <= 1
    g = f.copy()
    g[where[0], where[1], 1] = 2
    batch = torch.from_numpy(np.stack([f, g, f])).cuda()
    human, cat = Cat().visualize_batch(batch)
    for i, fr in enumerate((f, g, f)):
        ref_h, ref_c = M.cat_visualize(fr)
        _check(human[i].cpu().numpy(), cat[i].cpu().numpy(), ref_h, ref_c, f"witness {where} frame {i}")
    assert not np.array_equal(cat[0].cpu().numpy(), cat[1].cpu().numpy()), "the two branches must differ on this frame"


def test_cat_without_fov_warp(golden, golden_meta):
    """Class switch ENABLE_FOV_WARP = False (cat.py:21): centre zoom + L/M merge + blur, no warp."""
    from animal_vision_b200.animals import Cat
    NoWarp = type("NoWarp", (Cat,), {"ENABLE_FOV_WARP": False})
    h, w = golden_meta["small_hw"]
    g = golden("cat")
    fr = dict(frames.parity_set(h, w))
    n = 0
    for key in g:
        if key.startswith("nowarp_cat/"):
            name = key.split("/")[1]
            human, cat = NoWarp().visualize(fr[name])
            _check(human, cat, g[f"nowarp_human/{name}"], g[key], key)
            n += 1
    assert n >= 3
    f = frames.natural(270, 480)
    ref_h, ref_c = M.cat_visualize(f, fov_warp=False)
    _check(*NoWarp().visualize(f), ref_h, ref_c, "nowarp 270x480")
