"""Float / non-uint8 frames: "float in => float [0,1] out, no quantisation" (reference
animals/dog.py:56-59) and the data-dependent normalisation (animals/animal_utils.py:41-50).
Golden vectors: the unmodified reference on tests/frames.py float_set (tools/make_golden_float.py).
CPU: the oracle reproduces them; GPU: the float32 device path (avb_dichromat_f32) stays within the
north_star's 1e-5 relative bar on fp32 data (and 1 unit on the uint16 case)."""
import numpy as np
import pytest

import frames
from oracle import mammals as M

FLOAT_HW = (48, 64)
SPECIES = ["dog", "squirrel", "rat", "cow", "panda", "pig"]
REL_TOL = 1e-5          # BASELINE.json north_star: <= 1e-5 relative error on fp32 data
ABS_FLOOR = 2e-6        # values near 0 (encoded black): a few float32 ulps of 1.0


def _cases():
    return frames.float_set(*FLOAT_HW)


def test_oracle_reproduces_reference_on_float_frames(golden):
    g = golden("mammals_float")
    assert len(g) == len(SPECIES) * 4
    for sp in SPECIES:
        for name, f in _cases():
            base, out = M.mammal_visualize(f.copy(), sp)
            ref = g[f"{sp}/{name}"]
            assert base.dtype == f.dtype and out.dtype == ref.dtype == f.dtype, (sp, name)
            if np.issubdtype(f.dtype, np.integer):
                assert np.array_equal(out, ref), (sp, name)
            else:
                assert out.min() >= 0.0 and out.max() <= 1.0
                np.testing.assert_allclose(out, ref, rtol=1e-6, atol=1e-7, err_msg=f"{sp}/{name}")


@pytest.mark.gpu
@pytest.mark.parametrize("sp", SPECIES)
def test_gpu_float_path_against_reference(sp, golden):
    import animal_vision_b200.animals as A
    g = golden("mammals_float")
    animal = A.MAMMALS[sp]()
    for name, f in _cases():
        src = f.copy()
        base, out = animal.visualize(src)
        ref = g[f"{sp}/{name}"]
        assert base is src and np.array_equal(src, f), "baseline is the caller's own, untouched, array (dog.py:61)"
        assert out.dtype == f.dtype and out.shape == f.shape, (sp, name)
        if np.issubdtype(f.dtype, np.integer):
            d = np.abs(out.astype(np.int64) - ref.astype(np.int64))
            assert d.max() <= 1 and (d > 0).mean() <= 0.02, f"{sp}/{name}: max {d.max()}, {(d > 0).mean():.4f} differ"
        else:
            err = np.abs(out.astype(np.float64) - ref.astype(np.float64))
            bound = REL_TOL * np.abs(ref.astype(np.float64)) + ABS_FLOOR
            assert (err <= bound).all(), f"{sp}/{name}: worst {np.max(err / np.maximum(np.abs(ref), 1e-3)):.3e} relative"


@pytest.mark.gpu
def test_gpu_float_batch_entry_point():
    """engine.dichromat_f32 on a batch whose frames take different normalisation branches."""
    import torch
    import animal_vision_b200.animals as A
    from animal_vision_b200 import tables
    from animal_vision_b200._abi import AVB_F32_GAUSS
    from animal_vision_b200.engine import get_engine
    h, w = FLOAT_HW
    a = frames.natural(h, w).astype(np.float32) / np.float32(255.0)      # max <= 1: used as is
    b = frames.noise(h, w, 7).astype(np.float32)                         # max > 1: divided by 255
    batch = torch.from_numpy(np.stack([a, b])).cuda()
    eng = get_engine(batch.device)
    out, tmp = torch.empty_like(batch), torch.empty_like(batch)
    dog = A.Dog()
    taps = tables.gaussian_taps(tables.gaussian_ksize(dog.SIGMA), dog.SIGMA)
    eng.dichromat_f32(batch, out, tmp, dog._matrix(), AVB_F32_GAUSS, taps=taps)
    got = out.cpu().numpy()
    for k, f in enumerate((a, b)):
        ref = M.mammal_visualize(f, "dog")[1]
        np.testing.assert_allclose(got[k], ref, rtol=REL_TOL, atol=ABS_FLOOR)
