"""Host-side tables (CPU): the pixel-independent tables the kernels consume, checked by running a
NumPy emulation of the DEVICE algorithm on them against the oracle / golden vectors."""
import numpy as np
import pytest

import frames
from animal_vision_b200 import tables
from oracle import colorimetry as C
from oracle import mammals as M


def _reflect(idx, n):
    if n == 1:
        return np.zeros_like(idx)
    p = 2 * n - 2
    m = np.mod(idx, p)
    return np.where(m >= n, p - m, m)


def _streak_device_emulation(img, name):
    """k2_streak.cu in NumPy: LUT decode -> per-row folded 3x3 -> one combined x correlation -> tail."""
    r = M.RECIPES[name]
    H, W = img.shape[:2]
    tab = tables.streak_row_table(H, tables.dichromat_matrix(r.alpha, r.s_scale), *r.streak)
    lin = tables.decode_lut(C.divides_by_255(img))[img]
    out = np.empty((H, W, 3), np.float32)
    for y in range(H):
        P = (lin[y] @ tab[y, 33:42].reshape(3, 3).T).astype(np.float32)
        rr = int(tab[y, 42])
        taps = tab[y, 16 - rr:16 + rr + 1]
        idx = _reflect(np.arange(-rr, W + rr), W)
        acc = np.zeros((W, 3), np.float32)
        for t, w in enumerate(taps):
            acc += w * P[idx[t:t + W]]
        out[y] = acc
    if r.chroma:
        out = M.chroma_compression(out, r.chroma)
    return C.encode_tail(out, np.uint8)


@pytest.mark.parametrize("name", ["cow", "panda", "pig"])
def test_streak_row_table_reproduces_reference(name, golden, golden_meta):
    h, w = golden_meta["small_hw"]
    fr = dict(frames.parity_set(h, w))
    g = golden("mammals")
    n = 0
    for key, ref in g.items():
        s, case = key.split("/")
        if s != name:
            continue
        d = np.abs(_streak_device_emulation(fr[case], name).astype(int) - ref.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3, key
        n += 1
    assert n >= 3


def test_streak_row_table_shape_and_radius():
    tab = tables.streak_row_table(2160, tables.dichromat_matrix(0.6, 0.95), 0.5, 0.8, 2.6, 8.0)
    assert tab.shape == (2160, tables.STREAK_TAB) and tab.dtype == np.float32
    assert tab[:, 42].max() <= 16
    np.testing.assert_allclose(tab[:, :33].sum(1), 1.0, atol=2e-6)      # composed taps stay normalised
    # rank-2 factors: (mix_y P) Q reproduces the row's 3x3 to float32 rounding for every dichromat matrix
    assert (tab[:, 55] == 1.0).all()
    for y in (0, 700, 1080, 2159):
        A = tab[y, 33:42].reshape(3, 3).astype(np.float64)
        PQ = tab[y, 43:49].reshape(3, 2).astype(np.float64) @ tab[y, 49:55].reshape(2, 3).astype(np.float64)
        assert np.abs(PQ - A).max() <= 4e-7 * np.abs(A).max()
    full = np.array([[0.9, 0.1, 0.0], [0.0, 0.3, 0.8], [0.5, 0.5, 0.1]], np.float32)       # rank 3: no factors
    assert tables.rank2_factor(full) is None
    assert (tables.streak_row_table(64, full, 0.5, 0.8, 2.6, 8.0)[:, 55] == 0.0).all()


def test_encode_thresholds_are_the_reference_step_function():
    thr = tables.encode_thresholds(False)
    assert thr.shape == (255,) and np.all(np.diff(thr) > 0)
    q = C.encode_tail(thr.reshape(-1, 1, 1).repeat(3, 2), np.uint8)[:, 0, 0]
    assert np.array_equal(q, np.arange(1, 256))
    below = np.nextafter(thr, np.float32(0)).astype(np.float32)
    q = C.encode_tail(below.reshape(-1, 1, 1).repeat(3, 2), np.uint8)[:, 0, 0]
    assert np.array_equal(q, np.arange(0, 255))


def test_decode_luts_match_reference_expressions():
    v = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    assert np.array_equal(tables.decode_lut(True)[v], C.decode_srgb(C.normalize_frame(v)))
    assert np.array_equal(tables.decode_lut(True), C.decode_lut_u8(True))
    assert np.array_equal(tables.decode_lut(False), C.decode_lut_u8(False))
