"""Error behaviour of the C ABI: bad arguments come back as negative AVB_E_* codes with a message,
never as a crash; the Python layer turns them into AvbError."""
import ctypes as C

import numpy as np
import pytest


def _lib():
    from animal_vision_b200 import _abi
    return _abi.load()


def test_encode_table_rejects_bad_thresholds():
    lib = _lib()
    out = np.zeros(2048, np.uint32)
    bad = np.linspace(0.1, 0.9, 255).astype(np.float32)
    bad[100] = bad[99]                                      # not strictly increasing
    rc = lib.avb_build_encode_table(bad.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), out.size)
    assert rc < 0 and b"strictly increasing" in lib.avb_last_error()
    rc = lib.avb_build_encode_table(None, out.ctypes.data_as(C.c_void_p), out.size)
    assert rc == -1


def test_encode_table_reproduces_the_reference_quantiser_on_the_host():
    """Decode the bucketed table in NumPy and compare with the reference's step function."""
    from animal_vision_b200 import tables
    from oracle import colorimetry as Cm
    lib = _lib()
    thr = np.ascontiguousarray(tables.encode_thresholds(False))
    tab = np.zeros(2048, np.uint32)
    n = lib.avb_build_encode_table(thr.ctypes.data_as(C.c_void_p), tab.ctypes.data_as(C.c_void_p), tab.size)
    assert 4 < n <= 2048
    key_min, shift, nb = int(tab[0]), int(tab[1]), int(tab[2])
    ent = tab[4:4 + nb]
    x = np.concatenate([np.random.default_rng(0).random(200000).astype(np.float32) ** 3,
                        thr, np.nextafter(thr, np.float32(0)), np.float32([0, 1, 1e-9, -0.5, 7.0])]).astype(np.float32)
    b = np.clip(x, 0, 1).astype(np.float32).view(np.uint32)
    b = np.maximum(b, np.uint32(key_min << shift))
    e = ent[(b - np.uint32(key_min << shift)) >> np.uint32(shift)]
    got = (e & 0xFF) + ((b & np.uint32((1 << shift) - 1)) >= (e >> 8))
    ref = Cm.encode_tail(x.reshape(-1, 1, 1).repeat(3, 2), np.uint8)[:, 0, 0]
    assert np.array_equal(got.astype(np.uint8), ref)


def test_workspace_queries_reject_bad_geometry():
    lib = _lib()
    assert lib.avb_uv_workspace_bytes(0, 10, 10, 0) == 0
    assert lib.avb_uv_workspace_bytes(1, 10, 10, 9) == 0
    assert lib.avb_uv_workspace_bytes(2, 100, 200, 0) > 2 * 100 * 200 * 4
    assert lib.avb_mstpp_workspace_bytes(1, 1, 10, 8, 0) == 0
    assert lib.avb_mstpp_workspace_bytes(1, 64, 64, 12, 0) == 0          # pad multiple must be a multiple of 8
    assert lib.avb_mstpp_workspace_bytes(1, 64, 64, 8, 0) > 0


def test_mstpp_create_rejects_wrong_parameter_count():
    lib = _lib()
    h = C.c_void_p()
    blob = np.zeros(100, np.float32)
    rc = lib.avb_mstpp_create(blob.ctypes.data_as(C.c_void_p), blob.size, C.byref(h))
    assert rc == -1 and b"1619625" in lib.avb_last_error()


@pytest.mark.gpu
def test_compute_entry_points_validate_arguments():
    import torch
    from animal_vision_b200 import tables
    from animal_vision_b200._abi import AvbError
    from animal_vision_b200.engine import get_engine
    eng = get_engine()
    f = torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(f)
    T = tables.dichromat_matrix(0.58, 0.65)
    with pytest.raises(AvbError, match="ksize"):
        eng.dichromat_blur(f, out, T, np.ones(4, np.float32) / 4)          # even tap count
    with pytest.raises(AvbError, match="ksize"):
        eng.dichromat_blur(f, out, T, np.ones(35, np.float32) / 35)        # radius beyond the instantiated range
    with pytest.raises(AvbError, match="expected a CUDA uint8 tensor"):
        eng.dichromat_blur(f.float(), out, T, np.ones(3, np.float32) / 3)
    with pytest.raises(AvbError, match="expected a CUDA uint8 tensor"):
        eng.dichromat_blur(f[..., :2], out, T, np.ones(3, np.float32) / 3)
    lib = eng.lib
    rc = lib.avb_colorimetric_u8(None, out.data_ptr(), 1, 8, 8, 192, 24, 192, 24, eng.dec.data_ptr(), eng.dec_raw.data_ptr(),
                                 eng.enc.data_ptr(), T.ctypes.data_as(C.c_void_p), None, 0, None, None)
    assert rc == -1 and b"null frame pointer" in lib.avb_last_error()
    rc = lib.avb_colorimetric_u8(f.data_ptr(), out.data_ptr(), 1, 8, 8, 192, 8, 192, 24, eng.dec.data_ptr(), eng.dec_raw.data_ptr(),
                                 eng.enc.data_ptr(), T.ctypes.data_as(C.c_void_p), None, 0, None, None)
    assert rc == -1 and b"row stride" in lib.avb_last_error()
    # a failed call leaves the library usable
    eng.dichromat_blur(f, out, T, np.ones(3, np.float32) / 3)
    torch.cuda.synchronize()
