"""Regression tests for the code-level findings of the round-1 review: per-device shared-memory opt-in,
content-keyed band-table cache, output count of Cat subclasses, per-thread staging, the stale-library
guard and the guard bands around every kernel family's outputs (no sanitizer on the pool)."""
import threading

import numpy as np
import pytest

import frames as F


# ------------------------------------------------------------------ CPU
def test_band_table_cache_key_is_content_not_id():
    from animal_vision_b200.animals import HoneyBee
    a = HoneyBee(spectral_mode="bands")
    b = HoneyBee(spectral_mode="bands")
    c = HoneyBee(spectral_mode="bands", hsi_band_centers_nm=np.linspace(410.0, 690.0, 31))
    assert a._band_key == b._band_key != c._band_key


def test_n_outputs_follows_the_class_not_its_name():
    from animal_vision_b200.animals import Cat, Dog
    from animal_vision_b200.pipeline import HostBatchPipeline

    class NoWarp(Cat):
        ENABLE_FOV_WARP = False

    assert HostBatchPipeline.n_outputs(NoWarp()) == 2
    assert HostBatchPipeline.n_outputs(Cat()) == 2 and HostBatchPipeline.n_outputs(Dog()) == 1


def test_stale_library_with_other_header_is_refused(monkeypatch):
    from animal_vision_b200 import _abi, _build
    _abi.load()                                            # make sure a current library exists
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_build, "header_sha", lambda: "0123456789abcdef")
    monkeypatch.setattr(_build, "is_current", lambda: True)
    with pytest.raises(_abi.AvbError, match="another include/avb200.h"):
        _abi.load()


def test_uint16_frames_are_not_wrapped_to_uint8():
    """predict_torch.py:12-15 divides EVERY integer dtype by 255; 300 must not become 44."""
    import inspect
    from animal_vision_b200 import mstpp
    src = inspect.getsource(mstpp.MSTPlusPlus.predict_rgb_to_hsi)
    assert "astype(np.uint8)" not in src


# ------------------------------------------------------------------ GPU
def _canary_view(torch, shape, pad_rows=3, pad_cols=5, fill=0xA5):
    """A [N,H,W,3] uint8 view with padded row and frame strides inside a canary-filled buffer."""
    n, h, w, _ = shape
    big = torch.full((n + 2, h + 2 * pad_rows, (w + 2 * pad_cols) * 3), fill, dtype=torch.uint8, device="cuda")
    view = big[1:n + 1, pad_rows:pad_rows + h, pad_cols * 3:(pad_cols + w) * 3].unflatten(2, (w, 3))
    return big, view


def _canaries_intact(torch, big, view_shape, pad_rows=3, pad_cols=5, fill=0xA5):
    n, h, w, _ = view_shape
    mask = torch.ones_like(big, dtype=torch.bool)
    mask[1:n + 1, pad_rows:pad_rows + h, pad_cols * 3:(pad_cols + w) * 3] = False
    return bool((big[mask] == fill).all())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Rat", "Dog", "Squirrel", "Cow", "Panda", "HoneyBee", "Cat"])
@pytest.mark.parametrize("shape", [(2, 37, 53, 3), (1, 64, 131, 3), (3, 9, 17, 3)])
def test_outputs_never_write_outside_their_rows(name, shape):
    """Outputs live inside a larger canary-filled allocation (padded row and frame strides, odd sizes):
    every byte outside the H x W x 3 rows of each frame must survive the launch."""
    import torch
    import animal_vision_b200.animals as A
    sp = getattr(A, name)()
    n, h, w, _ = shape
    x = torch.from_numpy(np.stack([F.noise(h, w, s) for s in range(n)])).cuda()
    ref = sp.visualize_batch(x)                           # contiguous outputs: the reference result
    ref = ref if name == "Cat" else (ref[1],)
    bigs, views = zip(*[_canary_view(torch, shape) for _ in ref])
    in_big, in_view = _canary_view(torch, shape, fill=0x00)
    in_view.copy_(x)
    sp.visualize_batch(in_view, out=tuple(views) if name == "Cat" else views[0])
    torch.cuda.synchronize()
    for big, view, r in zip(bigs, views, ref):
        assert _canaries_intact(torch, big, shape), f"{name}: bytes outside the output rows were overwritten"
        assert torch.equal(view, r), f"{name}: strided output differs from the contiguous one"


@pytest.mark.gpu
def test_float_path_keeps_its_canaries():
    import torch
    from animal_vision_b200.animals import Dog, Cow
    from animal_vision_b200.engine import get_engine
    eng = get_engine()
    for sp in (Dog(), Cow()):
        x = torch.rand((2, 33, 47, 3), device="cuda")
        guard = 4096
        buf = torch.full((x.numel() + 2 * guard,), float("nan"), device="cuda")
        out = buf[guard:guard + x.numel()].view_as(x)
        tmp = torch.empty_like(x)
        sp._run_f32(eng, x, out, tmp, False)
        torch.cuda.synchronize()
        assert bool(torch.isnan(buf[:guard]).all()) and bool(torch.isnan(buf[guard + x.numel():]).all())
        assert not bool(torch.isnan(out).any())


@pytest.mark.gpu
def test_visualize_is_reentrant_across_threads():
    """Each calling thread stages through its own pinned buffers (engine.staging is keyed per thread)."""
    from animal_vision_b200.animals import Dog, HoneyBee
    frames = [F.natural(120, 200, s) for s in range(4)]
    expect = [(Dog().visualize(f)[1], HoneyBee().visualize(f)[1]) for f in frames]
    errs = []

    def work(i):
        try:
            for _ in range(6):
                d = Dog().visualize(frames[i])[1]
                b = HoneyBee().visualize(frames[i])[1]
                if not (np.array_equal(d, expect[i][0]) and np.array_equal(b, expect[i][1])):
                    errs.append(i)
        except Exception as e:   # noqa: BLE001
            errs.append(repr(e))
    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


@pytest.mark.gpu
def test_second_device_gets_its_own_smem_opt_in():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per device: K1 / K2 / K4 must launch on cuda:1 after cuda:0."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from animal_vision_b200.animals import Dog, Rat
    from animal_vision_b200.mstpp import MSTPlusPlus, synthetic_state_dict
    f = torch.from_numpy(np.stack([F.natural(64, 96)]))
    outs = []
    for d in (0, 1):
        x = f.to(f"cuda:{d}")
        outs.append((Rat().visualize_batch(x)[1].cpu(), Dog().visualize_batch(x)[1].cpu()))
        net = MSTPlusPlus(synthetic_state_dict(0), f"cuda:{d}")
        y = net.forward_nhwc(torch.rand(1, 40, 48, 3, generator=torch.Generator().manual_seed(1)).to(f"cuda:{d}"))
        outs[-1] += (y.cpu(),)
    for a, b in zip(*outs):
        assert torch.equal(a, b)
