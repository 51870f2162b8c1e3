"""GPU parity: the CUDA path (through the C-ABI) against the oracle and the reference's golden
vectors.  Tolerance from BASELINE.json north_star: <= 1 LSB on uint8 output."""
import numpy as np
import pytest

import frames
from oracle import mammals as M

pytestmark = pytest.mark.gpu

GAUSS = ["dog", "bear", "lion", "tiger", "elephant", "fox", "wolf", "raccoon", "squirrel"]
STREAK = ["cow", "deer", "goat", "horse", "kangaroo", "sheep", "panda", "rabbit", "pig"]


def _species(name):
    import animal_vision_b200.animals as A
    return A.MAMMALS[name]()


def _cmp(got, ref, what, max_frac=0.02):
    assert got.shape == ref.shape and got.dtype == ref.dtype == np.uint8, what
    d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    assert d.max() <= 1, f"{what}: max diff {d.max()} LSB"
    # (frames of a few bytes: a single 1-LSB byte is already several per cent)
    assert (d > 0).sum() <= max(max_frac * d.size, 2), f"{what}: {(d > 0).mean():.4f} of bytes differ by 1 LSB"


@pytest.mark.parametrize("name", GAUSS + ["rat"] + STREAK)
def test_against_golden(name, golden, golden_meta):
    h, w = golden_meta["small_hw"]
    fr = dict(frames.parity_set(h, w))
    g = golden("mammals")
    sp = _species(name)
    for key, ref in g.items():
        s, case = key.split("/")
        if s != name:
            continue
        base, out = sp.visualize(fr[case])
        assert base is fr[case]
        _cmp(out, ref, key)


@pytest.mark.parametrize("name", ["dog", "squirrel", "raccoon", "rat", "cow", "panda", "pig"])
@pytest.mark.parametrize("hw", [(61, 67), (8, 200), (270, 480), (1, 5), (37, 1), (40, 1100)])
def test_against_oracle_odd_shapes(name, hw):
    h, w = hw
    sp = _species(name)
    for case, f in frames.parity_set(h, w):
        _, ref = M.mammal_visualize(f, name)
        _, out = sp.visualize(f)
        _cmp(out, ref, f"{name}/{case}/{h}x{w}", max_frac=0.05)


def test_dog_1080p_batch_and_strides():
    import torch
    from animal_vision_b200.animals import Dog
    f0, f1 = frames.noise(1080, 1920, 0), frames.natural(1080, 1920)
    refs = [M.mammal_visualize(f, "dog")[1] for f in (f0, f1)]
    batch = torch.from_numpy(np.stack([f0, f1])).cuda()
    base, out = Dog().visualize_batch(batch)
    assert base is batch
    for i in range(2):
        _cmp(out[i].cpu().numpy(), refs[i], f"dog 1080p frame {i}")
    # non-contiguous rows: a view into a wider buffer
    wide = torch.zeros((2, 1080, 2000, 3), dtype=torch.uint8, device="cuda")
    wide[:, :, 40:1960] = batch
    view = wide[:, :, 40:1960]
    _, out2 = Dog().visualize_batch(view)
    assert torch.equal(out2, out)


def test_streak_1080p_batch_and_strides():
    import torch
    from animal_vision_b200.animals import Panda, Sheep
    f0, f1 = frames.noise(1080, 1920, 1), frames.natural(1080, 1920)
    batch = torch.from_numpy(np.stack([f0, f1])).cuda()
    for cls, name in ((Sheep, "sheep"), (Panda, "panda")):
        refs = [M.mammal_visualize(f, name)[1] for f in (f0, f1)]
        _, out = cls().visualize_batch(batch)
        for i in range(2):
            _cmp(out[i].cpu().numpy(), refs[i], f"{name} 1080p frame {i}")
        wide = torch.zeros((2, 1080, 2000, 3), dtype=torch.uint8, device="cuda")
        wide[:, :, 39:1959] = batch                      # unaligned rows: scalar store path
        _, out2 = cls().visualize_batch(wide[:, :, 39:1959])
        assert torch.equal(out2, out)


def test_host_batch_pipeline_matches_device_path():
    """e2e route of bench.py: pinned host frames -> H2D || kernels || D2H -> pinned host frames."""
    import torch
    from animal_vision_b200.animals import Cat, Dog, HoneyBee
    from animal_vision_b200.pipeline import HostBatchPipeline
    fs = np.stack([frames.noise(144, 256, s) for s in range(7)])
    host = torch.from_numpy(fs).pin_memory()
    pipe = HostBatchPipeline(chunk_frames=3)
    jobs, outs = [], {}
    for sp in (Dog(), Cat(), HoneyBee()):
        o = pipe.pinned_like(host, pipe.n_outputs(sp))
        outs[type(sp).__name__] = (sp, o)
        jobs.append((sp, host, o))
    h2d, d2h = pipe.run(jobs)
    assert h2d == 3 * fs.nbytes and d2h == 4 * fs.nbytes
    dev = host.cuda()
    for name, (sp, o) in outs.items():
        ref = sp.visualize_batch(dev)
        ref = ref if name == "Cat" else (ref[1],)
        for a, b in zip(o, ref):
            assert torch.equal(a, b.cpu()), name


def test_full_rank_matrix_takes_three_plane_path():
    """avb_dichromat_blur_u8 accepts any 3x3; species matrices are rank 2 (two-plane kernel), a
    generic full-rank matrix must give the same answer through the three-plane instantiation."""
    import torch
    from animal_vision_b200 import tables
    from animal_vision_b200.engine import get_engine
    from oracle import colorimetry as C
    from oracle import cvops as V
    eng = get_engine()
    T = np.array([[0.8, 0.15, 0.05], [0.1, 0.7, 0.2], [-0.05, 0.25, 0.8]], np.float32)
    assert abs(np.linalg.det(T.astype(np.float64))) > 0.1
    for sigma in (0.7, 2.0, 3.5):
        taps = tables.gaussian_taps(tables.gaussian_ksize(sigma), sigma)
        for f in (frames.noise(200, 300, 2), frames.natural(200, 300), frames.bars(131, 517)):
            lin = C.apply_matrix(C.decode_srgb(C.normalize_frame(f)), T)
            ref = C.encode_tail(V.gaussian_blur(lin, sigma), np.uint8)
            d_in = torch.from_numpy(f[None]).cuda()
            d_out = torch.empty_like(d_in)
            eng.dichromat_blur(d_in, d_out, T, taps)
            _cmp(d_out[0].cpu().numpy(), ref, f"full-rank sigma={sigma}", max_frac=0.03)
