"""Full-size checks (BASELINE configs: 1080p and 4K frames, the mixed-species batch): direct parity
against the oracle on a few whole frames, the reference's own output hashes from tests/golden
(1-LSB distance cannot be hashed, so the golden hash pins the ORACLE and the oracle pins the GPU),
and size-independent properties: batch invariance, determinism, stride invariance."""
import hashlib

import numpy as np
import pytest

import frames
from oracle import mammals as M
from oracle import uv

pytestmark = pytest.mark.gpu


def _lsb(got, ref, what, max_frac=0.02):
    d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    assert d.max() <= 1, f"{what}: max diff {d.max()} LSB"
    assert (d > 0).mean() <= max_frac, f"{what}: {(d > 0).mean():.4f} of bytes differ"


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_1080p_against_oracle_and_reference_hashes(golden_meta):
    """Config 1-3 inputs (default_rng(0) noise, 1080p).  The oracle must reproduce the hashes of the
    reference's own outputs; the CUDA path must sit within 1 LSB of the oracle."""
    from animal_vision_b200.animals import Cat, Dog, HoneyBee
    f = frames.noise(1080, 1920, 0)
    want = golden_meta["hashes"]["1080x1920"]
    assert _sha(f) == want["input"]
    ref_dog = M.mammal_visualize(f, "dog")[1]
    ref_h, ref_c = M.cat_visualize(f)
    ref_bee = uv.honeybee_visualize(f)[1]
    assert (_sha(ref_dog), _sha(ref_h), _sha(ref_c), _sha(ref_bee)) == (want["dog"], want["cat_human"], want["cat"], want["honeybee"])
    _lsb(Dog().visualize(f)[1], ref_dog, "dog 1080p")
    h, c = Cat().visualize(f)
    assert _sha(h) == want["cat_human"]                      # centre zoom is bit-exact: the reference's hash itself
    _lsb(c, ref_c, "cat 1080p")
    _lsb(HoneyBee().visualize(f)[1], ref_bee, "honeybee 1080p")


@pytest.mark.parametrize("species", ["Dog", "Cat", "HoneyBee"])
def test_4k_frame_against_oracle(species):
    """Config 5 frame size (3840x2160), natural-looking content (noise through a wide blur is flat)."""
    import animal_vision_b200.animals as A
    f = frames.natural(2160, 3840, seed=11)
    f[:, ::7] = frames.noise(2160, 3840, 3)[:, ::7]           # keep high-frequency columns too
    got = getattr(A, species)().visualize(f)
    if species == "Dog":
        _lsb(got[1], M.mammal_visualize(f, "dog")[1], "dog 4K")
    elif species == "Cat":
        ref_h, ref_c = M.cat_visualize(f)
        assert np.array_equal(got[0], ref_h)
        _lsb(got[1], ref_c, "cat 4K")
    else:
        _lsb(got[1], uv.honeybee_visualize(f)[1], "honeybee 4K")


def test_batch_invariance_determinism_and_strides():
    """A frame's result does not depend on what else is in the batch (all statistics are per frame),
    on being run twice, or on living inside a wider / offset buffer."""
    import torch
    import animal_vision_b200.animals as A
    fs = [frames.noise(360, 640, s) for s in range(4)] + [frames.natural(360, 640), frames.le1(360, 640), frames.constant(360, 640, 200)]
    batch = torch.from_numpy(np.stack(fs)).cuda()
    for name in ("Dog", "Squirrel", "Cow", "Rat", "Cat", "HoneyBee"):
        sp = getattr(A, name)()
        full = sp.visualize_batch(batch)
        full = full if name == "Cat" else (full[1],)
        again = sp.visualize_batch(batch)
        again = again if name == "Cat" else (again[1],)
        for a, b in zip(full, again):
            assert torch.equal(a, b), f"{name}: not deterministic"
        for i in (0, 4, 5, 6):
            one = sp.visualize_batch(batch[i:i + 1].clone())
            one = one if name == "Cat" else (one[1],)
            for a, b in zip(full, one):
                assert torch.equal(a[i], b[0]), f"{name}: frame {i} depends on its batch"
        wide = torch.zeros((7, 360, 700, 3), dtype=torch.uint8, device="cuda")
        wide[:, :, 13:653] = batch
        view = sp.visualize_batch(wide[:, :, 13:653])
        view = view if name == "Cat" else (view[1],)
        for a, b in zip(full, view):
            assert torch.equal(a, b), f"{name}: result depends on strides / alignment"


def test_mixed_species_round_robin_batch():
    """Config 5 in miniature: Dog/Cat/HoneyBee round-robin over a 12-frame video, frame-sharded plan."""
    import torch
    import animal_vision_b200.animals as A
    from animal_vision_b200 import sharding
    species = ("Dog", "Cat", "HoneyBee")
    video = [frames.noise(270, 480, s) for s in range(12)]
    outs = {}
    for rank in range(2):
        plan = sharding.shard_plan(len(video), rank, 2, species)
        for sp in species:
            idx = [i for i, s in plan if s == sp]
            res = getattr(A, sp)().visualize_batch(torch.from_numpy(np.stack([video[i] for i in idx])).cuda())
            for j, i in enumerate(idx):
                outs[i] = res[1][j].cpu().numpy()
    assert sorted(outs) == list(range(12))
    for i, f in enumerate(video):
        sp = species[i % 3]
        ref = {"Dog": lambda: M.mammal_visualize(f, "dog")[1], "Cat": lambda: M.cat_visualize(f)[1],
               "HoneyBee": lambda: uv.honeybee_visualize(f)[1]}[sp]()
        _lsb(outs[i], ref, f"frame {i} ({sp})")
