"""Host-side plugin API (CPU): registry shape, constructor contracts, the batched frame feed."""
import numpy as np
import pytest


def test_registry_matches_reference_shape():
    from animal_vision_b200 import registry
    from animal_vision_b200.animals import Animal
    ch = registry.animal_choices()
    assert [c["name"] for c in ch][:3] == ["Cat", "Dog", "Sheep"] and len(ch) == 36          # utils.py:91-130
    assert [c["name"] for c in ch][20:24] == ["HoneyBee", "ReinDeer", "RatUV", "GoldFish"] and ch[-1]["name"] == "HummingBird"
    assert all(set(c) == {"name", "value"} and isinstance(c["value"], Animal) for c in ch)
    assert all(callable(getattr(c["value"], "visualize")) and callable(getattr(c["value"], "visualize_batch")) for c in ch)


def test_honeybee_constructor_contract():
    from animal_vision_b200.animals import HoneyBee
    b = HoneyBee()                                        # honeybee.py:47-66 defaults
    assert b.mapping_mode == "opponent" and b.adaptation == "white_patch" and b.blur_sigma_px == 0.2
    assert b.lambdas.shape == (31,) and b.lambdas[0] == 400.0 and b.lambdas[-1] == 700.0
    for c in (b.UV_curve, b.Blue_curve, b.Green_curve):
        assert abs(float(c.sum()) - 1.0) < 1e-5
    with pytest.raises(ValueError):
        HoneyBee(mapping_mode="nope")
    with pytest.raises(ValueError):
        HoneyBee(adaptation="nope")
    with pytest.raises(AssertionError):
        HoneyBee(mapping_mode="custom_matrix")           # honeybee.py:153-156 asserts a 3x3 matrix


def test_bad_frames_assert_before_any_gpu_work():
    from animal_vision_b200.animals import Cat, Dog, HoneyBee
    for sp in (Dog(), Cat(), HoneyBee()):
        with pytest.raises(AssertionError):
            sp.visualize(np.zeros((4, 4), np.uint8))      # dog.py:33 / cat.py:24 / honeybee.py:102-103
        with pytest.raises(AssertionError):
            sp.visualize(np.zeros((4, 4, 4), np.uint8))


class _FakeVideo:
    """get_image() contract of renderers/video.py:82-96: RGB uint8 HxWx3 or None at end of stream."""

    def __init__(self, n, h=6, w=8):
        self.frames = [np.full((h, w, 3), i, np.uint8) for i in range(n)]
        self.i = 0

    def get_image(self):
        if self.i >= len(self.frames):
            return None
        self.i += 1
        return self.frames[self.i - 1]


def test_batch_feed_gathers_frames_in_order():
    from animal_vision_b200.renderers.video import BatchFeed
    feed = BatchFeed(_FakeVideo(7), batch=3, pinned=False)
    sizes, seen = [], []
    while (b := feed.get_batch()) is not None:
        sizes.append(b.shape[0])
        seen += [int(f[0, 0, 0]) for f in b]
        assert b.shape[1:] == (6, 8, 3) and b.dtype == np.uint8
    assert sizes == [3, 3, 1] and seen == list(range(7))


@pytest.mark.gpu
def test_species_batches_on_concurrent_streams_match_serial_results():
    """Independent batches may be issued on different CUDA streams (bench.py does): every entry point is
    stream ordered and the engine keeps its scratch (normalisation flags, K3 workspace) per stream."""
    import torch
    import animal_vision_b200.animals as A
    rng = np.random.default_rng(21)
    frames = {name: torch.from_numpy(rng.integers(0, 256, (3, 270, 480, 3), dtype=np.uint8)).cuda()
              for name in ("Dog", "Cat", "HoneyBee", "Cow")}
    frames["Dog"][1] = frames["Dog"][1] // 200          # a frame whose maximum is <= 1: the other normalisation branch
    species = {name: getattr(A, name)() for name in frames}
    serial = {}
    for name, sp in species.items():
        base, out = sp.visualize_batch(frames[name])
        serial[name] = (base.clone() if name == "Cat" else None, out.clone())
    torch.cuda.synchronize()
    streams = {name: torch.cuda.Stream() for name in frames}
    got = {}
    for _ in range(3):                                   # a few rounds so that launches really interleave
        for name, sp in species.items():
            streams[name].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(streams[name]):
                got[name] = sp.visualize_batch(frames[name])
    torch.cuda.synchronize()
    for name in frames:
        base, out = got[name]
        assert torch.equal(out, serial[name][1]), name
        if name == "Cat":
            assert torch.equal(base, serial[name][0]), name


def test_uv_species_constructor_contract():
    """Keyword-only constructors with the reference's parameter names and defaults (reindeer.py:41-66 and siblings);
    tools/make_golden_uv.py checks the defaults against the reference's signatures when the goldens are generated."""
    import animal_vision_b200.animals as A
    r = A.Reindeer()
    assert r.hsi_scale == 0.25 and r.uv_band == (300.0, 410.0) and r.panorama_scale == 1.3 and r.N_OUTPUTS == 2
    assert r.lambdas.dtype == np.float32 and r.lambdas.shape == (81,) and r.lambdas[0] == 300.0 and r.lambdas[-1] == 700.0
    assert A.RatUV().lambdas.shape == (129,) and A.RatUV().lambdas[0] == 320.0                  # rat_uv.py:48
    assert A.Goldfish(uv_boost=2.0).uv_boost == 2.0 and A.Goldfish().uv_boost == 3.0
    assert len(A.MantisShrimp().bands) == 10 and A.MantisShrimp().bands[-1] == (610.0, 680.0)  # mantis_shrimp.py:49-60
    assert A.Morpho(mosaic_downscale=0.01).mosaic_downscale == 0.15                             # morpho.py:59 clip
    with pytest.raises(TypeError):
        A.Kestrel(no_such_parameter=1)
    with pytest.raises(TypeError):
        A.Reindeer(0.5)                                                                         # keyword-only, as the reference
    with pytest.raises(AssertionError):
        A.Dragonfly(lambdas=np.arange(5))                                                       # dragonfly.py:100
    for cls in (A.Reindeer, A.Anableps, A.Hummingbird):
        with pytest.raises(AssertionError):
            cls().visualize(np.zeros((4, 4), np.uint8))                                         # reindeer.py:82-83
