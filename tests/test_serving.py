"""Host logic of the server-side batcher (animal_vision_b200/serving.py) and the split-compare compose
(renderers/video.py, 8f-3) on the CPU; the GPU tests run the same batcher against the real species."""
import threading
import time

import numpy as np
import pytest

import frames
from animal_vision_b200 import serving


def _fake_run(log):
    def run(key, batch):
        log.append((key, batch.shape))
        time.sleep(0.002)
        return (batch.astype(np.int16) + len(key)).clip(0, 255).astype(np.uint8)     # recognisable per key
    return run


def test_requests_from_many_threads_are_grouped_into_batches():
    log = []
    with serving.FrameBatcher(_fake_run(log), max_batch=8, max_delay_ms=30.0) as fb:
        fs = [frames.noise(12, 16, s) for s in range(20)]
        keys = ["dog" if i % 3 else "cat" for i in range(20)]
        futs = [None] * 20

        def client(i):
            futs[i] = fb.submit(fs[i], keys[i])
        ths = [threading.Thread(target=client, args=(i,)) for i in range(20)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        outs = [f.result(timeout=10) for f in futs]
    for i in range(20):                                    # every request gets ITS frame's result back
        assert np.array_equal(outs[i], (fs[i].astype(np.int16) + len(keys[i])).clip(0, 255).astype(np.uint8))
    assert sum(n for _, n in fb.batches) == 20 and len(fb.batches) < 20                  # batched, nothing lost
    assert all(shape[0] <= 8 for _, shape in log) and {k for k, _ in log} == {"dog", "cat"}


def test_shapes_are_not_mixed_and_human_passes_through():
    log = []
    with serving.FrameBatcher(_fake_run(log), max_batch=16, max_delay_ms=20.0) as fb:
        a, b = frames.noise(8, 8, 1), frames.noise(10, 6, 2)
        fa, fb_, fh = fb.submit(a, "dog"), fb.submit(b, "dog"), fb.submit(a, "human")
        assert fh.result() is a                                                       # utils.py:146-147
        assert fa.result(5).shape == a.shape and fb_.result(5).shape == b.shape
    assert sorted(s[1:] for _, s in log) == sorted([a.shape, b.shape])


def test_errors_reach_the_caller_and_the_loop_survives():
    calls = []

    def run(key, batch):
        calls.append(key)
        if key == "cow":
            raise RuntimeError("device lost")
        return batch
    with serving.FrameBatcher(run, max_delay_ms=1.0) as fb:
        f = frames.noise(4, 4, 0)
        with pytest.raises(RuntimeError):
            fb.submit(f, "cow").result(5)
        assert np.array_equal(fb.submit(f, "dog").result(5), f)
        with pytest.raises(KeyError):
            fb.submit(f, "unicorn").result(1)                                         # utils.py:192-193 "no case implemented"
        with pytest.raises(AssertionError):
            fb.submit(f.astype(np.float32), "dog").result(1)
    with pytest.raises(RuntimeError):
        fb.submit(f, "dog")


def test_server_keys_cover_the_reference_and_the_registry():
    import animal_vision_b200.animals as A
    from animal_vision_b200 import registry
    ref_keys = ["cat", "cow", "goat", "pig", "sheep", "dog", "rat", "horse", "rabbit", "panda", "squirrel", "elephant", "lion", "wolf",
                "fox", "bear", "raccoon", "deer", "kangaroo", "tiger", "honeybee"]                  # utils.py:145-191
    assert all(k in serving.SERVER_KEYS for k in ref_keys)
    assert {getattr(A, c) for c in serving.SERVER_KEYS.values()} == set(registry.animal_classes().values())


def test_process_image_round_trip_with_a_fake_device():
    import base64
    import cv2
    img = frames.natural(48, 64)
    ok, enc = cv2.imencode(".jpg", img)
    with serving.FrameBatcher(lambda key, batch: 255 - batch, max_delay_ms=1.0) as fb:
        uri = serving.process_image(enc.tobytes(), "dog", batcher=fb)
    assert uri.startswith("data:image/jpeg;base64,")                                  # utils.py:197-198
    back = cv2.imdecode(np.frombuffer(base64.b64decode(uri.split(",", 1)[1]), np.uint8), cv2.IMREAD_COLOR)
    decoded = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    assert back.shape == img.shape and np.abs(back.astype(int) - (255 - decoded.astype(int))).mean() < 6


def test_split_compare_matches_reference_golden(golden):
    import torch
    from animal_vision_b200.renderers.video import draw_split_labels, split_compare_batch
    g = golden("split_compare")
    for name, (h, w) in (("small", (120, 200)), ("tall", (300, 161))):
        a, b = frames.natural(h, w, 1), frames.noise(h, w, 2)
        ta, tb = torch.from_numpy(np.stack([a, a])), torch.from_numpy(np.stack([b, b]))
        out = split_compare_batch(ta, tb)
        assert torch.equal(out[0], out[1])
        assert np.array_equal(draw_split_labels(out[0].numpy().copy()), g[f"{name}/out"])           # renderers/video.py:198-245
        out = split_compare_batch(ta, tb, draw_seam=False)
        assert np.array_equal(draw_split_labels(out[1].numpy().copy(), "A", "Dog view"), g[f"{name}/out_noseam_labels"])
