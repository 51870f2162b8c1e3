#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference in place (container only).

    python tools/make_golden.py            # needs /root/reference, cv2, torch

The reference ships no tests or golden vectors, so these fixtures -- outputs of the reference's
own `visualize()` / `MST_Plus_Plus.forward` on seeded and structured inputs -- are what pins the
oracle (tests/test_oracle_golden.py) and, on the GPU box where /root/reference does not exist,
the CUDA path (tests/test_gpu_*.py).  Inputs are regenerated from tests/frames.py, only outputs
(or their sha256 for the large frames) are stored.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import frames  # noqa: E402
import ref_loader as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SMALL = (72, 128)          # H, W of the stored-output cases
MAMMALS = ["dog", "bear", "lion", "tiger", "elephant", "fox", "wolf", "raccoon", "squirrel", "rat",
           "cow", "deer", "goat", "horse", "kangaroo", "sheep", "panda", "rabbit", "pig"]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def versions():
    import cv2
    import torch
    return {"numpy": np.__version__, "cv2": cv2.__version__, "torch": torch.__version__}


def main():
    os.makedirs(OUT, exist_ok=True)
    h, w = SMALL
    meta = {"versions": versions(), "small_hw": [h, w], "hashes": {}}

    # ---- mammals: every species on noise + natural; dog/cow/rat/panda on the full structured set
    store = {}
    for sp in MAMMALS:
        cls = R.species(sp, sp.capitalize())
        names = [n for n, _ in frames.parity_set(h, w)] if sp in ("dog", "cow", "rat", "panda", "squirrel") \
            else ["noise0", "natural", "le1"]
        for name, f in frames.parity_set(h, w):
            if name not in names:
                continue
            base, out = cls().visualize(f.copy())
            assert base.shape == f.shape and np.array_equal(base, f)
            store[f"{sp}/{name}"] = out
    np.savez_compressed(os.path.join(OUT, "mammals.npz"), **store)

    # ---- cat: both outputs on the full set
    Cat = R.species("cat", "Cat")
    store = {}
    for name, f in frames.parity_set(h, w):
        human, cat = Cat().visualize(f.copy())
        store[f"human/{name}"] = human
        store[f"cat/{name}"] = cat
    CatNoWarp = type("CatNoWarp", (Cat,), {"ENABLE_FOV_WARP": False})        # class switch, cat.py:21
    for name in ("noise0", "natural", "le1", "bars"):
        f = dict(frames.parity_set(h, w))[name]
        human, cat = CatNoWarp().visualize(f.copy())
        store[f"nowarp_human/{name}"] = human
        store[f"nowarp_cat/{name}"] = cat
    np.savez_compressed(os.path.join(OUT, "cat.npz"), **store)

    # ---- honeybee: default on the full set; other mappers / adaptation on noise + natural
    HB = R.species("honeybee", "HoneyBee")
    store = {}
    for name, f in frames.parity_set(h, w):
        store[f"opponent/white_patch/{name}"] = HB().visualize(f.copy())[1]
    for mode in ("falsecolor", "uv_purple_yellow", "falsecolor_uv_mixed"):
        for name in ("noise0", "natural", "ramps"):
            f = dict(frames.parity_set(h, w))[name]
            store[f"{mode}/white_patch/{name}"] = HB(mapping_mode=mode).visualize(f.copy())[1]
    for name in ("noise0", "natural"):
        f = dict(frames.parity_set(h, w))[name]
        store[f"opponent/gray_world/{name}"] = HB(adaptation="gray_world").visualize(f.copy())[1]
    M = np.array([[0.9, 0.1, 0.0], [0.0, 0.3, 0.8], [0.5, 0.5, 0.1]], np.float32)
    f = dict(frames.parity_set(h, w))["natural"]
    store["custom_matrix/white_patch/natural"] = HB(mapping_mode="custom_matrix", custom_matrix=M).visualize(f.copy())[1]
    np.savez_compressed(os.path.join(OUT, "honeybee.npz"), **store)

    # ---- fp32 intermediates of the reference for the <=1e-5 checks (natural frame)
    au = R.module("animals.animal_utils")
    uvh = R.module("uv_helpers")
    hb = R.species("honeybee", "HoneyBee")()
    f = frames.natural(h, w)
    lin = au.srgb_to_linear(au.get_normalized_image(f))
    dog_lin = (lin.reshape(-1, 3) @ au.collapse_LMS_matrix(0.58, 0.65).T).reshape(lin.shape)
    dog_blur = au.apply_acuity_blur(dog_lin, 3.5)
    lms = au.sRGB_to_LMS(lin.reshape(-1, 3)).reshape(lin.shape)
    hsi = R.classic_rgb_to_hsi()(uvh.to_float01(f), wavelengths=hb.lambdas)
    rad = hsi * uvh.D65_like(hb.lambdas).astype(hsi.dtype)[None, None, :]
    ubg = np.stack([np.tensordot(rad, c, axes=([2], [0])) for c in (hb.UV_curve, hb.Blue_curve, hb.Green_curve)], 2)
    np.savez_compressed(os.path.join(OUT, "intermediates.npz"),
                        lms=lms.astype(np.float32), dog_lin=dog_lin.astype(np.float32),
                        dog_blur=dog_blur.astype(np.float32), bee_ubg=ubg.astype(np.float32),
                        hsi_px=hsi[::9, ::16].astype(np.float32))

    # ---- hashes of large-frame outputs (inputs: default_rng(0) noise), cf. SURVEY.md 8c
    for (H, W) in ((270, 480), (1080, 1920)):
        f = frames.noise(H, W, 0)
        key = f"{H}x{W}"
        meta["hashes"][key] = {
            "input": sha(f),
            "dog": sha(R.species("dog", "Dog")().visualize(f)[1]),
            "cat_human": sha(Cat().visualize(f)[0]),
            "cat": sha(Cat().visualize(f)[1]),
            "honeybee": sha(HB().visualize(f)[1]),
        }
        print(key, meta["hashes"][key])

    # ---- MST++: the reference nn.Module with the oracle's synthetic weights loaded
    import torch
    from oracle import mstpp
    net = R.mstpp_module().MST_Plus_Plus().eval()
    net.load_state_dict(mstpp.make_weights(0))
    x = torch.rand(1, 3, 42, 52, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y = net(x)
    np.savez_compressed(os.path.join(OUT, "mstpp.npz"), y=y.numpy().astype(np.float32))
    meta["mstpp"] = {"weights": "oracle.mstpp.make_weights(0)", "input": "torch.rand(1,3,42,52, manual_seed(1))",
                     "mean": float(y.mean()), "std": float(y.std()), "absmax": float(y.abs().max())}
    # survey fingerprint: torch.manual_seed(0) module init, input 1x3x64x72 seed 1
    torch.manual_seed(0)
    net0 = R.mstpp_module().MST_Plus_Plus().eval()
    x0 = torch.rand(1, 3, 64, 72, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y0 = net0(x0)
    meta["mstpp_seed0_init"] = {"mean": float(y0.mean()), "std": float(y0.std()), "absmax": float(y0.abs().max())}

    with open(os.path.join(OUT, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
