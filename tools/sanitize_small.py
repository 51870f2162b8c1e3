"""Small end-to-end invocations of every kernel family (the command compute-sanitizer wraps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import animal_vision_b200.animals as A
import frames
from animal_vision_b200.mstpp import MSTPlusPlus
from animal_vision_b200.mstpp import synthetic_state_dict

for (h, w) in ((37, 53), (96, 400), (130, 1100)):
    fs = np.stack([frames.noise(h, w, 1), frames.natural(h, w), frames.le1(h, w)])
    batch = torch.from_numpy(fs).cuda()
    for name in ("Dog", "Squirrel", "Rat", "Cow", "Panda", "Cat", "HoneyBee"):
        getattr(A, name)().visualize_batch(batch)
    for mode in ("falsecolor", "uv_purple_yellow", "falsecolor_uv_mixed"):
        A.HoneyBee(mapping_mode=mode, blur_sigma_px=0.5).visualize_batch(batch)
    A.HoneyBee(spectral_mode="bands", blur_sigma_px=0.0, adaptation="gray_world").visualize_batch(batch)
    wide = torch.zeros((3, h, w + 9, 3), dtype=torch.uint8, device="cuda")
    for name in ("Dog", "Cow", "Cat", "HoneyBee"):
        getattr(A, name)().visualize_batch(wide[:, :, 5:5 + w])
net = MSTPlusPlus(synthetic_state_dict(0))
net(torch.rand(2, 3, 42, 52).cuda())
net.predict_rgb_to_hsi(np.random.default_rng(0).integers(0, 256, (45, 70, 3), dtype=np.uint8))
torch.cuda.synchronize()
print("sanitize_small ok")
