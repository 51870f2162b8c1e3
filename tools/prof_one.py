"""Run one or more species (comma separated) a few times on device-resident noise frames (the command ncu wraps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import animal_vision_b200.animals as A

names = sys.argv[1].split(",")
H, W, N = (int(v) for v in sys.argv[2:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
frames = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device="cuda")
for name in names:
    sp = getattr(A, name)()
    for _ in range(reps):
        sp.visualize_batch(frames)
torch.cuda.synchronize()
print("ok", names)
