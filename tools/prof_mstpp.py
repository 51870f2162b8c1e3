"""Run the MST++ forward a few times on one 482x512 patch (the command ncu wraps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from animal_vision_b200.mstpp import MSTPlusPlus
from animal_vision_b200.mstpp import synthetic_state_dict

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
net = MSTPlusPlus(synthetic_state_dict(0))
x = torch.rand(1, 482, 512, 3, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(reps):
    net.forward_nhwc(x)
torch.cuda.synchronize()
print("ok mstpp")
