import sys, os
sys.path.insert(0, os.getcwd())
import torch
from animal_vision_b200.mstpp import MSTPlusPlus
from animal_vision_b200.mstpp import synthetic_state_dict
net = MSTPlusPlus(synthetic_state_dict(0))
for nb, parts in ((4, 4), (8, 4), (8, 8), (16, 8)):
    x = torch.rand(nb, 482, 512, 3, generator=torch.Generator().manual_seed(1)).cuda()
    for _ in range(3): net.forward_nhwc_streams(x, parts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): net.forward_nhwc_streams(x, parts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(nb, parts, round(ms, 2), "ms", round(nb / ms * 1e3, 1), "patch/s")
