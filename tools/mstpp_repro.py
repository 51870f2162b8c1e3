import sys, os
sys.path.insert(0, os.getcwd())
import torch
from animal_vision_b200.mstpp import MSTPlusPlus
from animal_vision_b200.mstpp import synthetic_state_dict
net = MSTPlusPlus(synthetic_state_dict(0))
x = torch.rand(4, 482, 512, 3, generator=torch.Generator().manual_seed(1)).cuda()
y1 = net.forward_nhwc(x).clone(); y2 = net.forward_nhwc(x).clone(); y3 = net.forward_nhwc_streams(x, 4).clone()
torch.cuda.synchronize()
m = y1.abs().max()
print("single vs single:", float((y1 - y2).abs().max() / m), " single vs streams:", float((y1 - y3).abs().max() / m))
print("per patch single-vs-streams:", [float((y1[i] - y3[i]).abs().max() / m) for i in range(4)])
