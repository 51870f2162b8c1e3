"""Per-kernel CUDA-event times of HoneyBee on 20 4K frames (uses the library's profiling hook)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal_vision_b200 import _abi
import animal_vision_b200.animals as A
name = sys.argv[1] if len(sys.argv) > 1 else "HoneyBee"
lib = _abi.load()
fr = torch.randint(0, 256, (20, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
sp = getattr(A, name)()
for _ in range(3): sp.visualize_batch(fr)
torch.cuda.synchronize()
lib.avb_profile_begin()
for _ in range(3): sp.visualize_batch(fr)
names = C.create_string_buffer(256 * 48); ms = (C.c_float * 256)()
n = lib.avb_profile_end(names, 48, ms, 256)
agg = {}
for i in range(n):
    nm = names.raw[i * 48:(i + 1) * 48].split(b"\0", 1)[0].decode()
    agg[nm] = agg.get(nm, 0.0) + ms[i] / 3
print(os.environ.get("AVB_NVCC_EXTRA", "default"), {k: round(v, 3) for k, v in agg.items() if v > 0.05})
