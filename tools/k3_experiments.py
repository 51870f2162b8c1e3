"""K3 routes of VERDICT r1 item 4, measured: (a) frame groups small enough to stay L2-resident across the passes
(20 frames as 10 calls of 2 / 20 calls of 1 instead of one call of 20); per-kernel times from the profiling hook."""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal_vision_b200 import _abi
import animal_vision_b200.animals as A
lib = _abi.load()
fr = torch.randint(0, 256, (20, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
out = torch.empty_like(fr)
sp = A.HoneyBee()
res = {"flags": os.environ.get("AVB_NVCC_EXTRA", "")}
for group in (20, 4, 2, 1):
    def run():
        for a in range(0, 20, group):
            sp.visualize_batch(fr[a:a + group], out[a:a + group])
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    lib.avb_profile_begin(); run()
    names = C.create_string_buffer(2048 * 48); t = (C.c_float * 2048)()
    k = lib.avb_profile_end(names, 48, t, 2048)
    agg = {}
    for i in range(k):
        nm = names.raw[i * 48:(i + 1) * 48].split(b"\0", 1)[0].decode()
        agg[nm] = agg.get(nm, 0.0) + t[i]
    res[f"group{group}"] = {"ms_per_20_frames": round(e0.elapsed_time(e1) / 5, 3), "kernels": {a: round(b, 3) for a, b in agg.items() if b > 0.02}}
print(json.dumps(res))
