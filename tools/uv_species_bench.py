"""Device-resident time of every UV species on 1080p uint8 frames (CUDA events), with the launch count and per-kernel
breakdown from the library's profiling hook.   python tools/uv_species_bench.py [batch] [H W]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal_vision_b200 import _abi
from animal_vision_b200.engine import get_engine
import animal_vision_b200.animals as A

NAMES = ["Reindeer", "RatUV", "Goldfish", "Damselfish", "Anableps", "Anchovy", "Guppy", "Morpho", "Heliconius", "Pieris",
         "MantisShrimp", "Kestrel", "JumpingSpider", "Dragonfly", "Hummingbird"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1080, 1920)
lib = _abi.load()
eng = get_engine()
fr = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
res = {}
for name in NAMES:
    sp = getattr(A, name)()
    for _ in range(2): sp.visualize_batch(fr)
    torch.cuda.synchronize()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): sp.visualize_batch(fr)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    launches = (eng.launches - l0) // 5
    lib.avb_profile_begin()
    sp.visualize_batch(fr)
    names = C.create_string_buffer(1024 * 48); t = (C.c_float * 1024)()
    k = lib.avb_profile_end(names, 48, t, 1024)
    agg = {}
    for i in range(k):
        nm = names.raw[i * 48:(i + 1) * 48].split(b"\0", 1)[0].decode()
        agg[nm] = agg.get(nm, 0.0) + t[i]
    top = dict(sorted(((a, round(b, 3)) for a, b in agg.items()), key=lambda kv: -kv[1])[:6])
    res[name] = {"ms_per_frame": round(ms / n, 3), "mpix_per_s": round(n * H * W / ms / 1e3, 1), "launches": launches, "kernel_ms_sum": round(sum(agg.values()), 3), "top": top}
    print(name, json.dumps(res[name]), flush=True)
