"""A/B of the opponent mapper's two map routes (planes of the hist pass vs the second walk): identical bytes, per-kernel times.
Run twice (AVB_UV_NO_PLANE_MAP unset / =1); each run prints a digest of the outputs and the kernel times."""
import ctypes as C, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animal_vision_b200 import _abi
import animal_vision_b200.animals as A

lib = _abi.load()
g = torch.Generator().manual_seed(5)
digests = []
for shape in ((3, 270, 480), (2, 1080, 1920), (1, 123, 236), (20, 2160, 3840)):
    fr = torch.randint(0, 256, (*shape, 3), dtype=torch.uint8, generator=g).cuda()
    fr[0, : shape[1] // 2] //= 3
    for kw in ({}, {"adaptation": "gray_world"}, {"blur_sigma_px": None}, {"blur_sigma_px": 0.6}, {"mapping_mode": "falsecolor"},
               {"mapping_mode": "uv_purple_yellow"}, {"mapping_mode": "falsecolor_uv_mixed"}):
        try:
            sp = A.HoneyBee(**kw)
        except TypeError:
            continue
        out = sp.visualize_batch(fr)
        out = out[-1] if isinstance(out, (tuple, list)) else out
        digests.append(hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:12])
print("route", "walk" if os.environ.get("AVB_UV_NO_PLANE_MAP") == "1" else "planes", "digests", " ".join(digests))
for mode in ("opponent", "falsecolor", "uv_purple_yellow", "falsecolor_uv_mixed"):
    sp = A.HoneyBee(mapping_mode=mode)
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
    print(mode, {k: round(v, 3) for k, v in agg.items() if v > 0.02}, "total", round(sum(agg.values()), 3))
