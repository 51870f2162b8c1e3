"""Time the MST++ forward (CUDA events) and print the parity margins; optionally per-kernel split."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from animal_vision_b200 import _abi
from animal_vision_b200.mstpp import MSTPlusPlus
from animal_vision_b200.mstpp import synthetic_state_dict

FLOP_PER_PATCH = 169.2e9     # SURVEY.md 8a-19, 482x512, 2 x MAC, unpadded channel counts


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (482, 512)
    check = "--check" in sys.argv
    sd = synthetic_state_dict(0)
    net = MSTPlusPlus(sd)
    x = torch.rand(n, h, w, 3, generator=torch.Generator().manual_seed(1)).cuda()
    for _ in range(3):
        y = net.forward_nhwc(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        y = net.forward_nhwc(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    import time
    t0 = time.perf_counter()
    for _ in range(iters):
        y = net.forward_nhwc(x)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host enqueue time {1e3 * (t1 - t0) / iters:.3f} ms/forward (asynchronous part of the call)")
    scale = (h * w) / (482 * 512)
    print(f"MST++ {n} x {h}x{w}: {ms:.3f} ms/forward, {n / ms * 1e3:.1f} patch/s, {FLOP_PER_PATCH * scale * n / ms / 1e9:.1f} TFLOP/s algorithmic")
    lib = _abi.load()
    lib.avb_profile_begin()
    net.forward_nhwc(x)
    cap = 4096
    names = C.create_string_buffer(cap * 48)
    msbuf = (C.c_float * cap)()
    nrec = lib.avb_profile_end(names, 48, msbuf, cap)
    agg = {}
    for i in range(nrec):
        nm = names.raw[i * 48:(i + 1) * 48].split(b"\0", 1)[0].decode()
        a = agg.setdefault(nm, [0.0, 0])
        a[0] += msbuf[i]; a[1] += 1
    tot = sum(v[0] for v in agg.values())
    for nm, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {nm:22s} {t:8.3f} ms  {c:4d} launches  {100 * t / tot:5.1f}%")
    print(f"  sum of kernels {tot:.3f} ms, {nrec} launches")
    if "--seq" in sys.argv:
        for i in range(min(nrec, 64)):
            nm = names.raw[i * 48:(i + 1) * 48].split(b"\0", 1)[0].decode()
            print(f"    {i:3d} {nm:22s} {msbuf[i] * 1e3:8.1f} us")
    for parts in (2, 4):
        if n >= parts:
            for _ in range(2):
                y2 = net.forward_nhwc_streams(x, parts)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                y2 = net.forward_nhwc_streams(x, parts)
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / iters
            print(f"  {parts} streams: {ms2:.3f} ms/forward, {n / ms2 * 1e3:.1f} patch/s, bitwise equal to the single-stream output: {bool(torch.equal(y2, y))}")
    if check:
        xs = x[:1].cpu().permute(0, 3, 1, 2)
        from oracle import mstpp as O
        ref = O.forward(xs, sd).permute(0, 2, 3, 1).numpy()
        got = y[:1].cpu().numpy()
        d = np.abs(got - ref)
        print(f"  parity: max-abs rel {d.max() / np.abs(ref).max():.3e}, rel-L2 {np.linalg.norm(got - ref) / np.linalg.norm(ref):.3e}")


if __name__ == "__main__":
    main()
