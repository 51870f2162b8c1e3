"""Scratch timing of individual species at a given size (device-resident frames, CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import animal_vision_b200.animals as A

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["dog"]
    H, W, N = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (2160, 3840, 20)))
    frames = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device="cuda")
    for name in names:
        sp = getattr(A, name)()
        out = [torch.empty_like(frames)]
        ms = timeit(lambda: sp.visualize_batch(frames))
        px = N * H * W
        print(f"{name}: {ms:.3f} ms / {N} frames {W}x{H}  -> {px/ms/1e6:.1f} Gpx/s, {6*px/ms/1e6:.0f} GB/s algorithmic, {N/ms*1e3:.0f} fps")
