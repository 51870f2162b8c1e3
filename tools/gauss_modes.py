"""A/B of the K2 Gaussian kernels: AVB_GAUSS_MMA=0 (CUDA cores), 1 (tensor cores, hi+lo f16), 2 (tensor cores, single f16).
For each mode (own process: the mode is read once): differing-byte fraction and max |diff| against the oracle over
the structured parity set at several shapes, and the time of 20 4K frames.   python tools/gauss_modes.py [child mode]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def child():
    import numpy as np, torch
    import frames
    from oracle import mammals as M
    import animal_vision_b200.animals as A
    res = {"mode": os.environ.get("AVB_GAUSS_MMA", "default"), "parity": {}, "ms": {}}
    for name in ("dog", "squirrel", "bear", "raccoon", "wolf"):
        sp = A.MAMMALS[name]()
        worst, wfrac, tot, dif = 0, 0.0, 0, 0
        for (h, w) in ((270, 480), (61, 67), (200, 1100), (37, 1), (540, 960)):
            for case, f in frames.parity_set(h, w):
                ref = M.mammal_visualize(f, name)[1]
                out = sp.visualize(f)[1]
                d = np.abs(out.astype(np.int16) - ref.astype(np.int16))
                worst = max(worst, int(d.max())); wfrac = max(wfrac, float((d > 0).mean()))
                tot += d.size; dif += int((d > 0).sum())
                if d.max() > 1 or (d > 0).mean() > 0.004:
                    print(f"  !! {name} {case} {h}x{w}: max {d.max()} frac {(d > 0).mean():.4f}", flush=True)
        res["parity"][name] = {"max_lsb": worst, "worst_frame_frac": round(wfrac, 5), "overall_frac": round(dif / tot, 6)}
    fr = torch.randint(0, 256, (20, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
    for name in ("Dog", "Squirrel", "Bear", "Raccoon"):
        sp = getattr(A, name)()
        for _ in range(3): sp.visualize_batch(fr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): sp.visualize_batch(fr)
        e1.record(); torch.cuda.synchronize()
        res["ms"][name] = round(e0.elapsed_time(e1) / 10, 4)
    print("RESULT " + json.dumps(res), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for mode in (sys.argv[1:] or ["0", "1", "2"]):
            env = dict(os.environ, AVB_GAUSS_MMA=mode)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
