"""Copy-only ceiling of the end-to-end leg: every rank moves what bench.py's e2e step moves (1.49 GB pinned H2D + 1.99 GB
D2H per step for 60 4K frames: Dog + Cat x2 + HoneyBee outputs) with NO kernels, all ranks concurrently.  One
cudaMemcpyAsync per chunk on its own stream per direction.  Prints one JSON line on rank 0: aggregate GB/s per direction
and the Gpix/s the e2e pipeline could reach if the kernels were free -- bench.py's e2e value divided by this is the
fraction of the host / PCIe ceiling the pipeline reaches.

    python tools/copy_probe.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/copy_probe.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

H, W, FRAMES = 2160, 3840, 60
CHUNK = int(os.environ.get("PROBE_CHUNK", "2"))          # frames per copy, as bench.py --chunk


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fb = H * W * 3
    h_in = torch.empty((FRAMES, fb), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((FRAMES * 4 // 3, fb), dtype=torch.uint8).pin_memory()          # 20 Dog + 40 Cat + 20 HoneyBee output frames
    d_in = torch.empty_like(h_in, device="cuda")
    d_out = torch.empty_like(h_out, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def step(do_in=True, do_out=True):
        if do_in:
            with torch.cuda.stream(s_in):
                for a in range(0, FRAMES, CHUNK):
                    d_in[a:a + CHUNK].copy_(h_in[a:a + CHUNK], non_blocking=True)
        if do_out:
            with torch.cuda.stream(s_out):
                for a in range(0, h_out.shape[0], CHUNK):
                    h_out[a:a + CHUNK].copy_(d_out[a:a + CHUNK], non_blocking=True)
        s_in.synchronize(); s_out.synchronize()

    res = {}
    for name, kw in (("both", {}), ("h2d_only", {"do_out": False}), ("d2h_only", {"do_in": False})):
        for _ in range(2):
            step(**kw)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            step(**kw)
        dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t.item())
    if rank == 0:
        gb_in, gb_out = FRAMES * fb / 1e9, h_out.shape[0] * fb / 1e9
        line = {"n_gpus": world, "chunk_frames": CHUNK, "h2d_gb_per_rank": gb_in, "d2h_gb_per_rank": gb_out,
                "both_ms": res["both"] * 1e3, "h2d_only_gbs_aggregate": world * gb_in / res["h2d_only"],
                "d2h_only_gbs_aggregate": world * gb_out / res["d2h_only"],
                "both_gbs_aggregate": world * (gb_in + gb_out) / res["both"],
                "e2e_ceiling_gpix_per_s": world * FRAMES * H * W / res["both"] / 1e9,
                "cpu_affinity": len(os.sched_getaffinity(0))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
