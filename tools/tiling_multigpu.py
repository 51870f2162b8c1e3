"""visualize_band on real GPUs over NCCL: one 7680x4320 frame, each rank owns a band of rows, halo rows travel by
NCCL send/recv over NVLink, the result is compared with the whole-frame output computed on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/tiling_multigpu.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import animal_vision_b200.animals as A
from animal_vision_b200 import tiling

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, W = 4320, 7680
res = {"n_gpus": world, "frame": f"{W}x{H}"}
for name in ("Dog", "Squirrel"):
    sp = getattr(A, name)()
    full = torch.from_numpy(np.random.default_rng(7).integers(0, 256, (1, H, W, 3), dtype=np.uint8))
    y0, y1 = tiling.band_rows(H, rank, world)
    band = full[:, y0:y1].cuda()
    for _ in range(2):
        out = tiling.visualize_band(sp, band, H, rank, world)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = tiling.visualize_band(sp, band, H, rank, world)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gathered = [torch.empty((1, tiling.band_rows(H, r, world)[1] - tiling.band_rows(H, r, world)[0], W, 3), dtype=torch.uint8, device="cuda") for r in range(world)]
    dist.all_gather(gathered, out.contiguous())
    if rank == 0:
        tiled = torch.cat(gathered, dim=1)
        _, whole = sp.visualize_batch(full.cuda())
        d = (tiled.to(torch.int16) - whole.to(torch.int16)).abs()
        e0.record(); sp.visualize_batch(full.cuda()); e1.record(); torch.cuda.synchronize()
        res[name] = {"tiled_ms_max_over_ranks": float(ms.item()), "max_lsb_vs_whole_frame": int(d.max()), "bytes_differing": float((d > 0).float().mean()),
                     "gpix_per_s": H * W / float(ms.item()) / 1e6}
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
