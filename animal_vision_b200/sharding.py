"""Frame-wise sharding of a video / frame batch over the GPUs of one box.

Frames are independent units -- every global statistic of the path (frame max, percentiles, MST++
attention statistics) is per frame -- so rank r simply owns a contiguous block of frame indices
(keeps decode order) and there is NO collective on the data path.  torch.distributed is used only
to agree on timings (max over ranks)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the contiguous block of `total` frames owned by `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def species_of_frame(index: int, species: Sequence[str]) -> str:
    """Round-robin species assignment by GLOBAL frame index (BASELINE configs[4]: mixed species)."""
    return species[index % len(species)]


def shard_plan(total: int, rank: int, world: int, species: Sequence[str]) -> List[Tuple[int, str]]:
    """(global frame index, species) for every frame this rank processes."""
    b, e = shard_range(total, rank, world)
    return [(i, species_of_frame(i, species)) for i in range(b, e)]


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over all ranks (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    if device is None:                   # NCCL reduces CUDA tensors only
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
