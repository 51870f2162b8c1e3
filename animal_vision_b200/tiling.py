"""Spatial tiling of ONE large frame (e.g. 7680x4320) over the GPUs of a box -- the only place the
path has a real exchange step: each rank owns a horizontal band of rows; before the
blur it receives the `radius` rows above and below its band from its neighbours (NCCL send/recv
over NVLink through torch.distributed), runs the ordinary fused kernel on band + halo and keeps
its own rows.  The kernel reflects (REFLECT_101) only at the TRUE image border: inside the frame the
reflected rows of the extended buffer fall into the discarded halo, so the tiled result is
bit-identical to the single-GPU result.

Supported: the nine Gaussian-blur dichromat mammals (Dog, Bear, Lion, ...).  The row-dependent
species (Rat's per-row gain, the streak species' per-row sigma) need no halo at all and are sharded
as whole frames instead.  `get_normalized_image`'s frame-global branch (divide by 255 only
if the frame maximum exceeds 1, animals/animal_utils.py:41-50) is resolved with one tiny MAX
all-reduce.  Cat (vertical resampling of the centre zoom) and the UV species (global percentiles)
are not tiled: whole frames are sharded across GPUs instead (sharding.py)."""
from __future__ import annotations

from typing import Tuple

from . import tables
from ._abi import AVB_NORM_DIV255, AvbError
from .engine import get_engine
from .sharding import shard_range


def band_rows(H: int, rank: int, world: int) -> Tuple[int, int]:
    """[y0, y1) of the rows owned by `rank` (contiguous, sizes differ by <= 1)."""
    return shard_range(H, rank, world)


def halo_radius(species) -> int:
    """Rows a band needs from each neighbour: the Gaussian radius, 0 for row-independent species."""
    sigma = getattr(species, "SIGMA", None)
    if sigma is None or not hasattr(species, "ALPHA"):
        return 0
    return tables.gaussian_ksize(float(sigma)) // 2


def exchange_halos(band, radius: int, rank: int, world: int, group=None):
    """band: [1, h, W, 3] uint8 tensor of this rank's rows.  Returns (extended, top, bottom): the band
    with up to `radius` neighbour rows attached above / below, and how many were attached.
    Uses point-to-point sends over the process group (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if radius == 0 or world == 1:
        return band, 0, 0
    h = band.shape[1]
    if h < radius:
        raise AvbError(f"band of {h} rows is thinner than the blur radius {radius}: use fewer ranks")
    ops, top_buf, bot_buf = [], None, None
    if rank > 0:                                        # neighbour above: send my first rows, receive its last rows
        top_buf = torch.empty_like(band[:, :radius])
        ops.append(dist.P2POp(dist.isend, band[:, :radius].contiguous(), rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, top_buf, rank - 1, group))
    if rank < world - 1:                                # neighbour below
        bot_buf = torch.empty_like(band[:, :radius])
        ops.append(dist.P2POp(dist.isend, band[:, h - radius:].contiguous(), rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, bot_buf, rank + 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = ([top_buf] if top_buf is not None else []) + [band] + ([bot_buf] if bot_buf is not None else [])
    return torch.cat(parts, dim=1), (radius if top_buf is not None else 0), (radius if bot_buf is not None else 0)


def frame_divides_by_255(band, group=None) -> bool:
    """The frame-global branch of get_normalized_image, agreed on by all ranks (one MAX all-reduce)."""
    import torch.distributed as dist
    m = band.max().to(dtype=band.new_empty(()).float().dtype).reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return bool(m.item() > 1.0)


def visualize_band(species, band, H: int, rank: int, world: int, group=None):
    """Run `species` on this rank's band of a frame of `H` rows.  band: CUDA uint8 [1, h, W, 3] holding
    rows band_rows(H, rank, world).  Returns the band of the output frame (same shape)."""
    eng = get_engine(band.device)
    y0, y1 = band_rows(H, rank, world)
    if band.shape[1] != y1 - y0:
        raise AvbError(f"rank {rank} must hold rows [{y0}, {y1}) of the frame, got {band.shape[1]} rows")
    if not hasattr(species, "_matrix") or getattr(species, "SIGMA", None) is None:
        raise AvbError(f"{type(species).__name__} cannot be tiled spatially (row-dependent parameters or frame-global "
                       "resampling / statistics): shard whole frames instead")
    radius = halo_radius(species)
    ext, top, bot = exchange_halos(band, radius, rank, world, group)
    if not frame_divides_by_255(band, group):
        raise AvbError("frames whose maximum is <= 1 are not tiled (whole-frame path handles them)")
    out = eng.torch.empty_like(ext)
    taps = tables.gaussian_taps(tables.gaussian_ksize(species.SIGMA), species.SIGMA)
    eng.dichromat_blur(ext, out, species._matrix(), taps, norm=AVB_NORM_DIV255)
    return out[:, top:top + (y1 - y0)]
