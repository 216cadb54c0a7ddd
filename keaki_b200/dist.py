"""Multi-GPU sharding of the hot path: one process per GPU, `torch.distributed`.

  * MSM / commit: point-range shards.  Rank k owns SRS points [lo_k, hi_k) and the matching scalars, computes
    one partial sum (kb_msm_g1), the partials (64 B + flag each) are exchanged with ONE all_gather and every
    rank adds them on its GPU (kb_g1_sum).  There is no other collective on the path.
  * encrypt / decrypt / openings: independent per index -> contiguous index ranges, no collective.

The compute backend is any object with `msm_g1(scalars, n=, first=)` and `g1_sum(pts_xy, inf)` (a
`keaki_b200.Context`); the CPU tests inject an oracle-backed stand-in to exercise the exchange logic under gloo."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) for `rank`; the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_partial(xy: np.ndarray, inf: int) -> np.ndarray:
    """17 x int32: 16 coordinate limbs + infinity flag (int32 because NCCL/gloo tensors here are signed)."""
    out = np.zeros(17, np.uint32)
    out[:16] = xy
    out[16] = 1 if inf else 0
    return out.view(np.int32)


def unpack_partials(buf: np.ndarray):
    g = np.ascontiguousarray(buf).view(np.uint32).reshape(-1, 17)
    return np.ascontiguousarray(g[:, :16]), np.ascontiguousarray(g[:, 16].astype(np.uint8))


def sharded_commit(ctx, local_scalars, first: int, group=None, device=None):
    """Commit to a polynomial whose coefficients are sharded by point range: this rank holds
    `local_scalars` for SRS points [first, first + len).  Returns (xy, inf) of the full commitment
    on every rank."""
    import torch
    import torch.distributed as dist

    n_local = local_scalars.shape[0]
    xy, inf = ctx.msm_g1(local_scalars, n=n_local, first=first)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return xy, inf
    world = dist.get_world_size(group)
    part = torch.from_numpy(pack_partial(xy, inf).copy())
    if device is not None:
        part = part.to(device)
    gathered = torch.empty(world * 17, dtype=torch.int32, device=part.device)
    dist.all_gather_into_tensor(gathered, part, group=group)
    pts, infs = unpack_partials(gathered.cpu().numpy())
    return ctx.g1_sum(pts, infs)
