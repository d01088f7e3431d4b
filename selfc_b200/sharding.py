"""GOP sharding across ranks and the final metric gather.

The rescaling path shards by independent units: every dependency is closed inside one 7-frame GOP (zero-padded
temporal convs, clip-local GlobalAgg).  Units are dealt to ranks as contiguous blocks; there is NO collective on the
data path.  `gather_metrics` is the only communication (one small all_gather at the end of an evaluation), and works
on any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

GOP = 7


def gop_indices(frames: int, gop: int = GOP) -> List[Tuple[List[int], int]]:
    """Frame indices of each GOP of a clip: full GOPs, then a tail padded with copies of the last frame, as
    models/SelfC_model.py:204-209 does.  Returns [(indices, n_real_frames), ...]."""
    out = []
    for g0 in range(0, frames, gop):
        ids = list(range(g0, min(frames, g0 + gop)))
        real = len(ids)
        ids += [frames - 1] * (gop - real)
        out.append((ids, real))
    return out


def partition(n_units: int, world: int, rank: int) -> range:
    """Contiguous block of unit indices owned by `rank` (sizes differ by at most one; covers 0..n_units-1 exactly once)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank {world}/{rank}")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_metrics(local: torch.Tensor, counts: Sequence[int] | None = None) -> torch.Tensor:
    """all_gather of per-unit metric rows [n_local, k] -> [n_total, k] in rank order (ragged n_local allowed).
    Without an initialised process group this is the identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes) if sizes else 0
    padded = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)
