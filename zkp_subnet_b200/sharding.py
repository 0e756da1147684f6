"""Multi-GPU plumbing of the hot path: the path shards with no inner-loop collective,
so all that crosses GPUs is one 48-byte G1 point per rank and MSM (commitment / proof partials).

`torch.distributed` is used for the process group only (NCCL on the GPU box, gloo in the CPU tests);
the combine itself is `zkp_g1_sum` (host arithmetic in libzkp_b200.so)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import native


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Point range [lo, hi) of rank `rank` when a row of n points is split across `world` GPUs."""
    if n % world:
        raise ValueError("row length must be divisible by the number of shards")
    per = n // world
    return rank * per, (rank + 1) * per


def gather_bytes(dist, mine: bytes, device: str = "cpu") -> List[bytes]:
    """all_gather of a fixed-size byte string (the N partial points) over the process group."""
    import torch
    t = torch.tensor(list(mine), dtype=torch.uint8, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [bytes(o.cpu().tolist()) for o in out]


def combine_partials(parts: Sequence[bytes]) -> Tuple[bytes, ...]:
    """Each part is k concatenated 48-byte points (e.g. commitment || proof); returns the k sums."""
    if not parts:
        raise ValueError("nothing to combine")
    k = len(parts[0]) // 48
    return tuple(native.g1_sum(b"".join(p[48 * j:48 * (j + 1)] for p in parts)) for j in range(k))
