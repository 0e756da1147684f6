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
    world, n = dist.get_world_size(), len(mine)
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(device)
    out = torch.empty(world * n, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, t)
    raw = out.cpu().numpy().tobytes()
    return [raw[r * n:(r + 1) * n] for r in range(world)]


class HostExchange:
    """Gather of a small fixed-size byte string from every rank of ONE node to rank 0 through a POSIX shared-memory
    segment -- no device round trip, no collective launch.  The partial results of this path are HOST-resident (the
    window fold of an MSM runs on the host, see msm_driver.cuh), so the cross-GPU combine of a multi-process job is a
    host-to-host exchange of ~200 bytes per rank; going through NCCL means an H2D copy, a kernel launch and a D2H copy
    for data that never needed to be on the device (measured: 0.30 ms per step against ~20 us here).  Each rank owns one
    64-byte-aligned slot: [u64 sequence | payload].  gather(step, mine) publishes this rank's payload under `step`; on
    rank 0 it returns every rank's payload once all have published that step, elsewhere it returns None at once --
    after waiting until rank 0 has consumed the step before last (two buffers per rank, so a writer is never more than
    one step ahead of the reader).  The process group (NCCL) stays in charge of barriers and reductions."""

    SLOT = 512

    def __init__(self, rank: int, world: int, nbytes: int, tag: str):
        from multiprocessing import shared_memory
        import struct
        if nbytes + 8 > self.SLOT:
            raise ValueError("payload too large for a slot")
        self.rank, self.world, self.n, self.struct = rank, world, nbytes, struct
        size = self.SLOT * (2 * world + 1)
        name = f"zkpb200_{tag}"
        if rank == 0:
            try:
                old = shared_memory.SharedMemory(name=name)
                old.close()
                old.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=size)
            self.shm.buf[:size] = bytes(size)
        else:
            self.shm = None
        self.name, self.size = name, size

    def attach(self) -> None:
        """non-zero ranks: call after a barrier that follows rank 0's constructor"""
        if self.shm is None:
            from multiprocessing import resource_tracker, shared_memory
            self.shm = shared_memory.SharedMemory(name=self.name)
            try:  # the segment belongs to rank 0: keep this process's tracker from unlinking (and warning about) it
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass

    def _slot(self, rank: int, step: int) -> int:
        return self.SLOT * (2 * rank + (step & 1))

    def gather(self, step: int, mine: bytes):
        """step = 1, 2, 3, ... (strictly increasing, the same on every rank)"""
        buf, st = self.shm.buf, self.struct
        ack = self.SLOT * 2 * self.world
        if self.rank != 0:
            while st.unpack_from("<Q", buf, ack)[0] + 2 < step:  # rank 0 still reads the buffer this step reuses
                pass
        o = self._slot(self.rank, step)
        buf[o + 8:o + 8 + self.n] = mine
        st.pack_into("<Q", buf, o, step)
        if self.rank != 0:
            return None
        out = []
        for r in range(self.world):
            o = self._slot(r, step)
            while st.unpack_from("<Q", buf, o)[0] != step:
                pass
            out.append(bytes(buf[o + 8:o + 8 + self.n]))
        st.pack_into("<Q", buf, ack, step)
        return out

    def close(self) -> None:
        if self.shm is not None:
            self.shm.close()
            if self.rank == 0:
                try:
                    self.shm.unlink()
                except FileNotFoundError:
                    pass
            self.shm = None


def combine_partials(parts: Sequence[bytes]) -> Tuple[bytes, ...]:
    """Each part is k concatenated 48-byte points (e.g. commitment || proof); returns the k sums."""
    if not parts:
        raise ValueError("nothing to combine")
    k = len(parts[0]) // 48
    return tuple(native.g1_sum(b"".join(p[48 * j:48 * (j + 1)] for p in parts)) for j in range(k))


def expand_partials(points48: bytes) -> bytes:
    """k concatenated compressed points -> k uncompressed (96-byte) points: done by every rank on its own partials,
    so that the combining rank adds affine points instead of taking 2 N square roots."""
    return b"".join(native.g1_uncompress(points48[i:i + 48]) for i in range(0, len(points48), 48))


def combine_expanded(parts: Sequence[bytes]) -> Tuple[bytes, ...]:
    """Like combine_partials for parts made by expand_partials (96 bytes per point); returns compressed sums."""
    if not parts:
        raise ValueError("nothing to combine")
    k = len(parts[0]) // 96
    return tuple(native.g1_sum_uncompressed(b"".join(p[96 * j:96 * (j + 1)] for p in parts)) for j in range(k))


def sharded_commit_open(dist, ctx, row: int, slice_be: bytes, x_be: bytes, log_n: int, device: str = "cpu") -> Tuple[bytes, bytes, bytes]:
    """Commit + open of ONE polynomial of 2^log_n evaluations split by point range over the ranks of `dist`
    (rank g holds the shard made by srs_generate_shard(..., g, log2(world)) and `slice_be`, its slice of the
    evaluations).  Two small all-gathers (80 and 48 bytes per rank), no collective inside a kernel; every rank
    returns the same (commitment, y, proof).  `ctx` needs worker_commit / shard_eval_partial / shard_open_partial."""
    com_g = ctx.worker_commit(row, slice_be)
    s_g = ctx.shard_eval_partial(row, slice_be, x_be)
    parts = gather_bytes(dist, com_g + s_g, device)
    com = native.g1_sum(b"".join(p[:48] for p in parts))
    y = native.shard_eval_combine(b"".join(p[48:80] for p in parts), log_n, x_be)
    pi_g = ctx.shard_open_partial(row, slice_be, x_be, y)
    proof = native.g1_sum(b"".join(gather_bytes(dist, pi_g, device)))
    return com, y, proof
