"""Multi-GPU plumbing: one process per GPU via torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards naturally: tracks are block-partitioned by global id, fields are replicated,
and the only exchange is one sum all-reduce of the presence raster per (case, realisation).  Because the RNG
is keyed by (seed, global track id, step) and counts are integers, results are bit-identical for any number
of ranks.  Everything degrades to a no-op in a single process.
"""
from __future__ import annotations


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def rank() -> int:
    d = _dist()
    return d.get_rank() if d else 0


def world_size() -> int:
    d = _dist()
    return d.get_world_size() if d else 1


def barrier() -> None:
    d = _dist()
    if d:
        d.barrier()


def shard_range(n: int, r: int, w: int):
    """Contiguous block [lo, hi) of global track ids owned by rank r of w (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(w))
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def allreduce_sum(t):
    """In-place sum all-reduce of a tensor (presence raster / counters); returns it."""
    d = _dist()
    if d:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def gather_tracks(tracks):
    """Concatenates per-rank lists of trajectories in rank (= global id) order on every rank."""
    d = _dist()
    if not d:
        return tracks
    parts = [None] * d.get_world_size()
    d.all_gather_object(parts, tracks)
    return [t for p in parts for t in p]
