"""Multi-GPU plumbing: one process per GPU via torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards naturally: tracks are block-partitioned by global id, fields are replicated,
and the only exchange is one sum all-reduce of the presence raster per (case, realisation).  Because the RNG
is keyed by (seed, global track id, step) and counts are integers, results are bit-identical for any number
of ranks.  Everything degrades to a no-op in a single process.

The row-sharded potential solve (BASELINE config 5) and `ssrs_presence_allreduce` use the library's own NCCL
communicator (`native_comm()`, `ssrs_comm` in include/ssrs_b200.h); torch.distributed only carries the 128-byte
NCCL id from rank 0 to the others.
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


def agree(flag: bool) -> bool:
    """Rank 0's value of `flag` on every rank (decisions that lead into a collective must not diverge)."""
    d = _dist()
    if not d:
        return bool(flag)
    box = [bool(flag)]
    d.broadcast_object_list(box, src=0)
    return bool(box[0])


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


_native_comm = None


def native_comm():
    """The library's NCCL communicator (`ssrs_comm*` as a ctypes void pointer) spanning the torch.distributed
    world, created once per process; None in a single process.  Collective: every rank must call it."""
    global _native_comm
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return None
    if _native_comm is None:
        import ctypes as C

        from . import _native as N
        torch = N.require_cuda()
        lib = N.load()
        buf = (C.c_uint8 * 128)()
        if d.get_rank() == 0:
            N.check(lib.ssrs_nccl_unique_id(buf), "ssrs_nccl_unique_id")
        box = [bytes(buf)]
        d.broadcast_object_list(box, src=0)
        C.memmove(buf, box[0], 128)
        torch.cuda.synchronize()
        out = C.c_void_p()
        N.check(lib.ssrs_comm_create_nccl(buf, d.get_rank(), d.get_world_size(), C.byref(out)), "ssrs_comm_create_nccl")
        _native_comm = out
    return _native_comm


def halo_mode() -> str:
    """How the native communicator exchanges the sharded solve's halos: 'peer' (stores into the neighbours' memory over
    NVLink, one kernel per exchange; SSRS_COMM_HALO=peer), 'nccl' (grouped ncclSend/ncclRecv, the default) or
    'none' (single rank)."""
    comm = native_comm()
    if comm is None:
        return "none"
    from . import _native as N
    return {1: "peer", 0: "nccl"}.get(int(N.load().ssrs_comm_halo_mode(comm)), "none")


def destroy_native_comm() -> None:
    global _native_comm
    if _native_comm is not None:
        from . import _native as N
        N.load().ssrs_comm_destroy(_native_comm)
        _native_comm = None


def presence_allreduce(presence):
    """In-place sum of a CUDA int32/uint32 presence raster over all ranks through `ssrs_presence_allreduce`."""
    comm = native_comm()
    if comm is None:
        return presence
    from . import _native as N
    if not presence.is_cuda or presence.element_size() != 4 or not presence.is_contiguous():
        raise ValueError("presence must be a contiguous CUDA raster of 32-bit counts")
    N.check(N.load().ssrs_presence_allreduce(N.ptr(presence), presence.numel(), comm, N.current_stream()),
            "ssrs_presence_allreduce")
    return presence
