"""TEST INFRASTRUCTURE ONLY — builds ssrs_b200/csrc/potential.cu with -DSSRS_HOST_EMU (pfor() becomes a
serial host loop) into tests/_build/ so `-m "not gpu"` tests can check the solver's logic on CPU.
The product package never loads this library."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "ssrs_b200", "csrc", "potential.cu")
DEPS = [SRC, os.path.join(ROOT, "ssrs_b200", "csrc", "pfor.cuh"), os.path.join(ROOT, "include", "ssrs_b200.h")]
OUT = os.path.join(ROOT, "tests", "_build", "libssrs_solver_emu.so")


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("restarts", C.c_int32), ("levels", C.c_int32), ("converged", C.c_int32),
                ("rel_residual", C.c_double), ("setup_ms", C.c_double), ("solve_ms", C.c_double),
                ("operator_complexity", C.c_double), ("level_rows", C.c_int64 * 24), ("coarsest_rows", C.c_int64),
                ("workspace_bytes", C.c_int64)]


def build():
    """SSRS_EMU_FLAGS (e.g. "-fsanitize=address -g -O1", run with LD_PRELOAD=libasan.so) builds a separate library."""
    extra = os.environ.get("SSRS_EMU_FLAGS", "").split()
    out = OUT if not extra else OUT.replace(".so", f"_{zlib.crc32(' '.join(extra).encode()):08x}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in DEPS):
        tmp = f"{out}.{os.getpid()}.tmp"                    # concurrent ranks may build at the same time
        subprocess.run(["g++", "-O2", "-std=c++17", "-DSSRS_HOST_EMU", *extra, "-x", "c++", "-shared", "-fPIC", "-o", tmp, SRC],
                       check=True)
        os.replace(tmp, out)
    return out


_lib = None


def solve(K, bnodes, bvals, rtol=0.0, max_iter=300):
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ssrs_emu_potential_solve.restype = C.c_int
        _lib.ssrs_emu_last_error.restype = C.c_char_p
    K = np.ascontiguousarray(K, dtype=np.float32)
    rows, cols = K.shape
    bn = np.ascontiguousarray(bnodes, dtype=np.int64)
    bv = np.ascontiguousarray(bvals, dtype=np.float64)
    phi = np.zeros((rows, cols), dtype=np.float32)
    st = Stats()
    rc = _lib.ssrs_emu_potential_solve(K.ctypes.data_as(C.c_void_p), C.c_int(rows), C.c_int(cols),
                                       bn.ctypes.data_as(C.c_void_p), bv.ctypes.data_as(C.c_void_p), C.c_int64(bn.size),
                                       C.c_double(rtol), C.c_int(max_iter), phi.ctypes.data_as(C.c_void_p), C.byref(st),
                                       None)
    return rc, phi, st, _lib.ssrs_emu_last_error().decode()


# ---- row-sharded solve over gloo (world_size > 1 CPU tests) ---------------------------------------------
class Comm(C.Structure):
    """Mirror of `ssrs_comm` (include/ssrs_b200.h)."""
    EXCHANGE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, *([C.c_int64] * 8), C.c_void_p)
    ALLREDUCE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int32, C.c_void_p)
    ALLGATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p)
    ALLREDUCE_U32 = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
    _fields_ = [("rank", C.c_int32), ("size", C.c_int32), ("ctx", C.c_void_p), ("exchange", EXCHANGE),
                ("allreduce_sum", ALLREDUCE), ("allgather", ALLGATHER), ("allreduce_u32", ALLREDUCE_U32)]


def gloo_comm():
    """`ssrs_comm` whose callbacks move HOST memory with torch.distributed (gloo); counts the calls."""
    import torch
    import torch.distributed as dist
    rank, size = dist.get_rank(), dist.get_world_size()
    calls = {"exchange": 0, "exchange_bytes": 0, "allreduce": 0, "allgather": 0}

    def view(base, off, nbytes):
        return torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(base + off)))

    def exchange(ctx, base, su_off, su_n, ru_off, ru_n, sd_off, sd_n, rd_off, rd_n, stream):
        try:
            ops = []
            if su_n: ops.append(dist.P2POp(dist.isend, view(base, su_off, su_n).clone(), rank - 1))
            if ru_n: ops.append(dist.P2POp(dist.irecv, view(base, ru_off, ru_n), rank - 1))
            if sd_n: ops.append(dist.P2POp(dist.isend, view(base, sd_off, sd_n).clone(), rank + 1))
            if rd_n: ops.append(dist.P2POp(dist.irecv, view(base, rd_off, rd_n), rank + 1))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            calls["exchange"] += 1
            calls["exchange_bytes"] += su_n + sd_n
            return 0
        except Exception as e:      # never raise through the C frame
            print("exchange failed:", e, flush=True)
            return -1

    def allreduce(ctx, vals, count, stream):
        try:
            if not 1 <= count <= 4 * 16:      # the bound include/ssrs_b200.h documents (4 * SSRS_MAX_RANKS)
                raise ValueError(f"all-reduce of {count} scalars")
            t = torch.from_numpy(np.ctypeslib.as_array(vals, shape=(count,)))
            dist.all_reduce(t)
            calls["allreduce"] += 1
            return 0
        except Exception as e:
            print("allreduce failed:", e, flush=True)
            return -1

    def allgather(ctx, base, offs, stream):
        try:
            o = [offs[i] for i in range(size + 1)]
            for r in range(size):       # variable-size pieces: one broadcast per owner
                if o[r + 1] > o[r]:
                    dist.broadcast(view(base, o[r], o[r + 1] - o[r]), src=r)
            calls["allgather"] += 1
            return 0
        except Exception as e:
            print("allgather failed:", e, flush=True)
            return -1

    def allreduce_u32(ctx, base, count, stream):
        try:
            t = torch.from_numpy(np.ctypeslib.as_array((C.c_int32 * count).from_address(base)))
            dist.all_reduce(t)
            return 0
        except Exception as e:
            print("allreduce_u32 failed:", e, flush=True)
            return -1

    cbs = (Comm.EXCHANGE(exchange), Comm.ALLREDUCE(allreduce), Comm.ALLGATHER(allgather), Comm.ALLREDUCE_U32(allreduce_u32))
    comm = Comm(rank, size, None, *cbs)
    comm._keep = cbs            # keep the callback objects alive
    return comm, calls


def solve_sharded(K, bnodes, bvals, comm, rtol=0.0, max_iter=300):
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ssrs_emu_potential_solve.restype = C.c_int
        _lib.ssrs_emu_last_error.restype = C.c_char_p
    _lib.ssrs_emu_potential_solve_sharded.restype = C.c_int
    K = np.ascontiguousarray(K, dtype=np.float32)
    rows, cols = K.shape
    bn = np.ascontiguousarray(bnodes, dtype=np.int64)
    bv = np.ascontiguousarray(bvals, dtype=np.float64)
    phi = np.zeros((rows, cols), dtype=np.float32)
    st = Stats()
    rc = _lib.ssrs_emu_potential_solve_sharded(K.ctypes.data_as(C.c_void_p), C.c_int(rows), C.c_int(cols),
                                               bn.ctypes.data_as(C.c_void_p), bv.ctypes.data_as(C.c_void_p),
                                               C.c_int64(bn.size), C.c_double(rtol), C.c_int(max_iter),
                                               phi.ctypes.data_as(C.c_void_p), C.byref(st), C.byref(comm), None)
    return rc, phi, st, _lib.ssrs_emu_last_error().decode()
