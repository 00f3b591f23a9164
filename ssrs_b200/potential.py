"""Stage 2 host side: the directional potential on the GPU (reference `ssrs/movmodel.py:86-128`).

`solve_potential_device(K, move_dirn)` is what the Simulator uses (device tensor in, device tensor out);
`solve_potential_nodes(K, bnodes, benergy)` backs the reference-signature
`MovModel.solve_sparse_linear_system`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


class SolveStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("restarts", C.c_int32), ("levels", C.c_int32), ("converged", C.c_int32),
                ("rel_residual", C.c_double), ("setup_ms", C.c_double), ("solve_ms", C.c_double),
                ("operator_complexity", C.c_double), ("level_rows", C.c_int64 * 24), ("coarsest_rows", C.c_int64),
                ("workspace_bytes", C.c_int64)]

    def as_dict(self):
        return {"iterations": self.iterations, "restarts": self.restarts, "levels": self.levels,
                "converged": self.converged, "rel_residual": self.rel_residual, "setup_ms": self.setup_ms,
                "solve_ms": self.solve_ms, "operator_complexity": self.operator_complexity,
                "level_rows": [int(v) for v in self.level_rows[:self.levels]], "coarsest_rows": int(self.coarsest_rows),
                "workspace_bytes": int(self.workspace_bytes)}


def _solve(K_dev, bnodes, bvals, rtol, max_iter, strict, comm=None, want_f64=False):
    torch = N.require_cuda()
    lib = N.load()
    rows, cols = K_dev.shape
    bn = np.ascontiguousarray(bnodes, dtype=np.int64)
    bv = np.ascontiguousarray(bvals, dtype=np.float64)
    if bn.shape != bv.shape or bn.ndim != 1:
        raise ValueError("bnodes and benergy must be 1-D arrays of equal length")
    phi = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
    st = SolveStats()
    phi64 = None
    if want_f64:
        if comm is not None:
            raise ValueError("the float64 iterate is available from single-rank solves only")
        phi64 = torch.empty((rows, cols), dtype=torch.float64, device="cuda")
        rc = lib.ssrs_potential_solve_f64(N.ptr(K_dev), rows, cols, bn.ctypes.data_as(C.POINTER(C.c_int64)),
                                          bv.ctypes.data_as(C.POINTER(C.c_double)), bn.size, float(rtol), int(max_iter),
                                          N.ptr(phi), N.ptr(phi64), C.byref(st), N.current_stream())
    elif comm is None:
        rc = lib.ssrs_potential_solve(N.ptr(K_dev), rows, cols, bn.ctypes.data_as(C.POINTER(C.c_int64)),
                                      bv.ctypes.data_as(C.POINTER(C.c_double)), bn.size, float(rtol), int(max_iter),
                                      N.ptr(phi), C.byref(st), N.current_stream())
    else:
        rc = lib.ssrs_potential_solve_sharded(N.ptr(K_dev), rows, cols, bn.ctypes.data_as(C.POINTER(C.c_int64)),
                                              bv.ctypes.data_as(C.POINTER(C.c_double)), bn.size, float(rtol),
                                              int(max_iter), N.ptr(phi), C.byref(st), comm, N.current_stream())
    stats = st.as_dict()
    if rc == -4 and not strict:                       # SSRS_ERR_NOT_CONVERGED: potential was still written
        print(f"ssrs_b200: potential solve stopped at relative residual {stats['rel_residual']:.2e}")
    else:
        N.check(rc, "ssrs_potential_solve")
    if want_f64:
        stats["potential_f64"] = phi64
    return phi, stats


def solve_potential_device(conductivity, move_dirn: float, rtol: float = 0.0, max_iter: int = 0, strict: bool = True,
                           sharded: bool = False, want_f64: bool = False):
    """K: CUDA float32 tensor [rows, cols] (or array-like) -> (phi CUDA float32 tensor, stats dict).
    Dirichlet sets come from `MovModel(move_dirn, shape).get_boundary_nodes()`.
    sharded=True (collective over the torch.distributed world; every rank passes the same K): the solve phase is
    row-sharded over the ranks with NCCL halo exchanges (`ssrs_potential_solve_sharded`); every rank gets the
    full potential.
    want_f64=True (single rank): stats["potential_f64"] holds the float64 iterate the potential was rounded from."""
    torch = N.require_cuda()
    from .movmodel import MovModel
    if isinstance(conductivity, torch.Tensor):
        K = conductivity.to(device="cuda", dtype=torch.float32).contiguous()
    else:
        K = torch.from_numpy(np.ascontiguousarray(conductivity, dtype=np.float32)).to("cuda")
    if K.dim() != 2:
        raise ValueError("conductivity must be a 2-D raster")
    bn, bv = MovModel(move_dirn, tuple(K.shape)).get_boundary_nodes()
    comm = None
    if sharded:
        from . import dist as D
        comm = D.native_comm()
    return _solve(K, bn, bv, rtol, max_iter, strict, comm, want_f64)


def solve_potential_nodes(conductivity, bnodes, benergy, rtol: float = 0.0, max_iter: int = 0, return_stats: bool = False):
    """numpy in / numpy float32 out, explicit Dirichlet nodes (reference signature)."""
    torch = N.require_cuda()
    K = torch.from_numpy(np.ascontiguousarray(conductivity, dtype=np.float32)).to("cuda")
    phi, stats = _solve(K, bnodes, benergy, rtol, max_iter, strict=True)
    out = phi.cpu().numpy()
    return (out, stats) if return_stats else out
