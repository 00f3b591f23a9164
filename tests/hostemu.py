"""TEST INFRASTRUCTURE ONLY — builds ssrs_b200/csrc/potential.cu with -DSSRS_HOST_EMU (pfor() becomes a
serial host loop) into tests/_build/ so `-m "not gpu"` tests can check the solver's logic on CPU.
The product package never loads this library."""
import ctypes as C
import os
import subprocess

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
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        subprocess.run(["g++", "-O2", "-std=c++17", "-DSSRS_HOST_EMU", "-x", "c++", "-shared", "-fPIC", "-o", OUT, SRC],
                       check=True)
    return OUT


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
