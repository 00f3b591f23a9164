"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/ssrs_oracle.c (the C restatement of the stepper).

Imported by tests/, `__graft_entry__.smoke()` and bench.py's CPU-baseline legs only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle
from .oracle_np import directional_probs

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
        _lib.oracle_step_tracks.restype = C.c_int64
        _lib.oracle_step_tracks.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                            C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_void_p, C.c_int64,
                                            C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib.oracle_presence_counts.restype = None
        _lib.oracle_presence_counts.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        _lib.oracle_philox_uniform.restype = C.c_double
        _lib.oracle_philox_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        _lib.oracle_philox_words.restype = None
        _lib.oracle_philox_words.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def step_tracks(U, P, shape, start_rc, move_dirn, memory=1, nu=1.0, seed=0, track_id0=0, uniforms=None,
                traj_cap=0, want_presence=True, nthreads=1, fast=False):
    """Runs the C oracle.  U, P: float32 [rows, cols] or both None ('drw').  start_rc int32 [n,2].
    uniforms: float64 [n, stride] (verification) or None (Philox).  fast=False: the reference's exact operation
    order; fast=True: the CUDA stepper's production arithmetic (same distribution, fewer divisions).  Returns dict(total_steps, traj_len,
    traj [n,cap,2] or None, presence int32 or None)."""
    rows, cols = shape
    start_rc = np.ascontiguousarray(start_rc, dtype=np.int32)
    n = start_rc.shape[0]
    if U is not None:
        U = np.ascontiguousarray(U, dtype=np.float32)
        P = np.ascontiguousarray(P, dtype=np.float32)
        assert U.shape == (rows, cols) and P.shape == (rows, cols)
    dirp = np.ascontiguousarray(directional_probs(move_dirn * np.pi / 180.0), dtype=np.float64)
    ustride = 0
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        ustride = uniforms.shape[1]
    traj = np.zeros((n, traj_cap, 2), dtype=np.int16) if traj_cap > 0 else None
    traj_len = np.zeros(n, dtype=np.int32)
    presence = np.zeros((rows, cols), dtype=np.int32) if want_presence else None
    total = lib().oracle_step_tracks(_ptr(U), _ptr(P), rows, cols, _ptr(start_rc), n, track_id0, _ptr(dirp),
                                     int(memory), float(nu), int(seed), _ptr(uniforms), ustride, _ptr(traj),
                                     traj_cap, _ptr(traj_len), _ptr(presence), int(nthreads), _mode(fast, U, memory, nu, uniforms, traj_cap))
    return dict(total_steps=int(total), traj_len=traj_len, traj=traj, presence=presence)


def _mode(fast, U, memory, nu, uniforms, traj_cap):
    """0: the reference's exact operation order; 1 (fast=True): production arithmetic of tracks.cu; 2 (fast='table'): the
    transition-table walk of walk.cu (needs fields, memory 1, nu 1, Philox streams, no trajectories)."""
    if fast == "table":
        if U is None or memory != 1 or nu != 1.0 or uniforms is not None or traj_cap:
            raise ValueError("table mode needs fields, memory 1, nu 1, Philox streams and no trajectory output")
        return 2
    return int(bool(fast))


def presence_counts(traj, traj_len, shape):
    rows, cols = shape
    traj = np.ascontiguousarray(traj, dtype=np.int16)
    traj_len = np.ascontiguousarray(traj_len, dtype=np.int32)
    out = np.zeros((rows, cols), dtype=np.int32)
    lib().oracle_presence_counts(_ptr(traj), traj.shape[1], _ptr(traj_len), traj.shape[0], rows, cols, _ptr(out))
    return out


def philox_words(seed, track, block):
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox_words(int(seed), int(track), int(block), _ptr(out))
    return out


def philox_uniform(seed, track, step):
    return lib().oracle_philox_uniform(int(seed), int(track), int(step))
