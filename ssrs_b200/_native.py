"""ctypes binding of `libssrs_b200.so` (the C-ABI in include/ssrs_b200.h).

There is no CPU fallback: importing this module without the built library, or calling a compute entry
point without a CUDA device, raises.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSRS_B200_LIB") or os.path.join(_HERE, "libssrs_b200.so")    # override: kernel A/B experiments

EXPORTS = {
    # name: (restype, argtypes)
    "ssrs_abi_version": (C.c_int, []),
    "ssrs_last_error": (C.c_char_p, []),
    "ssrs_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "ssrs_updraft": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                               C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssrs_orographic_updraft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                          C.c_void_p, C.c_int64, C.c_void_p]),
    "ssrs_threshold": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "ssrs_potential_solve": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_int64,
                                       C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssrs_potential_solve_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_int64,
                                           C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssrs_potential_solve_sharded": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                               C.c_int64, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssrs_interp_wind": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                   C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssrs_interp_wind_nearest": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssrs_interp_wind_cubic_scratch_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "ssrs_interp_wind_cubic": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssrs_thermal_seeds": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p]),
    "ssrs_gaussian_blur": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p]),
    "ssrs_release_workspace": (C.c_int, []),
    "ssrs_reserve_workspace": (C.c_int, [C.c_int, C.c_int]),
    "ssrs_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "ssrs_comm_create_nccl": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ssrs_comm_destroy": (C.c_int, [C.c_void_p]),
    "ssrs_comm_halo_mode": (C.c_int, [C.c_void_p]),
    "ssrs_presence_allreduce": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ssrs_step_tracks": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                   C.POINTER(C.c_double), C.c_int, C.c_double, C.c_uint64, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ssrs_walk_table_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "ssrs_walk_workspace_bytes": (C.c_int64, [C.c_int64]),
    "ssrs_transition_table": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]),
    "ssrs_walk_tracks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                   C.POINTER(C.c_double), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_int, C.c_void_p]),
    "ssrs_step_phase_count": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "ssrs_step_tracks_phased": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                          C.POINTER(C.c_double), C.c_int, C.c_double, C.c_uint64, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "ssrs_interleave_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ssrs_pack_trajectories": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ssrs_row_prefix_sums": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ssrs_smooth_presence": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ssrs_presence_counts": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def load():
    """Loads the shared library (building is `python -m ssrs_b200.build`); never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -m ssrs_b200.build` (nvcc, sm_100a). "
                "ssrs_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.ssrs_abi_version() != 1:
            raise NativeError("libssrs_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().ssrs_last_error().decode("utf-8", "replace")
        if status == -1:
            raise ValueError(msg or what)
        raise NativeError(f"{what} failed with status {status}: {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device visible: ssrs_b200 runs on B200 (sm_100a) only and has no CPU fallback")
    return torch


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
