"""Host-side mirror of the reference's `ssrs/layers.py` hot-path functions, backed by the CUDA library.

Same names and argument meaning as the reference (`/root/reference/ssrs/layers.py`):
  compute_slope_degrees(z_mat, res)                      :63-93
  compute_aspect_degrees(z_mat, res)                     :96-128
  compute_orographic_updraft(wspeed, wdirn, slope, aspect, min_updraft_val=0.)   :11-22
  get_above_threshold_speed(in_array, threshold)         :171-185
plus `updraft_fields`, the fused form the Simulator uses (one kernel, one pass over the DEM).
Inputs may be numpy arrays (copied to the GPU and back, results as numpy) or CUDA torch tensors
(results stay on the device).  Arithmetic is float32 on the device (the reference computes in float64
and stores float32, `ssrs/simulator.py:198`); agreement is within 1e-5 of each field's maximum.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def _to_device(a, torch, dtype=None):
    dtype = dtype or torch.float32
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype).contiguous(), True
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to("cuda"), False


def _back(t, was_tensor):
    return t if was_tensor else t.cpu().numpy()


def updraft_fields(z_mat, res, wspeed, wdirn, threshold=0.75, want=("slope", "aspect", "orograph", "updraft")):
    """Fused stage 1.  `wspeed`/`wdirn`: scalars (uniform mode) or [rows, cols] arrays (per-cell wind).
    Returns a dict of the requested rasters (float32)."""
    torch = N.require_cuda()
    lib = N.load()
    z, was_tensor = _to_device(z_mat, torch)
    if z.dim() != 2:
        raise ValueError("elevation raster must be 2-D")
    rows, cols = z.shape
    per_cell = not (np.isscalar(wspeed) and np.isscalar(wdirn))
    ws = wd = None
    if per_cell:
        ws, _ = _to_device(wspeed if not np.isscalar(wspeed) else np.full((rows, cols), wspeed, np.float32), torch)
        wd, _ = _to_device(wdirn if not np.isscalar(wdirn) else np.full((rows, cols), wdirn, np.float32), torch)
        if ws.shape != z.shape or wd.shape != z.shape:
            raise ValueError("wind rasters must match the elevation raster")
    out = {k: torch.empty_like(z) for k in want}
    N.check(lib.ssrs_updraft(N.ptr(z), rows, cols, float(res), N.ptr(ws), N.ptr(wd),
                             0.0 if per_cell else float(wspeed), 0.0 if per_cell else float(wdirn), float(threshold),
                             N.ptr(out.get("slope")), N.ptr(out.get("aspect")), N.ptr(out.get("orograph")),
                             N.ptr(out.get("updraft")), N.current_stream()), "ssrs_updraft")
    return {k: _back(v, was_tensor) for k, v in out.items()}


def compute_slope_degrees(z_mat, res: float):
    """Reference `layers.py:63-93`."""
    return updraft_fields(z_mat, res, 0.0, 0.0, want=("slope",))["slope"]


def compute_aspect_degrees(z_mat, res: float):
    """Reference `layers.py:96-128`."""
    return updraft_fields(z_mat, res, 0.0, 0.0, want=("aspect",))["aspect"]


def compute_orographic_updraft(wspeed, wdirn, slope, aspect, min_updraft_val: float = 0.0):
    """Reference `layers.py:11-22`.  Elementwise on already computed slope/aspect rasters; kept for API
    parity (the Simulator uses the fused `updraft_fields`).  Runs as torch elementwise ops on the GPU."""
    torch = N.require_cuda()
    s, was_tensor = _to_device(slope, torch)
    a, _ = _to_device(aspect, torch)
    ws = float(wspeed) if np.isscalar(wspeed) else _to_device(wspeed, torch)[0]
    wd = float(wdirn) if np.isscalar(wdirn) else _to_device(wdirn, torch)[0]
    diff = torch.clamp(torch.cos((a - wd) * (np.pi / 180.0)), min=0.0)
    out = torch.clamp(ws * (torch.sin(s * (np.pi / 180.0)) * diff), min=float(min_updraft_val))
    return _back(out, was_tensor)


def get_above_threshold_speed(in_array, threshold: float):
    """Reference `layers.py:171-185` (vectorised form)."""
    torch = N.require_cuda()
    lib = N.load()
    x, was_tensor = _to_device(in_array, torch)
    out = torch.empty_like(x)
    N.check(lib.ssrs_threshold(N.ptr(x), N.ptr(out), x.numel(), float(threshold), N.current_stream()), "ssrs_threshold")
    return _back(out, was_tensor)
