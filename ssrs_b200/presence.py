"""Presence post-processing on the GPU (reference `ssrs/movmodel.py:422-439`, `ssrs/simulator.py:520-546`)."""
from __future__ import annotations

import numpy as np

from . import _native as N


def smooth_presence_counts(counts, radius: int):
    """Disk-kernel smoothing of a count raster (CUDA int tensor or numpy) -> float32 (same container kind).
    Equals `convolve2d(counts, disk / disk.sum(), mode='same')`."""
    torch = N.require_cuda()
    was_tensor = isinstance(counts, torch.Tensor)
    c = counts.to("cuda") if was_tensor else torch.from_numpy(np.ascontiguousarray(counts)).to("cuda")
    if c.dim() != 2:
        raise ValueError("counts must be a 2-D raster")
    rows, cols = c.shape
    if c.dtype not in (torch.int32, torch.uint32) or not c.is_contiguous():
        if bool((c < 0).any()) or bool((c > 2 ** 32 - 1).any()):
            raise ValueError("counts must be non-negative 32-bit values")
        c = c.to(torch.int64).to(torch.int32).contiguous()      # bit pattern of the uint32 counts
    prefix = torch.empty((rows, cols + 1), dtype=torch.int64, device="cuda")
    lib = N.load()
    N.check(lib.ssrs_row_prefix_sums(N.ptr(c), rows, cols, N.ptr(prefix), N.current_stream()), "ssrs_row_prefix_sums")
    out = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
    N.check(N.load().ssrs_smooth_presence(N.ptr(prefix), rows, cols, int(radius), N.ptr(out), N.current_stream()),
            "ssrs_smooth_presence")
    return out if was_tensor else out.cpu().numpy()


def compute_smooth_presence_counts(tracks, gridshape, radius: float):
    """Reference signature (`movmodel.py:422-439`): tracks -> counts -> disk smoothing, float32 numpy."""
    from .movmodel import compute_presence_counts
    counts = compute_presence_counts(tracks, gridshape)
    return smooth_presence_counts(counts, int(radius))
