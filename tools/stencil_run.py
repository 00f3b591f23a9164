"""Stage-1 stencil at 5000 x 6000 in its output variants (for timing and ncu): all four rasters (20 B/cell), {orograph,
updraft} (12 B/cell), {orograph} (8 B/cell, what Simulator asks for), and per-cell wind (28 B/cell)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssrs_b200 import layers
from ssrs_b200.synth import synthetic_dem
rows, cols, res = 5000, 6000, 10.0
z = torch.from_numpy(synthetic_dem(rows, cols, res)).cuda()
ws = torch.full((rows, cols), 10.0, device="cuda"); wd = torch.full((rows, cols), 270.0, device="cuda")
def timeit(label, nbytes, **kw):
    for _ in range(3):
        layers.updraft_fields(z, res, kw.get("ws", 10.0), kw.get("wd", 270.0), 0.75, want=kw["want"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        layers.updraft_fields(z, res, kw.get("ws", 10.0), kw.get("wd", 270.0), 0.75, want=kw["want"])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label:40s} {ms*1e3:8.1f} us  {nbytes*rows*cols/ms/1e6:8.1f} GB/s algorithmic ({nbytes} B/cell)")
timeit("all four rasters, uniform wind", 20, want=("slope", "aspect", "orograph", "updraft"))
timeit("orograph + updraft", 12, want=("orograph", "updraft"))
timeit("orograph only", 8, want=("orograph",))
timeit("all four rasters, per-cell wind", 28, want=("slope", "aspect", "orograph", "updraft"), ws=ws, wd=wd)
