"""Wind-site interpolation at BASELINE config 4's shape (924 sites on a jittered 2 km lattice -> 5000 x 6000) by the three
methods, timed with CUDA events around the library calls (device arrays and triangulation prepared once, so the figures
are the kernels', not the uploads'); run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssrs_b200 import _native as N, layers
from ssrs_b200.synth import synthetic_wind_lattice
rows, cols, res = 5000, 6000, 10.0
xl, yl, spd, drn = synthetic_wind_lattice(rows, cols, res, spacing_m=2000.0, seed=7)
lib = N.load()
tri, nbr, indptr, indices = layers.delaunay_topology(xl, yl)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
east, north = spd * np.sin(np.deg2rad(drn)), spd * np.cos(np.deg2rad(drn))
dx, dy, de, dn, dt, dnb, dip, dix = (dev(a) for a in (xl, yl, east, north, tri, nbr, indptr, indices))
owner = torch.empty((rows, cols), dtype=torch.int32, device="cuda")
ws = torch.empty((rows, cols), dtype=torch.float32, device="cuda"); wd = torch.empty_like(ws)
scratch = torch.empty(int(lib.ssrs_interp_wind_cubic_scratch_bytes(len(xl), len(tri))) // 8, dtype=torch.float64, device="cuda")
st = N.current_stream()
calls = {
    "linear": lambda: lib.ssrs_interp_wind(N.ptr(dx), N.ptr(dy), N.ptr(de), N.ptr(dn), len(xl), N.ptr(dt), len(tri), 0.0, 0.0, res,
                                           rows, cols, N.ptr(owner), N.ptr(ws), N.ptr(wd), st),
    "nearest": lambda: lib.ssrs_interp_wind_nearest(N.ptr(dx), N.ptr(dy), N.ptr(de), N.ptr(dn), len(xl), 0.0, 0.0, res, rows, cols,
                                                    N.ptr(ws), N.ptr(wd), st),
    "cubic": lambda: lib.ssrs_interp_wind_cubic(N.ptr(dx), N.ptr(dy), N.ptr(de), N.ptr(dn), len(xl), N.ptr(dt), N.ptr(dnb), len(tri),
                                                N.ptr(dip), N.ptr(dix), 0.0, 0.0, res, rows, cols, N.ptr(owner), N.ptr(scratch),
                                                N.ptr(ws), N.ptr(wd), st),
}
for name, call in calls.items():
    N.check(call(), name)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        N.check(call(), name)
    e1.record(); torch.cuda.synchronize()
    extra = f", sweeps {scratch[-1:].view(torch.int32).tolist()}" if name == "cubic" else ""
    print(f"{name:8s} {len(xl)} sites, {len(tri)} triangles -> {rows}x{cols}: {e0.elapsed_time(e1) / 3:.3f} ms per call{extra}")
