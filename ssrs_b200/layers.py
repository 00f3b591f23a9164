"""Host-side mirror of the reference's `ssrs/layers.py` hot-path functions, backed by the CUDA library.

Same names and argument meaning as the reference (`/root/reference/ssrs/layers.py`):
  compute_slope_degrees(z_mat, res)                      :63-93
  compute_aspect_degrees(z_mat, res)                     :96-128
  compute_orographic_updraft(wspeed, wdirn, slope, aspect, min_updraft_val=0.)   :11-22
  get_above_threshold_speed(in_array, threshold)         :171-185
  compute_thermals(aspect, thermal_intensity_scale)      :188-214   (distributional parity: own RNG stream)
plus `interpolate_wind_to_grid` (the reference's `Simulator._get_interpolated_wind_conditions`,
`ssrs/simulator.py:765-792`), `gaussian_filter_constant` and `updraft_fields`, the fused form the Simulator uses (one kernel, one pass over the DEM).
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
    """Reference `layers.py:11-22`.  Elementwise on already computed slope/aspect rasters (`ssrs_orographic_updraft`);
    kept for API parity — the Simulator uses the fused `updraft_fields`."""
    torch = N.require_cuda()
    lib = N.load()
    s, was_tensor = _to_device(slope, torch)
    a, _ = _to_device(aspect, torch)
    if s.shape != a.shape:
        raise ValueError("slope and aspect rasters differ in shape")
    scalar_s, scalar_d = np.isscalar(wspeed), np.isscalar(wdirn)
    ws = None if scalar_s else _to_device(wspeed, torch)[0]
    wd = None if scalar_d else _to_device(wdirn, torch)[0]
    for t in (ws, wd):
        if t is not None and t.shape != s.shape:
            raise ValueError("wind rasters must match the slope raster")
    out = torch.empty_like(s)
    N.check(lib.ssrs_orographic_updraft(N.ptr(s), N.ptr(a), N.ptr(ws), N.ptr(wd), float(wspeed) if scalar_s else 0.0,
                                        float(wdirn) if scalar_d else 0.0, float(min_updraft_val), N.ptr(out), s.numel(),
                                        N.current_stream()), "ssrs_orographic_updraft")
    return _back(out, was_tensor)


def get_above_threshold_speed(in_array, threshold: float):
    """Reference `layers.py:171-185` (vectorised form)."""
    torch = N.require_cuda()
    lib = N.load()
    x, was_tensor = _to_device(in_array, torch)
    out = torch.empty_like(x)
    N.check(lib.ssrs_threshold(N.ptr(x), N.ptr(out), x.numel(), float(threshold), N.current_stream()), "ssrs_threshold")
    return _back(out, was_tensor)


def gaussian_filter_constant(in_array, sigma: float = 4.0, truncate: float = 4.0):
    """scipy.ndimage.gaussian_filter(in_array, sigma, mode='constant', truncate=truncate) on the GPU (float32)."""
    torch = N.require_cuda()
    lib = N.load()
    x, was_tensor = _to_device(in_array, torch)
    if x.dim() != 2:
        raise ValueError("gaussian_filter_constant expects a 2-D raster")
    rows, cols = x.shape
    out, tmp = torch.empty_like(x), torch.empty_like(x)
    w = torch.empty(2 * int(truncate * sigma + 0.5) + 1, dtype=torch.float32, device="cuda")
    N.check(lib.ssrs_gaussian_blur(N.ptr(x), N.ptr(out), N.ptr(tmp), rows, cols, float(sigma), float(truncate), N.ptr(w),
                                   N.current_stream()), "ssrs_gaussian_blur")
    return _back(out, was_tensor)


def thermal_seeds(aspect, thermal_intensity_scale: float, seed: int):
    """The un-smoothed seed field of `compute_thermals` (reference `layers.py:192-206`), Philox keyed by (seed, cell)."""
    torch = N.require_cuda()
    lib = N.load()
    a, was_tensor = _to_device(aspect, torch)
    rows, cols = a.shape
    out = torch.empty_like(a)
    N.check(lib.ssrs_thermal_seeds(N.ptr(a), rows, cols, float(thermal_intensity_scale), int(seed) & (2 ** 64 - 1),
                                   N.ptr(out), N.current_stream()), "ssrs_thermal_seeds")
    return _back(out, was_tensor)


def compute_thermals(aspect, thermal_intensity_scale: float, seed=None):
    """Reference `layers.py:188-214`: smoothed random thermals.  The reference consumes numpy's global stream cell
    by cell; here `seed` (default: one draw from numpy's global stream, so `np.random.seed` still controls it)
    keys a counter-based generator — same distribution, different realisation."""
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    torch = N.require_cuda()
    a, was_tensor = _to_device(aspect, torch)
    wt = gaussian_filter_constant(thermal_seeds(a, thermal_intensity_scale, seed), 4.0, 4.0)      # :211
    return _back(wt, was_tensor)


def delaunay_triangles(xlocs, ylocs):
    """Qhull Delaunay triangulation of the wind sites — what scipy's griddata(method='linear') builds internally."""
    from scipy.spatial import Delaunay
    pts = np.stack([np.asarray(xlocs, dtype=np.float64), np.asarray(ylocs, dtype=np.float64)], 1)
    return np.ascontiguousarray(Delaunay(pts).simplices, dtype=np.int32)


def delaunay_topology(xlocs, ylocs):
    """The triangulation with the adjacency 'cubic' needs: (triangles [nt, 3], neighbors [nt, 3] — the triangle opposite
    vertex k, -1 on the hull —, vertex_nb_indptr [n + 1], vertex_nb_indices), all int32.  Pass the tuple as `triangles=`
    to reuse it for many wind cases."""
    from scipy.spatial import Delaunay
    pts = np.stack([np.asarray(xlocs, dtype=np.float64), np.asarray(ylocs, dtype=np.float64)], 1)
    d = Delaunay(pts)
    indptr, indices = d.vertex_neighbor_vertices
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return i32(d.simplices), i32(d.neighbors), i32(indptr), i32(indices)


def interpolate_wind_to_grid(xlocs, ylocs, wspeed, wdirn, x0: float, y0: float, res: float, gridsize,
                             triangles=None, method: str = 'linear'):
    """Reference `Simulator._get_interpolated_wind_conditions` (`simulator.py:778-792`): wind speed/direction at
    scattered sites -> CUDA float32 rasters `[rows, cols]` (speed, direction in degrees).  `method` is
    `Config.wtk_interp_type`: 'linear' (the default; NaN outside the sites' convex hull; `triangles` may be passed to
    reuse one triangulation for many wind cases), 'nearest' (closest site, defined everywhere) or 'cubic' (griddata's
    Clough-Tocher scheme: C1 piecewise cubic, NaN outside the hull; `triangles` is then `delaunay_topology`'s tuple)."""
    if method not in ('linear', 'nearest', 'cubic'):
        raise ValueError(f"Unknown interpolation method {method!r} for 2 dimensional data")      # griddata's error
    torch = N.require_cuda()
    lib = N.load()
    x = np.ascontiguousarray(xlocs, dtype=np.float64)
    y = np.ascontiguousarray(ylocs, dtype=np.float64)
    ws = np.asarray(wspeed, dtype=np.float64)
    wd = np.asarray(wdirn, dtype=np.float64)
    if not (x.shape == y.shape == ws.shape == wd.shape) or x.ndim != 1 or x.size < 3:
        raise ValueError("site coordinates and wind values must be 1-D arrays of equal length (>= 3 sites)")
    east = np.multiply(ws, np.sin(wd * np.pi / 180.))               # simulator.py:784-785
    north = np.multiply(ws, np.cos(wd * np.pi / 180.))
    if method == 'nearest':
        rows, cols = int(gridsize[0]), int(gridsize[1])
        dev = lambda a: torch.from_numpy(a).to("cuda")
        dx, dy, de, dn = dev(x), dev(y), dev(east), dev(north)
        out_s = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
        out_d = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
        N.check(lib.ssrs_interp_wind_nearest(N.ptr(dx), N.ptr(dy), N.ptr(de), N.ptr(dn), x.size, float(x0), float(y0),
                                             float(res), rows, cols, N.ptr(out_s), N.ptr(out_d), N.current_stream()),
                "ssrs_interp_wind_nearest")
        return out_s, out_d
    rows, cols = int(gridsize[0]), int(gridsize[1])
    dev = lambda a: torch.from_numpy(a).to("cuda")
    owner = torch.empty((rows, cols), dtype=torch.int32, device="cuda")
    out_s = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
    out_d = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
    if method == 'cubic':
        topo = delaunay_topology(x, y) if triangles is None else triangles
        if not (isinstance(topo, tuple) and len(topo) == 4):
            raise ValueError("method='cubic' takes the (triangles, neighbors, indptr, indices) tuple of delaunay_topology")
        tri, nbr, indptr, indices = (np.ascontiguousarray(a, dtype=np.int32) for a in topo)
        if tri.ndim != 2 or tri.shape[1] != 3 or nbr.shape != tri.shape or indptr.shape != (x.size + 1,):
            raise ValueError("inconsistent triangulation arrays")
        dx, dy, de, dn = dev(x), dev(y), dev(east), dev(north)
        dt, dnb, dip, dix = dev(tri), dev(nbr), dev(indptr), dev(indices)
        nbytes = int(lib.ssrs_interp_wind_cubic_scratch_bytes(x.size, tri.shape[0]))
        scratch = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
        N.check(lib.ssrs_interp_wind_cubic(N.ptr(dx), N.ptr(dy), N.ptr(de), N.ptr(dn), x.size, N.ptr(dt), N.ptr(dnb),
                                           tri.shape[0], N.ptr(dip), N.ptr(dix), float(x0), float(y0), float(res), rows, cols,
                                           N.ptr(owner), N.ptr(scratch), N.ptr(out_s), N.ptr(out_d), N.current_stream()),
                "ssrs_interp_wind_cubic")
        sweeps = scratch[-1:].view(torch.int32).cpu().numpy()
        if (sweeps == 0).any():
            import warnings
            warnings.warn("Gradient estimation did not converge, the results may be inaccurate")      # scipy's wording
        return out_s, out_d
    tri = delaunay_triangles(x, y) if triangles is None else np.ascontiguousarray(triangles, dtype=np.int32)
    dx, dy, de, dn, dt = dev(x), dev(y), dev(east), dev(north), dev(tri)
    N.check(lib.ssrs_interp_wind(N.ptr(dx), N.ptr(dy), N.ptr(de), N.ptr(dn), x.size, N.ptr(dt), tri.shape[0], float(x0),
                                 float(y0), float(res), rows, cols, N.ptr(owner), N.ptr(out_s), N.ptr(out_d),
                                 N.current_stream()), "ssrs_interp_wind")
    return out_s, out_d
