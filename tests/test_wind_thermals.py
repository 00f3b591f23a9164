"""Rows f-2 / f-3 of SURVEY.md §8f: wind interpolation (`ssrs/simulator.py:765-792`) and thermals
(`ssrs/layers.py:188-214`).  CPU part: the oracle against the committed fixtures (tests/golden/wind_thermals.npz,
written by oracle/make_golden.py with scipy's griddata — the routine the reference calls — and the reference's own
compute_thermals).  GPU part: the CUDA kernels through the C-ABI against the same fixtures."""
import numpy as np
import pytest

from oracle import oracle_np as O

LOGN_MEAN = float(np.exp(5.0 + 0.5 * 0.5 ** 2))          # E[lognormal(2 + 3, 0.5)], layers.py:203-204


def _grid(g):
    rows, cols = (int(v) for v in g["a_shape"])
    res = float(g["a_res"])
    return rows, cols, res, np.linspace(0.0, (cols - 1) * res, cols), np.linspace(0.0, (rows - 1) * res, rows)


def test_oracle_wind_interpolation_matches_fixture(golden):
    g = golden("wind_thermals")
    rows, cols, res, xg, yg = _grid(g)
    for k in "ab":
        ws, wd = O.interpolated_wind_conditions(g[f"{k}_x"], g[f"{k}_y"], g[f"{k}_speed"], g[f"{k}_dirn"], xg, yg)
        assert np.array_equal(np.isnan(ws), np.isnan(g[f"{k}_ws"]))
        assert np.allclose(ws, g[f"{k}_ws"], rtol=1e-12, atol=1e-12, equal_nan=True)
        assert np.allclose(wd, g[f"{k}_wd"], rtol=1e-12, atol=1e-9, equal_nan=True)
    assert not np.isnan(g["a_ws"]).any() and np.isnan(g["b_ws"]).any()       # case b leaves the hull


def test_oracle_thermal_statistics_match_reference(golden):
    g = golden("wind_thermals")
    p = O.thermal_hit_probability(g["t_aspect"].astype(np.float64))
    assert p[:8].sum() == 0 and p[:, :10].sum() == 0 and (p[8:72, 10:90] > 0).all()
    assert 1 / 2999 <= p[p > 0].min() and p.max() <= 1 / 999
    # the Gaussian smoothing (zero padding) keeps almost all of the mass: E[mean(wt)] ~ mean(p) * E[lognormal]
    expect = p.mean() * LOGN_MEAN
    means = g["t_reference_means"]
    sem = means.std(ddof=1) / np.sqrt(len(means))
    assert abs(means.mean() - expect) < 4 * sem + 0.02 * expect
    sm = O.smooth_thermals(g["t_seeds"].astype(np.float64))
    assert np.allclose(sm, g["t_smoothed"], rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_gpu_wind_interpolation(golden, case):
    from ssrs_b200 import layers
    g = golden("wind_thermals")
    rows, cols, res, xg, yg = _grid(g)
    ws, wd = layers.interpolate_wind_to_grid(g[f"{case}_x"], g[f"{case}_y"], g[f"{case}_speed"], g[f"{case}_dirn"],
                                             0.0, 0.0, res, (rows, cols))
    ws, wd = ws.cpu().numpy().astype(np.float64), wd.cpu().numpy().astype(np.float64)
    ref_s, ref_d = g[f"{case}_ws"], g[f"{case}_wd"]
    nan_ref = np.isnan(ref_s)
    # hull membership may differ only for cells within rounding of a hull edge
    assert (np.isnan(ws) != nan_ref).sum() <= 2
    ok = ~nan_ref & ~np.isnan(ws)
    assert np.abs(ws[ok] - ref_s[ok]).max() <= 1e-5 * np.abs(ref_s[ok]).max()          # float32 output vs float64
    dd = np.abs(wd[ok] - ref_d[ok])
    dd = np.minimum(dd, 360.0 - dd)                                                      # directions wrap at north
    assert dd.max() <= 1e-5 * 360.0


@pytest.mark.gpu
def test_gpu_wind_interpolation_feeds_stage1(golden):
    """The interpolated rasters drive the per-cell-wind stencil exactly like the reference's
    compute_orographic_updrafts_using_wtk (simulator.py:200-215)."""
    from ssrs_b200 import layers
    from ssrs_b200.synth import synthetic_dem
    g = golden("wind_thermals")
    rows, cols, res, xg, yg = _grid(g)
    z = synthetic_dem(rows, cols, res, seed=5, rough_rms=10.0)
    ws, wd = layers.interpolate_wind_to_grid(g["a_x"], g["a_y"], g["a_speed"], g["a_dirn"], 0.0, 0.0, res, (rows, cols))
    oro = layers.updraft_fields(z, res, ws, wd, 0.75, want=("orograph",))["orograph"]
    oro = oro if isinstance(oro, np.ndarray) else oro.cpu().numpy()
    _, _, ref_oro, _ = O.updraft_pipeline(z, res, g["a_ws"], g["a_wd"], 0.75)
    assert np.abs(oro - ref_oro).max() <= 1e-5 * max(1.0, np.abs(ref_oro).max())


@pytest.mark.gpu
def test_gpu_gaussian_blur_matches_scipy(golden):
    from ssrs_b200 import layers
    g = golden("wind_thermals")
    out = layers.gaussian_filter_constant(g["t_seeds"], 4.0, 4.0)
    assert np.abs(out - g["t_smoothed"]).max() <= 1e-5 * np.abs(g["t_smoothed"]).max()
    rng = np.random.RandomState(0)
    x = rng.rand(37, 53).astype(np.float32)                      # ragged size, kernel wider than the borders
    from scipy import ndimage
    ref = ndimage.gaussian_filter(x.astype(np.float64), sigma=2.5, mode='constant', truncate=3.0)
    assert np.abs(layers.gaussian_filter_constant(x, 2.5, 3.0) - ref).max() <= 2e-6


@pytest.mark.gpu
def test_gpu_thermals_distribution(golden):
    """Distributional parity (the reference draws from numpy's global stream cell by cell)."""
    from ssrs_b200 import layers
    g = golden("wind_thermals")
    asp = np.tile(g["t_aspect"], (10, 10))                      # 800 x 1000 cells: ~400 seeds per realisation
    p = O.thermal_hit_probability(asp.astype(np.float64))
    seeds = layers.thermal_seeds(asp, 2.0, seed=123)
    assert (seeds[p == 0] == 0).all()                           # 10 % border stays empty
    hits = seeds > 0
    n_exp, n_sd = p.sum(), np.sqrt((p * (1 - p)).sum())
    assert abs(hits.sum() - n_exp) <= 4 * n_sd
    logs = np.log(seeds[hits].astype(np.float64))
    assert abs(logs.mean() - 5.0) <= 4 * 0.5 / np.sqrt(hits.sum()) and abs(logs.std() - 0.5) <= 0.06
    assert not np.array_equal(layers.thermal_seeds(asp, 2.0, seed=124), seeds)
    assert np.array_equal(layers.thermal_seeds(asp, 2.0, seed=123), seeds)       # counter-based: reproducible
    # full compute_thermals on the fixture's grid: mean over realisations against the reference's realisations
    means = np.array([layers.compute_thermals(g["t_aspect"], 2.0, seed=s).mean() for s in range(40)])
    ref = g["t_reference_means"]
    sem = np.sqrt(means.var(ddof=1) / len(means) + ref.var(ddof=1) / len(ref))
    assert abs(means.mean() - ref.mean()) <= 4 * sem


@pytest.mark.gpu
def test_gpu_wind_interpolation_nearest(golden):
    """Config.wtk_interp_type = 'nearest': every cell takes its closest site's wind, exactly the site griddata's k-d tree
    returns (the interpolated value IS a site value, so the comparison is to float32 rounding, on every cell); the
    filter that prunes the sites per CTA is exercised with few sites (fixture, ~25) and with many (3000 random ones,
    dozens of survivors per CTA), and outside the sites' hull where the nearest site is far away."""
    from ssrs_b200 import layers
    g = golden("wind_thermals")
    rows, cols, res, xg, yg = _grid(g)
    rng = np.random.RandomState(3)
    n = 3000
    dense = (rng.uniform(-0.2, 0.6, n) * cols * res, rng.uniform(0.3, 1.2, n) * rows * res, rng.uniform(2.0, 14.0, n),
             rng.uniform(0.0, 360.0, n))
    for xl, yl, spd, drn in ((g["a_x"], g["a_y"], g["a_speed"], g["a_dirn"]), (g["b_x"], g["b_y"], g["b_speed"], g["b_dirn"]), dense):
        ws, wd = layers.interpolate_wind_to_grid(xl, yl, spd, drn, 0.0, 0.0, res, (rows, cols), method='nearest')
        ws, wd = ws.cpu().numpy().astype(np.float64), wd.cpu().numpy().astype(np.float64)
        ref_s, ref_d = O.interpolated_wind_conditions(xl, yl, spd, drn, xg, yg, method='nearest')
        assert not np.isnan(ws).any() and not np.isnan(wd).any()
        # a cell exactly between two sites may go either way in the tree: allow a handful of such cells
        bad = np.abs(ws - ref_s) > 1e-6 * np.abs(ref_s).max()
        dd = np.abs(wd - ref_d); dd = np.minimum(dd, 360.0 - dd)
        bad |= dd > 1e-5 * 360.0
        assert bad.sum() <= 2, bad.sum()
    with pytest.raises(ValueError):          # griddata's own error for a method it does not know
        layers.interpolate_wind_to_grid(g["a_x"], g["a_y"], g["a_speed"], g["a_dirn"], 0.0, 0.0, res, (rows, cols), method='quintic')


def _ct_case(seed, n, narrow=False):
    rng = np.random.RandomState(seed)
    pts = rng.rand(n, 2) * np.array([6000.0, 5000.0])
    if narrow:                              # a sliver on the hull and two nearly coincident sites
        pts[0] = (3000.0, -1e-3); pts[1] = (10.0, 0.0); pts[2] = (5990.0, 0.0); pts[3] = pts[4] + 1e-6
    vals = 8.0 + 3.0 * np.sin(pts[:, 0] / 900.0) * np.cos(pts[:, 1] / 700.0) + 0.3 * rng.randn(n)
    return pts, vals


@pytest.mark.parametrize("seed,n,narrow", [(0, 12, False), (1, 60, False), (2, 400, False), (3, 60, True)])
def test_clough_tocher_arithmetic_matches_scipy(seed, n, narrow):
    """Config.wtk_interp_type = 'cubic' (config.py:60 -> griddata(method='cubic'), simulator.py:772): the lines of
    csrc/clough_tocher.cuh, compiled for the host, against scipy's CloughTocher2DInterpolator — the gradient sweeps
    (same count, same values) and the cubic inside every triangle, on the hull's triangles and in slivers."""
    from scipy.interpolate import CloughTocher2DInterpolator
    from scipy.spatial import Delaunay
    import ctemu
    pts, vals = _ct_case(seed, n, narrow)
    tri = Delaunay(pts)
    ref = CloughTocher2DInterpolator(tri, vals)          # griddata's defaults: tol 1e-6, maxiter 400, no rescaling
    grad, sweeps = ctemu.gradients(pts, vals, *tri.vertex_neighbor_vertices)
    assert sweeps > 0
    scale = np.abs(ref.grad).max()
    assert np.abs(grad - ref.grad[:, 0, :]).max() <= 1e-12 * scale
    rng = np.random.RandomState(seed + 100)
    q = rng.rand(4000, 2) * np.array([6400.0, 5400.0]) - 200.0            # some queries fall outside the hull
    simplex = tri.find_simplex(q)
    got = ctemu.interpolate(pts, vals, grad, tri.simplices, tri.neighbors, q[:, 0], q[:, 1], simplex)
    want = ref(q)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and (simplex < 0).any() and (simplex >= 0).sum() > 1000
    ok = ~np.isnan(want)
    assert np.abs(got[ok] - want[ok]).max() <= 1e-9 * np.abs(want[ok]).max()
    # the interpolant reproduces the site values (two sites 1e-6 m apart: the barycentric coordinates lose digits)
    used = np.unique(tri.simplices)
    s_at = tri.find_simplex(pts[used])
    at_sites = ctemu.interpolate(pts, vals, grad, tri.simplices, tri.neighbors, pts[used, 0], pts[used, 1], s_at)
    assert np.abs(at_sites - vals[used]).max() <= (1e-6 if narrow else 1e-9) * np.abs(vals).max()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_gpu_wind_interpolation_cubic(golden, case):
    """'cubic' through the C-ABI (gradient sweeps, Bezier ordinates and the per-cell cubic all on the device) against
    griddata(method='cubic') on every cell of the fixtures' grids; case b leaves the hull."""
    from ssrs_b200 import layers
    g = golden("wind_thermals")
    rows, cols, res, xg, yg = _grid(g)
    xl, yl, spd, drn = (g[f"{case}_{k}"] for k in ("x", "y", "speed", "dirn"))
    ws, wd = layers.interpolate_wind_to_grid(xl, yl, spd, drn, 0.0, 0.0, res, (rows, cols), method='cubic')
    ws, wd = ws.cpu().numpy().astype(np.float64), wd.cpu().numpy().astype(np.float64)
    ref_s, ref_d = O.interpolated_wind_conditions(xl, yl, spd, drn, xg, yg, method='cubic')
    nan_ref = np.isnan(ref_s)
    assert (np.isnan(ws) != nan_ref).sum() <= 2                   # cells within rounding of a hull edge
    assert nan_ref.any() == (case == "b")
    ok = ~nan_ref & ~np.isnan(ws)
    assert np.abs(ws[ok] - ref_s[ok]).max() <= 1e-5 * np.abs(ref_s[ok]).max()
    dd = np.abs(wd[ok] - ref_d[ok])
    dd = np.minimum(dd, 360.0 - dd)
    weight = ref_s[ok] / np.abs(ref_s[ok]).max()                  # the direction of a vanishing vector is ill-conditioned
    assert (dd * np.minimum(1.0, weight * 1e3)).max() <= 1e-5 * 360.0
    # it is not the linear interpolant in disguise
    lin_s, _ = O.interpolated_wind_conditions(xl, yl, spd, drn, xg, yg)
    assert np.nanmax(np.abs(lin_s - ref_s)) > 1e-3 * np.nanmax(ref_s)


@pytest.mark.gpu
def test_gpu_wind_interpolation_full_size():
    """BASELINE config 4 shape: ~800 sites on a jittered 2 km lattice -> (5000, 6000) rasters; compared with
    griddata on every 40th grid line (griddata over all 3e7 cells takes minutes on the host)."""
    import time
    import torch
    from ssrs_b200 import layers
    from ssrs_b200.synth import synthetic_wind_lattice
    rows, cols, res = 5000, 6000, 10.0
    xl, yl, spd, drn = synthetic_wind_lattice(rows, cols, res, spacing_m=2000.0, seed=7)
    tri = layers.delaunay_triangles(xl, yl)
    layers.interpolate_wind_to_grid(xl, yl, spd, drn, 0.0, 0.0, res, (rows, cols), triangles=tri)       # warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ws, wd = layers.interpolate_wind_to_grid(xl, yl, spd, drn, 0.0, 0.0, res, (rows, cols), triangles=tri)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"interpolate_wind_to_grid {len(xl)} sites -> {rows}x{cols}: {dt * 1e3:.2f} ms")
    assert not torch.isnan(ws).any() and not torch.isnan(wd).any()          # the lattice is padded beyond the region
    xg = np.linspace(0.0, (cols - 1) * res, cols)[::40]
    yg = np.linspace(0.0, (rows - 1) * res, rows)[::40]
    ref_s, ref_d = O.interpolated_wind_conditions(xl, yl, spd, drn, xg, yg)
    got_s = ws[::40, ::40].cpu().numpy().astype(np.float64)
    got_d = wd[::40, ::40].cpu().numpy().astype(np.float64)
    assert np.abs(got_s - ref_s).max() <= 1e-5 * ref_s.max()
    dd = np.abs(got_d - ref_d); dd = np.minimum(dd, 360.0 - dd)
    assert dd.max() <= 1e-5 * 360.0
    # 'nearest' at the same size
    layers.interpolate_wind_to_grid(xl, yl, spd, drn, 0.0, 0.0, res, (rows, cols), method='nearest')
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ws, wd = layers.interpolate_wind_to_grid(xl, yl, spd, drn, 0.0, 0.0, res, (rows, cols), method='nearest')
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"interpolate_wind_to_grid(nearest) {len(xl)} sites -> {rows}x{cols}: {dt * 1e3:.2f} ms")
    ref_s, ref_d = O.interpolated_wind_conditions(xl, yl, spd, drn, xg, yg, method='nearest')
    got_s = ws[::40, ::40].cpu().numpy().astype(np.float64)
    got_d = wd[::40, ::40].cpu().numpy().astype(np.float64)
    dd = np.abs(got_d - ref_d); dd = np.minimum(dd, 360.0 - dd)
    assert ((np.abs(got_s - ref_s) > 1e-6 * ref_s.max()) | (dd > 1e-5 * 360.0)).sum() <= 2
    # 'cubic' at the same size (one triangulation + adjacency for all cases)
    topo = layers.delaunay_topology(xl, yl)
    layers.interpolate_wind_to_grid(xl, yl, spd, drn, 0.0, 0.0, res, (rows, cols), triangles=topo, method='cubic')
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ws, wd = layers.interpolate_wind_to_grid(xl, yl, spd, drn, 0.0, 0.0, res, (rows, cols), triangles=topo, method='cubic')
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"interpolate_wind_to_grid(cubic) {len(xl)} sites -> {rows}x{cols}: {dt * 1e3:.2f} ms")
    assert not torch.isnan(ws).any() and not torch.isnan(wd).any()
    ref_s, ref_d = O.interpolated_wind_conditions(xl, yl, spd, drn, xg, yg, method='cubic')
    got_s = ws[::40, ::40].cpu().numpy().astype(np.float64)
    got_d = wd[::40, ::40].cpu().numpy().astype(np.float64)
    assert np.abs(got_s - ref_s).max() <= 1e-5 * ref_s.max()
    dd = np.abs(got_d - ref_d); dd = np.minimum(dd, 360.0 - dd)
    assert dd.max() <= 1e-5 * 360.0
