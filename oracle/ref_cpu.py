"""TEST INFRASTRUCTURE ONLY — the reference's OWN Python hot path on the host cores, for bench.py's CPU legs.

Runs the unmodified `ssrs/movmodel.py` / `ssrs/layers.py` (from /root/reference in the authoring container, from the
staged copy `oracle/_ref/` on the GPU box; oracle/ref_loader.py) the way `ssrs/simulator.py:360-369` runs them:

    with mp.Pool(num_cores) as pool:
        tracks = pool.map(lambda start_loc: generate_simulated_tracks(...), starting_locs)

with `multiprocess` (pathos's backend: fork + dill-pickled closure capturing the fields), `num_cores = os.cpu_count()`.
A track-step is one iteration of the loop at `movmodel.py:285-317`: total = sum(len(track) - 1).
Nothing under ssrs_b200/ imports this file.
"""
import os
import time

import numpy as np

from .ref_loader import available, load_reference


def pool_track_steps(updraft, potential, start_rows, start_cols, move_dirn=0.0, memory=1, nu=1.0, procs=None, seed=None):
    """(track_steps, seconds, procs) of the reference's pool pattern on the given fields.  `updraft` is widened to
    float64 as the reference's thresholded updraft is (layers.py:171-185 returns float64), `potential` stays float32
    (movmodel.py:128)."""
    import multiprocess as mp            # pathos.multiprocessing's backend (ssrs/simulator.py:11)
    _, M = load_reference()
    procs = int(procs or os.cpu_count() or 1)
    U = np.asarray(updraft, dtype=np.float64)
    P = np.asarray(potential, dtype=np.float32)
    starting_locs = [[int(r), int(c)] for r, c in zip(start_rows, start_cols)]
    procs = min(procs, max(1, len(starting_locs)))                   # num_cores = min(track_count, max_cores), :347
    if seed is not None:
        np.random.seed(int(seed))
    shape = U.shape
    t0 = time.perf_counter()
    with mp.Pool(procs) as pool:                                     # :360
        tracks = pool.map(lambda start_loc: M.generate_simulated_tracks(    # :361-369
            move_dirn, start_loc, shape, memory, nu, U, P), starting_locs)
    dt = time.perf_counter() - t0
    return int(sum(len(t) - 1 for t in tracks)), dt, procs


def config1_full(elevation32, resolution, wspeed, wdirn, threshold, start_rows, start_cols, move_dirn=0.0, procs=None):
    """BASELINE configs[0] in full through the reference: stencil, threshold, assembly, SuperLU solve, pooled stepping,
    presence counts (simulator.py:189-198, 230-243, 259-288, 332-369; movmodel.py:410-419).  Seconds per stage."""
    L, M = load_reference()
    z = np.asarray(elevation32, dtype=np.float64)
    out = {}
    t0 = time.perf_counter()
    slope = L.compute_slope_degrees(z, resolution)
    aspect = L.compute_aspect_degrees(z, resolution)
    oro = L.compute_orographic_updraft(wspeed * np.ones(z.shape), wdirn * np.ones(z.shape), slope, aspect).astype(np.float32)
    out["stencil_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    K = L.get_above_threshold_speed(oro, threshold)
    out["threshold_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    mm = M.MovModel(move_dirn, z.shape)
    bn, be = mm.get_boundary_nodes()
    ri, ci, fa = mm.assemble_sparse_linear_system()
    out["assembly_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    phi = mm.solve_sparse_linear_system(K, bn, be, ri, ci, fa)
    out["solve_s"] = time.perf_counter() - t0
    steps, dt, procs = pool_track_steps(K, phi, start_rows, start_cols, move_dirn, procs=procs)
    out.update(track_steps=steps, stepping_s=dt, procs=procs, track_steps_per_s=steps / dt)
    return out


__all__ = ["available", "pool_track_steps", "config1_full"]
