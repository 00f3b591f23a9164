"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):  python -m oracle.make_golden
Every array below is an output of `/root/reference/ssrs/{layers,movmodel}.py` loaded by
`oracle/ref_loader.py`; the glue between stages restates `ssrs/simulator.py:189-198,230-243,259-288`
(float32 save/reload of the orograph, threshold, MovModel calls).  Per-track random numbers are pinned
as SURVEY.md §4 describes: `np.random.seed(s)` before each serial reference call, and the same stream
re-drawn with `np.random.RandomState(s).random_sample(n)` and stored next to the trajectory.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402
from ssrs_b200.synth import synthetic_dem, synthetic_wind_lattice  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def reference_fields(L, M, z32, res, wspeed, wdirn, thr, dirn):
    z = z32.astype(np.float64)
    slope = L.compute_slope_degrees(z, res)
    aspect = L.compute_aspect_degrees(z, res)
    ws = wspeed * np.ones(z.shape) if np.isscalar(wspeed) else wspeed
    wd = wdirn * np.ones(z.shape) if np.isscalar(wdirn) else wdirn
    oro = L.compute_orographic_updraft(ws, wd, slope, aspect).astype(np.float32)   # simulator.py:196-198
    K = L.get_above_threshold_speed(oro, thr)                                      # simulator.py:240-242
    phi = None
    if dirn is not None:
        mm = M.MovModel(dirn, z.shape)
        bn, be = mm.get_boundary_nodes()
        ri, ci, fa = mm.assemble_sparse_linear_system()
        phi = mm.solve_sparse_linear_system(K, bn, be, ri, ci, fa)                 # simulator.py:276-283
    return slope, aspect, oro, K, phi


def main():
    L, M = load_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---- tables -------------------------------------------------------------------------------
    masks = np.stack([M.get_track_restrictions(i // 3 - 1, i % 3 - 1) for i in range(9)])
    thetas = np.array([0, 30, 45, 90, 135, 180, 225, 270, 315, -45, 17.3])
    dirw = np.stack([M.get_directional_probs(t * np.pi / 180.0) for t in thetas])
    thr_x = np.concatenate([np.linspace(0, 2.0, 81), [0.01, 0.0100001, 0.75, 0.7500001]]).astype(np.float32)
    thr_y = L.get_above_threshold_speed(thr_x, 0.75).astype(np.float64)
    tab = dict(masks=masks, thetas=thetas, dirw=dirw, thr_x=thr_x, thr_y=thr_y,
               norms_inv=M.neighbour_delta_norms_inv, deltas=np.array(M.neighbour_deltas))
    for th in (0, 30, 45, 90, 180, 270, 315, -45):
        for shp in ((23, 31), (60, 50)):
            bn, be = M.MovModel(th, shp).get_boundary_nodes()
            tab[f"bn_{th}_{shp[0]}x{shp[1]}"] = bn
            tab[f"be_{th}_{shp[0]}x{shp[1]}"] = be
    np.random.seed(4)
    r, c = M.get_starting_indices(64, (5, 55, 1, 2), 'random', (60., 50.), 100.)
    tab["start_random_rows"], tab["start_random_cols"] = r, c
    r, c = M.get_starting_indices(37, (5, 55, 1, 2), 'structured', (60., 50.), 100.)
    tab["start_struct_rows"], tab["start_struct_cols"] = r, c
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **tab)

    # ---- stage 1: uniform and per-cell wind on a 60x50 and a ragged 37x45 grid ------------------
    st = {}
    for name, (rows, cols, res) in {"a": (50, 60, 100.0), "b": (37, 45, 30.0)}.items():
        z = synthetic_dem(rows, cols, res, seed=11, rough_rms=8.0)
        z[5:9, 7:12] = z[5, 7]                        # a flat patch: dz_dx == 0 -> aspect 270 (layers.py:124)
        sl, asp, oro, K, _ = reference_fields(L, M, z, res, 10.0, 270.0, 0.75, None)
        st.update({f"{name}_z": z, f"{name}_res": res, f"{name}_slope": sl, f"{name}_aspect": asp,
                   f"{name}_oro": oro, f"{name}_K": K})
        rng = np.random.RandomState(5)
        ws = (8.0 + 3.0 * rng.rand(rows, cols)).astype(np.float32)
        wd = (270.0 + 60.0 * (rng.rand(rows, cols) - 0.5)).astype(np.float32)
        _, _, oro2, K2, _ = reference_fields(L, M, z, res, ws.astype(np.float64), wd.astype(np.float64), 0.75, None)
        st.update({f"{name}_ws": ws, f"{name}_wd": wd, f"{name}_oro_cell": oro2, f"{name}_K_cell": K2})
    np.savez_compressed(os.path.join(OUT, "stencil.npz"), **st)

    # ---- stage 2: potentials -------------------------------------------------------------------
    pot = {}
    rng = np.random.RandomState(2)
    Krand = rng.rand(23, 31) * (rng.rand(23, 31) > 0.4)            # random K with zeros
    Krand[3:6, 4:9] *= 1e-9                                        # tiny positive K: conducts less than 0
    for th in (0, 90, 180, 270, 45, -45, 30):
        mm = M.MovModel(th, Krand.shape)
        bn, be = mm.get_boundary_nodes()
        ri, ci, fa = mm.assemble_sparse_linear_system()
        pot[f"rand_phi_{th}"] = mm.solve_sparse_linear_system(Krand, bn, be, ri, ci, fa)
    pot["rand_K"] = Krand
    z = synthetic_dem(50, 60, 100.0, seed=11, rough_rms=8.0)
    for th in (0, 270, 45):
        _, _, oro, K, phi = reference_fields(L, M, z, 100.0, 10.0, 270.0, 0.75, th)
        pot[f"dem_phi_{th}"] = phi
    pot["dem_K"] = K
    z2 = synthetic_dem(120, 150, 100.0, seed=3, rough_rms=15.0)
    _, _, _, K2, phi2 = reference_fields(L, M, z2, 100.0, 10.0, 270.0, 0.75, 0.0)
    pot["dem2_K"] = K2.astype(np.float32)          # stored as float32 (what the GPU path consumes)
    _, _ = None, None
    mm = M.MovModel(0.0, K2.shape)
    bn, be = mm.get_boundary_nodes()
    ri, ci, fa = mm.assemble_sparse_linear_system()
    pot["dem2_phi_0"] = mm.solve_sparse_linear_system(K2.astype(np.float32).astype(np.float64), bn, be, ri, ci, fa)
    np.savez_compressed(os.path.join(OUT, "potential.npz"), **pot)

    # ---- stage 3+4: trajectories on the 50x60 fields --------------------------------------------
    _, _, oro, K, phi = reference_fields(L, M, z, 100.0, 10.0, 270.0, 0.75, 0.0)
    U32 = K.astype(np.float32)
    U = U32.astype(np.float64)                     # the stepper sees exactly these values on both sides
    tr = dict(U32=U32, P32=phi)
    cases = {"n0_m1": (0.0, 1, 1.0), "n0_m3": (0.0, 3, 1.0), "n0_m0": (0.0, 0, 1.0), "d45_m1": (45.0, 1, 1.0),
             "d270_m2": (270.0, 2, 1.0), "n0_nu05": (0.0, 1, 0.5), "n0_nu0": (0.0, 1, 0.0)}
    rng = np.random.RandomState(9)
    for name, (dirn, mem, nu) in cases.items():
        n = 8
        starts = np.stack([rng.randint(1, 48, n), rng.randint(1, 58, n)], 1).astype(np.int32)
        starts[0] = (0, 0)                         # exercises the burn-in relocation from a corner
        starts[1] = (49, 59)
        trajs, lens = [], []
        for t in range(n):
            np.random.seed(1000 + t)
            tj = M.generate_simulated_tracks(dirn, [int(starts[t, 0]), int(starts[t, 1])], U.shape, mem, nu, U, phi)
            trajs.append(tj)
            lens.append(len(tj))
        cap = max(lens)
        cap = min(cap, 4000)
        uni = np.stack([np.random.RandomState(1000 + t).random_sample(cap) for t in range(n)])
        packed = np.zeros((n, cap, 2), dtype=np.int16)
        for t in range(n):
            m = min(lens[t], cap)
            packed[t, :m] = trajs[t][:m]
        tr.update({f"{name}_params": np.array([dirn, mem, nu]), f"{name}_starts": starts, f"{name}_len": np.array(lens),
                   f"{name}_traj": packed, f"{name}_uni": uni,
                   f"{name}_presence": M.compute_presence_counts(trajs, U.shape)})
    # 'drw' model: no fields
    np.random.seed(77)
    tj = M.generate_simulated_tracks(30.0, [5, 30], U.shape, 1, 1.0)
    tr["drw_traj"] = tj
    tr["drw_uni"] = np.random.RandomState(77).random_sample(len(tj) + 1)
    np.savez_compressed(os.path.join(OUT, "tracks.npz"), **tr)

    # ---- smoothing ("next" row f-1) ---------------------------------------------------------------
    cnt = M.compute_presence_counts(trajs, U.shape)
    sm = {"counts": cnt}
    for rad in (2, 5):
        sm[f"smooth_{rad}"] = M.compute_smooth_presence_counts(trajs, U.shape, rad)
    np.savez_compressed(os.path.join(OUT, "smooth.npz"), **sm)
    # ---- "next" rows f-2 / f-3: wind interpolation and thermals --------------------------------------
    from oracle import oracle_np as O
    wn = {}
    rows, cols, res = 70, 90, 250.0
    xl, yl, spd, drn = synthetic_wind_lattice(rows, cols, res, spacing_m=2000.0, seed=7)
    xg = np.linspace(0.0, (cols - 1) * res, cols)                   # get_terrain_grid, simulator.py:177-185
    yg = np.linspace(0.0, (rows - 1) * res, rows)
    ws, wd = O.interpolated_wind_conditions(xl, yl, spd, drn, xg, yg)            # scipy griddata, as the reference calls it
    wn.update(dict(a_x=xl, a_y=yl, a_speed=spd, a_dirn=drn, a_shape=np.array([rows, cols]), a_res=res, a_ws=ws, a_wd=wd))
    # sites that do not cover the grid: NaN outside the hull; wrap-around directions near north
    rng = np.random.RandomState(12)
    xb = rng.uniform(2000, 18000, 40); yb = rng.uniform(1000, 15000, 40)
    sb = rng.uniform(3, 14, 40); db = np.mod(rng.normal(0.0, 40.0, 40), 360.0)
    ws2, wd2 = O.interpolated_wind_conditions(xb, yb, sb, db, xg, yg)
    wn.update(dict(b_x=xb, b_y=yb, b_speed=sb, b_dirn=db, b_ws=ws2, b_wd=wd2))
    # thermals: the reference function itself (layers.py:188-214) for distribution statistics, and its smoothing
    z = synthetic_dem(80, 100, 100.0, seed=11, rough_rms=8.0)
    asp = L.compute_aspect_degrees(z.astype(np.float64), 100.0)
    np.random.seed(3)
    reals = np.stack([L.compute_thermals(asp, 2.0) for _ in range(40)])
    rng = np.random.RandomState(8)
    seeds = np.where(rng.rand(80, 100) < 0.01, rng.lognormal(5.0, 0.5, (80, 100)), 0.0)
    wn.update(dict(t_aspect=asp.astype(np.float32), t_reference_means=reals.mean(axis=(1, 2)), t_reference_max=reals.max(axis=(1, 2)),
                   t_seeds=seeds.astype(np.float32), t_smoothed=O.smooth_thermals(seeds.astype(np.float32).astype(np.float64))))
    np.savez_compressed(os.path.join(OUT, "wind_thermals.npz"), **wn)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
