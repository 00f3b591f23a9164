"""Drop-in boundary on the GPU: Simulator(Config, elevation=...) end to end on BASELINE config 1
(uniform mode, 600x500 synthetic DEM at 100 m, wind 10 m/s from 270 deg, 1000 northbound tracks) against the
oracle pipeline, plus the presence smoothing ("next" row f-1) against the reference's golden output."""
import os
import pickle

import numpy as np
import pytest

from oracle import oracle_c as OC
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


def test_config1_end_to_end(tmp_path):
    from ssrs_b200 import Config, Simulator
    from ssrs_b200.synth import synthetic_dem
    cfg = Config(run_name="cfg1", out_dir=str(tmp_path), sim_seed=7, region_width_km=(60., 50.), resolution=100.,
                 sim_mode="uniform", uniform_windspeed=10., uniform_winddirn=270., track_direction=0.,
                 track_count=1000, track_start_region=(5, 55, 1, 2))
    z = synthetic_dem(500, 600, 100.0)
    sim = Simulator(cfg, elevation=z)
    assert sim.gridsize == (500, 600) and sim.case_ids == ["s10d270"]
    data = sim.mode_data_dir
    # stage 1 artefact: <case>_orograph.npy float32, within 1e-5 of the reference arithmetic
    oro = np.load(os.path.join(data, "s10d270_orograph.npy"))
    sl, asp, oro_ref, K_ref = O.updraft_pipeline(z, 100.0, 10.0, 270.0, 0.75)
    assert oro.dtype == np.float32 and np.abs(oro - oro_ref).max() <= 1e-5 * oro_ref.max()
    up = sim.load_updrafts("s10d270")[0]
    assert np.abs(up - K_ref).max() <= 1e-5 * K_ref.max()
    sim.simulate_tracks()
    # stage 2 artefact
    pot = np.load(os.path.join(data, "s10d270_d0_t75_fluidflow_r0_potential.npy"))
    pot_ref = O.solve_potential(up.astype(np.float64), 0.0)
    assert pot.dtype == np.float32 and np.abs(pot.astype(np.float64) - pot_ref).max() <= 1e-5 * 1000.0
    # stage 3 artefact: pickled list of int16 [L, 2]
    with open(os.path.join(data, "s10d270_d0_t75_fluidflow_r0_tracks.pkl"), "rb") as f:
        tracks = pickle.load(f)
    assert len(tracks) == 1000 and tracks[0].dtype == np.int16 and tracks[0].shape[1] == 2
    lens = np.array([len(t) for t in tracks])
    assert sim.total_track_steps == int((lens - 1).sum())
    # every track starts in the start region and ends on the border (or ran out of moves)
    assert all(9 <= t[0, 0] <= 19 and 49 <= t[0, 1] <= 549 for t in tracks)
    assert all(t[-1, 0] in (0, 499) or t[-1, 1] in (0, 599) for t in tracks)
    # stage 4: fused counts == counts recomputed from the stored tracks by the oracle, bit-exact
    assert np.array_equal(sim.presence_counts(), O.presence_counts(tracks, (500, 600)).astype(np.int32))
    # the same tracks are reproduced by the C oracle from the product's own fields and seed (Philox streams)
    np.random.seed(7)
    from ssrs_b200.movmodel import get_starting_indices
    sr, sc = get_starting_indices(1000, (5, 55, 1, 2), "random", (60., 50.), 100.)
    ref = OC.step_tracks(up, pot, (500, 600), np.stack([sr, sc], 1), 0.0, 1, 1.0, seed=sim._track_seed(0, 0),
                         traj_cap=int(lens.max()), nthreads=4, fast=True)
    assert np.array_equal(ref["traj_len"], lens)
    assert np.array_equal(ref["presence"], sim.presence_counts())
    # "next" row: summary presence written with the reference's name and dtype
    summ = sim.plot_presence_map(radius=1000.)
    saved = np.load(os.path.join(data, "summary_presence.npy"))
    assert saved.dtype == np.float32 and saved.max() == 1.0 and np.array_equal(saved, summ)
    sm_ref = O.smooth_presence(sim.presence_counts(), 10)
    assert np.allclose(saved, sm_ref / sm_ref.max(), rtol=1e-5, atol=1e-7)
    # second construction finds the cached potential (reference cache rule)
    sim2 = Simulator(cfg, elevation=z)
    sim2.simulate_tracks()
    assert np.array_equal(sim2.presence_counts(), sim.presence_counts())


def test_snapshot_mode_per_cell_wind(tmp_path):
    from ssrs_b200 import Config, Simulator
    from ssrs_b200.synth import synthetic_dem
    rows, cols = 120, 160
    z = synthetic_dem(rows, cols, 100.0, seed=4)
    rng = np.random.RandomState(0)
    ws = (8 + 2 * rng.rand(rows, cols)).astype(np.float32)
    wd = (270 + 40 * (rng.rand(rows, cols) - 0.5)).astype(np.float32)
    cfg = Config(run_name="snap", out_dir=str(tmp_path), sim_seed=3, sim_mode="snapshot", region_width_km=(16., 12.),
                 resolution=100., track_count=64, track_start_region=(2, 14, 0.5, 1), track_direction=0.)
    sim = Simulator(cfg, elevation=z, wind_cases={"y2014m12d01h15": (ws, wd)})
    oro = np.load(os.path.join(sim.mode_data_dir, "y2014m12d01h15_orograph.npy"))
    _, _, oro_ref, _ = O.updraft_pipeline(z, 100.0, ws, wd, 0.75)
    assert np.abs(oro - oro_ref).max() <= 1e-5 * oro_ref.max()
    sim.simulate_tracks()
    assert os.path.exists(os.path.join(sim.mode_data_dir, "y2014m12d01h15_d0_t75_fluidflow_r0_tracks.pkl"))
    with pytest.raises(ValueError):
        Simulator(cfg)                      # no terrain injected
    with pytest.raises(NotImplementedError):
        sim.plot_updrafts()


def test_smoothing_golden(golden):
    from ssrs_b200.presence import compute_smooth_presence_counts, smooth_presence_counts
    g, s = golden("tracks"), golden("smooth")
    for rad in (2, 5):
        out = smooth_presence_counts(s["counts"].astype(np.int32), rad)
        assert out.dtype == np.float32 and np.allclose(out, s[f"smooth_{rad}"], rtol=1e-6, atol=1e-7)
    lens = g["n0_nu0_len"]
    tracks = [g["n0_nu0_traj"][t, :lens[t]] for t in range(len(lens))]
    assert np.allclose(compute_smooth_presence_counts(tracks, g["U32"].shape, 5), s["smooth_5"], rtol=1e-6, atol=1e-7)


def test_smoothing_full_radius():
    """201x201 disk at 10 m (the default 1 km radius the reference cannot evaluate): checked on a window against
    the oracle's convolve2d, and by mass conservation away from the border."""
    import torch
    from ssrs_b200.presence import smooth_presence_counts
    rng = np.random.RandomState(1)
    cnt = (rng.rand(700, 900) < 0.02).astype(np.int32) * rng.randint(1, 50, (700, 900)).astype(np.int32)
    out = smooth_presence_counts(torch.from_numpy(cnt).cuda(), 100).cpu().numpy()
    ref = O.smooth_presence(cnt[150:550, 200:700], 100)
    assert np.allclose(out[250:450, 300:600], ref[100:300, 100:400], rtol=1e-5, atol=1e-7)
    inner = np.zeros_like(cnt); inner[300:400, 400:500] = cnt[300:400, 400:500]
    o2 = smooth_presence_counts(inner, 100)
    assert abs(o2.sum() - inner.sum()) <= 1e-3 * inner.sum()


def test_snapshot_mode_wind_sites_and_thermals(tmp_path):
    """BASELINE config 4 in miniature: wind given at scattered sites (a jittered 2 km lattice), interpolated to the
    terrain grid on the GPU (reference simulator.py:765-792), plus one thermal realisation (layers.py:188-214,
    simulator.py:217-243): realisation ids r0 (orographic only) and r1 (orographic + thermals)."""
    from ssrs_b200 import Config, Simulator
    from ssrs_b200.synth import synthetic_dem, synthetic_wind_lattice
    rows, cols, res = 120, 160, 100.0
    z = synthetic_dem(rows, cols, res, seed=4)
    xl, yl, spd, drn = synthetic_wind_lattice(rows, cols, res, spacing_m=2000.0, seed=7)
    cfg = Config(run_name="snap2", out_dir=str(tmp_path), sim_seed=3, sim_mode="snapshot", region_width_km=(16., 12.),
                 resolution=res, track_count=64, track_start_region=(2, 14, 0.5, 1), track_direction=0.,
                 thermals_realization_count=1)
    sim = Simulator(cfg, elevation=z, wind_points=(xl, yl), wind_cases={"y2014m12d01h15": (spd, drn)})
    data = sim.mode_data_dir
    xg, yg = sim.get_terrain_grid()
    ws_ref, wd_ref = O.interpolated_wind_conditions(xl, yl, spd, drn, xg, yg)
    _, _, oro_ref, _ = O.updraft_pipeline(z, res, ws_ref, wd_ref, 0.75)
    oro = np.load(os.path.join(data, "y2014m12d01h15_orograph.npy"))
    assert np.abs(oro - oro_ref).max() <= 1e-5 * oro_ref.max()
    th = np.load(os.path.join(data, "y2014m12d01h15_r0_thermals.npy"))
    assert th.dtype == np.float32 and th.shape == (rows, cols) and th.min() >= 0 and th.max() > 0
    ups = sim.load_updrafts("y2014m12d01h15")
    assert len(ups) == 2 and np.abs(ups[1] - O.get_above_threshold_speed(oro + th, 0.75)).max() <= 1e-5 * ups[1].max()
    sim.simulate_tracks()
    for r in (0, 1):
        assert os.path.exists(os.path.join(data, f"y2014m12d01h15_d0_t75_fluidflow_r{r}_tracks.pkl"))
        assert os.path.exists(os.path.join(data, f"y2014m12d01h15_d0_t75_fluidflow_r{r}_potential.npy"))
    assert not np.array_equal(sim.presence_counts(real_id=0), sim.presence_counts(real_id=1))
    summ = sim.plot_presence_map(radius=500.)
    assert summ.max() == 1.0
    # same seed -> the same thermal realisation (counter-based RNG keyed by sim_seed, case, realisation)
    sim2 = Simulator(cfg.__class__(**{**cfg.__dict__, "run_name": "snap3"}), elevation=z, wind_points=(xl, yl),
                     wind_cases={"y2014m12d01h15": (spd, drn)})
    assert np.array_equal(np.load(os.path.join(sim2.mode_data_dir, "y2014m12d01h15_r0_thermals.npy")), th)
    # Config.wtk_interp_type = 'cubic' (config.py:60): the orograph follows griddata(method='cubic'); an unknown method
    # fails in the constructor with griddata's error
    over = lambda **kw: cfg.__class__(**{**cfg.__dict__, "thermals_realization_count": 0, **kw})
    sim3 = Simulator(over(run_name="snap4", wtk_interp_type="cubic"), elevation=z, wind_points=(xl, yl),
                     wind_cases={"y2014m12d01h15": (spd, drn)})
    ws_c, wd_c = O.interpolated_wind_conditions(xl, yl, spd, drn, xg, yg, method='cubic')
    _, _, oro_c, _ = O.updraft_pipeline(z, res, ws_c, wd_c, 0.75)
    got_c = np.load(os.path.join(sim3.mode_data_dir, "y2014m12d01h15_orograph.npy"))
    assert np.abs(got_c - oro_c).max() <= 1e-5 * oro_c.max()
    assert np.abs(oro_c - oro_ref).max() > 1e-4 * oro_ref.max()          # and differs from the linear one
    with pytest.raises(ValueError):
        Simulator(over(run_name="snap5", wtk_interp_type="quintic"), elevation=z, wind_points=(xl, yl),
                  wind_cases={"y2014m12d01h15": (spd, drn)})


def test_seasonal_mode_three_cases(tmp_path):
    """BASELINE config 5 in miniature: several wind cases (the reference's seasonal loop over case_ids,
    simulator.py:200-215, 348-386, 520-546): one orograph, potential, track set and presence map per case; the
    summary map is the normalised sum of the per-case maps."""
    from ssrs_b200 import Config, Simulator
    from ssrs_b200.synth import seasonal_wind_conditions, synthetic_dem
    rows, cols, res = 120, 160, 100.0
    z = synthetic_dem(rows, cols, res, seed=4)
    spd, drn = seasonal_wind_conditions(3, seed=11)
    cases = {f"y2015m0{i + 3}d10h12": (float(spd[i]), float(drn[i])) for i in range(3)}
    cfg = Config(run_name="seas", out_dir=str(tmp_path), sim_seed=5, sim_mode="seasonal", region_width_km=(16., 12.),
                 resolution=res, track_count=200, track_start_region=(2, 14, 0.5, 1), track_direction=0.)
    sim = Simulator(cfg, elevation=z, wind_cases=cases)
    assert sim.case_ids == list(cases)
    sim.simulate_tracks()
    maps = []
    for i, cid in enumerate(cases):
        oro = np.load(os.path.join(sim.mode_data_dir, f"{cid}_orograph.npy"))
        _, _, oro_ref, _ = O.updraft_pipeline(z, res, float(spd[i]), float(drn[i]), 0.75)
        assert np.abs(oro - oro_ref).max() <= 1e-5 * max(oro_ref.max(), 1e-30)
        assert os.path.exists(os.path.join(sim.mode_data_dir, f"{cid}_d0_t75_fluidflow_r0_potential.npy"))
        cnt = sim.presence_counts(cid)
        assert cnt.sum() == sim._presence[sim._get_id_string(cid, 0)].sum().item() and cnt.sum() > 200
        sm = O.smooth_presence(cnt, 5)
        maps.append(sm / sm.max())
    summ = sim.plot_presence_map(radius=500.)
    ref = np.sum(maps, axis=0)
    assert np.allclose(summ, ref / ref.max(), rtol=1e-5, atol=1e-6)


def test_potential_cache_is_keyed_by_content(tmp_path, capsys):
    """The reference reuses `<id>_potential.npy` by id string alone (simulator.py:264-272) and would step tracks on a
    potential computed for other terrain.  Here the cache entry carries a content key: same inputs -> reused,
    other terrain under the same run name -> recomputed, a raster without a key -> stale, force_potential -> recomputed."""
    from ssrs_b200 import Config, Simulator
    from ssrs_b200.synth import synthetic_dem
    rows, cols, res = 120, 160, 100.0
    cfg = Config(run_name="cache", out_dir=str(tmp_path), sim_seed=3, region_width_km=(16., 12.), resolution=res,
                 sim_mode="uniform", uniform_windspeed=10., uniform_winddirn=270., track_direction=0., track_count=32,
                 track_start_region=(2, 14, 0.5, 1))
    z1, z2 = synthetic_dem(rows, cols, res, seed=1), synthetic_dem(rows, cols, res, seed=2)

    def run(z, **kw):
        capsys.readouterr()
        sim = Simulator(cfg, elevation=z, **kw)
        sim.simulate_tracks()
        out = capsys.readouterr().out
        fname = sim._get_potential_fname("s10d270", 0, sim.mode_data_dir)
        return ("Found saved potential" in out, "Computing potential" in out, np.load(f"{fname}.npy"), fname,
                sim.presence_counts())

    found, computed, pot1, fname, pres1 = run(z1)
    assert computed and not found and os.path.exists(f"{fname}.key")
    found, computed, pot1b, _, pres1b = run(z1)
    assert found and not computed and np.array_equal(pot1b, pot1) and np.array_equal(pres1b, pres1)
    found, computed, pot2, _, _ = run(z2)                       # other terrain, same id string
    assert computed and not found and not np.array_equal(pot2, pot1)
    key2 = open(f"{fname}.key").read()
    found, computed, pot2b, _, _ = run(z2, force_potential=True)
    assert computed and not found and np.abs(pot2b - pot2).max() <= 1e-5 * 1000.0 and open(f"{fname}.key").read() == key2
    os.remove(f"{fname}.key")                                   # a raster the reference (or an older run) left behind
    found, computed, _, _, _ = run(z2)
    assert computed and not found and os.path.exists(f"{fname}.key")
