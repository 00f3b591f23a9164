"""The oracle (oracle/oracle_np.py, oracle/ssrs_oracle.c) against outputs of the unmodified reference
stored in tests/golden/ (made by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle_c as OC
from oracle import oracle_np as O


def test_tables(golden):
    g = golden("tables")
    for i in range(9):
        assert np.array_equal(O.track_restrictions(i // 3 - 1, i % 3 - 1), g["masks"][i])
    for th, w in zip(g["thetas"], g["dirw"]):
        assert np.array_equal(O.directional_probs(th * np.pi / 180.0), w)
    assert np.allclose(O.get_above_threshold_speed(g["thr_x"], 0.75), g["thr_y"], rtol=1e-14, atol=0)
    assert np.array_equal(O.NORMS_INV, g["norms_inv"])
    assert np.array_equal(np.array(O.NEIGHBOUR_DELTAS), g["deltas"])
    for th in (0, 30, 45, 90, 180, 270, 315, -45):
        for shp in ((23, 31), (60, 50)):
            bn, be = O.boundary_nodes(th, *shp)
            assert np.array_equal(bn, g[f"bn_{th}_{shp[0]}x{shp[1]}"])
            assert np.array_equal(be, g[f"be_{th}_{shp[0]}x{shp[1]}"])
    r, c = O.starting_indices(64, (5, 55, 1, 2), "random", (60., 50.), 100., rng=np.random.RandomState(4))
    assert np.array_equal(r, g["start_random_rows"]) and np.array_equal(c, g["start_random_cols"])
    r, c = O.starting_indices(37, (5, 55, 1, 2), "structured", (60., 50.), 100.)
    assert np.array_equal(r, g["start_struct_rows"]) and np.array_equal(c, g["start_struct_cols"])


@pytest.mark.parametrize("name", ["a", "b"])
def test_stencil(golden, name):
    g = golden("stencil")
    z, res = g[f"{name}_z"], float(g[f"{name}_res"])
    sl, asp, oro, K = O.updraft_pipeline(z, res, 10.0, 270.0, 0.75)
    assert np.array_equal(sl, g[f"{name}_slope"]) and np.array_equal(asp, g[f"{name}_aspect"])
    assert np.array_equal(oro, g[f"{name}_oro"])
    assert np.allclose(K, g[f"{name}_K"], rtol=1e-12, atol=1e-15)   # exp()-1 cancels: scalar vs SIMD exp differ by an ulp
    assert (sl[0] == 0).all() and (sl[:, -1] == 0).all() and (asp[-1] == 0).all()
    assert (asp[6:8, 8:11] == 270.0).all()          # flat patch: dz_dx == 0 -> 1e-10 -> aspect 270
    _, _, oro2, K2 = O.updraft_pipeline(z, res, g[f"{name}_ws"], g[f"{name}_wd"], 0.75)
    assert np.array_equal(oro2, g[f"{name}_oro_cell"])
    assert np.allclose(K2, g[f"{name}_K_cell"], rtol=1e-12, atol=1e-15)


def test_potential(golden):
    g = golden("potential")
    ulp = np.spacing(np.float32(1000.0))
    for th in (0, 90, 180, 270, 45, -45, 30):
        phi = O.solve_potential(g["rand_K"], th)
        assert np.abs(phi.astype(np.float64) - g[f"rand_phi_{th}"]).max() <= 2 * ulp, th
    for th in (0, 270, 45):
        phi = O.solve_potential(g["dem_K"], th)
        assert np.abs(phi.astype(np.float64) - g[f"dem_phi_{th}"]).max() <= 2 * ulp, th
    phi = O.solve_potential(g["dem2_K"].astype(np.float64), 0.0)
    assert np.abs(phi.astype(np.float64) - g["dem2_phi_0"]).max() <= 2 * ulp


def test_operator_residual(golden):
    """The [row, col] restatement of the operator annihilates the reference's own potential at free nodes."""
    g = golden("potential")
    K = g["rand_K"]
    gw = O.edge_weights(K)
    for th in (0, 45, 270):
        phi = O.solve_potential(K, th).astype(np.float64)
        mask, _ = O.boundary_grid(th, *K.shape)
        res = O.apply_operator(gw, phi)
        scale = gw.sum(axis=0) * 1000.0
        assert (np.abs(res[~mask]) / scale[~mask]).max() < 1e-6      # float32 rounding of phi


CASES = ["n0_m1", "n0_m3", "n0_m0", "d45_m1", "d270_m2", "n0_nu05", "n0_nu0"]


@pytest.mark.parametrize("name", CASES)
def test_tracks_c_oracle(golden, name):
    g = golden("tracks")
    dirn, mem, nu = g[f"{name}_params"]
    U32, P32 = g["U32"], g["P32"]
    lens = g[f"{name}_len"]
    cap = g[f"{name}_traj"].shape[1]
    out = OC.step_tracks(U32, P32, U32.shape, g[f"{name}_starts"], float(dirn), int(mem), float(nu),
                         uniforms=g[f"{name}_uni"], traj_cap=cap)
    for t in range(len(lens)):
        if lens[t] <= cap:
            assert out["traj_len"][t] == lens[t]
            assert np.array_equal(out["traj"][t, :lens[t]], g[f"{name}_traj"][t, :lens[t]])
    if (lens <= cap).all():
        assert np.array_equal(out["presence"], g[f"{name}_presence"].astype(np.int32))


def test_tracks_numpy_oracle(golden):
    g = golden("tracks")
    U = g["U32"].astype(np.float64)
    for name in ("n0_m1", "d45_m1", "n0_m0"):
        dirn, mem, nu = g[f"{name}_params"]
        for t in range(3):
            L = int(g[f"{name}_len"][t])
            tj = O.simulate_track(float(dirn), g[f"{name}_starts"][t], U.shape, int(mem), float(nu), U, g["P32"],
                                  g[f"{name}_uni"][t])
            assert np.array_equal(tj, g[f"{name}_traj"][t, :L])


def test_drw(golden):
    g = golden("tracks")
    tj = g["drw_traj"]
    out = OC.step_tracks(None, None, g["U32"].shape, np.array([[5, 30]], dtype=np.int32), 30.0, 1, 1.0,
                         uniforms=g["drw_uni"][None, :], traj_cap=len(tj) + 1)
    assert out["traj_len"][0] == len(tj) and np.array_equal(out["traj"][0, :len(tj)], tj)


def test_presence_and_smoothing(golden):
    g = golden("tracks")
    s = golden("smooth")
    name = "n0_nu0"
    lens = g[f"{name}_len"]
    cnt = OC.presence_counts(g[f"{name}_traj"], np.minimum(lens, g[f"{name}_traj"].shape[1]), g["U32"].shape)
    tracks = [g[f"{name}_traj"][t, :lens[t]] for t in range(len(lens))]
    assert np.array_equal(O.presence_counts(tracks, g["U32"].shape), cnt)
    assert np.array_equal(cnt, s["counts"].astype(np.int32))
    for rad in (2, 5):
        assert np.allclose(O.smooth_presence(cnt, rad), s[f"smooth_{rad}"], rtol=1e-6, atol=1e-7)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors): counter/key all zero ->
    6627e8d5 e169c58d bc57ac4c 9b00dbd8; all ones -> 408f276d 41c83b0e a20bc7c6 6d5451fd;
    counter 243f6a88 85a308d3 13198a2e 03707344, key a4093822 299f31d0 -> d16cfe09 94fdcceb 5001e420 24126ea1."""
    assert [hex(v) for v in OC.philox_words(0, 0, 0)] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    # step -> (block = step >> 1, word pair = step & 1), 52 mantissa bits
    for step, (a, b) in enumerate([(0x6627e8d5, 0xe169c58d), (0xbc57ac4c, 0x9b00dbd8)]):
        bits = 0x3FF0000000000000 | (a << 20) | (b >> 12)
        expect = np.array([bits], dtype=np.uint64).view(np.float64)[0] - 1.0
        assert OC.philox_uniform(0, 0, step) == expect
    u = np.array([OC.philox_uniform(7, t, s) for t in range(50) for s in range(40)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.03
