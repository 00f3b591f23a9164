"""BASELINE configs[0] (500 x 600 at 100 m, 1000 northbound tracks) against outputs of the UNMODIFIED reference
(tests/golden/config1.npz, made by oracle/make_golden_config1.py: fields, two independent 1000-track realisations A and
B with per-track seeding, presence counts, lengths, the reference's smoothed map, 128 full trajectories).

Three claims (BASELINE.json north_star):
  * verification mode — fed the reference's own uniforms the stepper reproduces all 1000 tracks step for step:
    equal lengths, equal positions, bit-exact presence counts;
  * production mode, distributional — presence maps from an independent counter-based stream lie within
    1.5 x the reference-vs-reference noise floor of the normalised-L1 (total-variation) distance, raw and after the
    reference's disk smoothing (krad = 10), and the mean track length agrees within 1 %;
  * the potential the GPU solves from the reference's K agrees with the reference's SuperLU potential.
The CPU tests pin the oracles (C and numpy) to the fixture; the `gpu` tests do the same through the C-ABI.
"""
import numpy as np
import pytest

from oracle import oracle_c as OC
from oracle import oracle_np as O

L1_FACTOR = 1.5            # stated bound: distance to the reference <= 1.5 x reference-vs-reference distance
LEN_RTOL = 0.01            # mean track length within 1 %
SHAPE = (500, 600)


def tv(a, b):
    """Normalised-L1 distance of two presence rasters: sum |p - q| / 2 with p, q normalised to sum 1."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return 0.5 * np.abs(a / a.sum() - b / b.sum()).sum()


def reference_uniforms(base, lens):
    """The per-step uniforms the reference consumed: `np.random.seed(base + t)` then one random_sample() per step."""
    stride = int(lens.max())
    out = np.zeros((lens.size, stride))
    for t, n in enumerate(lens):
        out[t, :n - 1] = np.random.RandomState(int(base) + t).random_sample(int(n) - 1)
    return out


@pytest.fixture(scope="module")
def c1(golden):
    g = golden("config1")
    return {k: g[k] for k in g.files}


def _starts(c1):
    return np.stack([c1["start_rows"], c1["start_cols"]], 1).astype(np.int32)


def _check_verification(c1, lens, presence, traj):
    assert np.array_equal(lens, c1["A_len"])
    assert np.array_equal(presence.astype(np.int64), c1["A_presence"].astype(np.int64))
    ref = c1["A_traj"]
    for t in range(ref.shape[0]):
        n = c1["A_len"][t]
        assert np.array_equal(traj[t, :n], ref[t, :n]), t


def _check_distribution(c1, runs):
    """runs: list of (presence, lens) from independent production streams."""
    A, B = c1["A_presence"].astype(np.float64), c1["B_presence"].astype(np.float64)
    floor_raw = tv(A, B)
    sA, sB = O.smooth_presence(A, 10), O.smooth_presence(B, 10)
    floor_smooth = tv(sA, sB)
    ref_len = 0.5 * (c1["A_len"].mean() + c1["B_len"].mean())
    report = [f"reference vs reference: raw {floor_raw:.4f}, smoothed {floor_smooth:.4f}, mean length {ref_len:.1f}"]
    for pres, lens in runs:
        sP = O.smooth_presence(pres.astype(np.float64), 10)
        d_raw = max(tv(pres, A), tv(pres, B))
        d_smooth = max(tv(sP, sA), tv(sP, sB))
        report.append(f"production vs reference: raw {d_raw:.4f}, smoothed {d_smooth:.4f}, mean length {lens.mean():.1f}")
        assert d_raw <= L1_FACTOR * floor_raw, report
        assert d_smooth <= L1_FACTOR * floor_smooth, report
        assert abs(lens.mean() - ref_len) <= LEN_RTOL * ref_len, report
    print("\n".join(report))


def test_fixture_is_config1(c1):
    assert c1["U32"].shape == SHAPE and c1["P32"].dtype == np.float32
    assert c1["A_len"].size == 1000 and c1["B_len"].size == 1000
    # the reference's smoothed map (movmodel.py:422-439) from its own counts pins the numpy restatement at this size
    assert np.allclose(O.smooth_presence(c1["A_presence"].astype(np.float64), 10), c1["A_smooth10"], rtol=1e-5, atol=1e-7)


def test_c_oracle_step_for_step(c1):
    uni = reference_uniforms(c1["bases"][0], c1["A_len"])
    cap = int(c1["A_len"].max())
    out = OC.step_tracks(c1["U32"], c1["P32"], SHAPE, _starts(c1), 0.0, 1, 1.0, uniforms=uni, traj_cap=cap, nthreads=8)
    _check_verification(c1, out["traj_len"], out["presence"], out["traj"])
    assert out["total_steps"] == int((c1["A_len"] - 1).sum())


def test_c_oracle_distribution(c1):
    """The production arithmetic (what the GPU reproduces bit for bit) on Philox streams vs the reference."""
    runs = []
    for seed in (101, 202):
        out = OC.step_tracks(c1["U32"], c1["P32"], SHAPE, _starts(c1), 0.0, 1, 1.0, seed=seed, nthreads=8, fast=True)
        runs.append((out["presence"], out["traj_len"]))
    _check_distribution(c1, runs)


@pytest.mark.parametrize("dirn,mem,nu", [(0.0, 1, 1.0), (45.0, 1, 1.0), (0.0, 3, 1.0), (0.0, 1, 0.5)])
def test_production_arithmetic_picks_the_reference_moves(c1, dirn, mem, nu):
    """The link between the two parity statements: on the reference's own config-1 fields and IDENTICAL uniform streams
    (Philox, same seed) the production arithmetic (normalisations cancelled; what the GPU steps with, bit for bit) and the
    reference's exact operation order (pinned step for step by the reference's trajectories above) choose the same move
    at every step of every track — ~8e5 steps per run — so a production run IS the reference's algorithm
    (movmodel.py:264-318) on another random stream, not merely a process with a similar distribution."""
    for seed in (101, 202):
        exact = OC.step_tracks(c1["U32"], c1["P32"], SHAPE, _starts(c1), dirn, mem, nu, seed=seed, nthreads=8, fast=False)
        prod = OC.step_tracks(c1["U32"], c1["P32"], SHAPE, _starts(c1), dirn, mem, nu, seed=seed, nthreads=8, fast=True)
        assert exact["total_steps"] == prod["total_steps"] > 5e5
        assert np.array_equal(exact["traj_len"], prod["traj_len"])
        assert np.array_equal(exact["presence"], prod["presence"])


def test_production_arithmetic_picks_the_reference_moves_at_10m(golden):
    """The same at the resolution of the large configs (1000 x 1200 cells at 10 m, the refined-truth fixture's fields):
    two thirds of the float32 potential are plateaus there, so the fallback chain of movmodel.py:228-240 decides far
    more steps than at 100 m — and still every one of ~1.2e7 steps agrees."""
    g = golden("potential_truth10m")
    U, P = g["K32"], g["phi_truth32"]
    rng = np.random.RandomState(5)
    n = 2000
    starts = np.stack([rng.randint(2, 40, n), rng.randint(2, U.shape[1] - 2, n)], 1).astype(np.int32)
    for seed in (7, 8):
        for mem, nu in ((1, 1.0), (2, 1.0), (1, 0.3)):
            exact = OC.step_tracks(U, P, U.shape, starts, 0.0, mem, nu, seed=seed, nthreads=8, fast=False)
            prod = OC.step_tracks(U, P, U.shape, starts, 0.0, mem, nu, seed=seed, nthreads=8, fast=True)
            assert exact["total_steps"] == prod["total_steps"] > 1.5e6
            assert np.array_equal(exact["traj_len"], prod["traj_len"])
            assert np.array_equal(exact["presence"], prod["presence"])


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_step_for_step(c1):
    from ssrs_b200 import movmodel as mm
    uni = reference_uniforms(c1["bases"][0], c1["A_len"])
    cap = int(c1["A_len"].max())
    res = mm.simulate_tracks_batch(0.0, c1["start_rows"], c1["start_cols"], SHAPE, 1, 1.0, updraft_field=c1["U32"],
                                   potential_field=c1["P32"], uniforms=uni, record=True, traj_cap=cap)
    traj = res.traj.permute(1, 0, 2).cpu().numpy()
    _check_verification(c1, res.traj_len.cpu().numpy(), res.presence.cpu().numpy(), traj)
    assert res.total_steps == int((c1["A_len"] - 1).sum())


@pytest.mark.gpu
def test_gpu_distribution(c1):
    """Production mode (Philox streams, production arithmetic) vs the reference's two realisations."""
    from ssrs_b200 import movmodel as mm
    runs = []
    for seed in (101, 202):
        res = mm.simulate_tracks_batch(0.0, c1["start_rows"], c1["start_cols"], SHAPE, 1, 1.0, updraft_field=c1["U32"],
                                       potential_field=c1["P32"], seed=seed)
        runs.append((res.presence.cpu().numpy(), res.traj_len.cpu().numpy()))
    _check_distribution(c1, runs)


@pytest.mark.gpu
def test_gpu_potential_vs_reference(c1):
    """The reference's own K (float32) through the GPU solver vs the reference's SuperLU potential at config-1 size."""
    from ssrs_b200.potential import solve_potential_device
    phi, stats = solve_potential_device(c1["U32"], 0.0)
    d = np.abs(phi.cpu().numpy().astype(np.float64) - c1["P32"].astype(np.float64))
    ulp = np.spacing(np.float32(1000.0))
    assert stats["converged"] != 0
    assert d.max() <= 2 * ulp, (d.max() / ulp, stats)
