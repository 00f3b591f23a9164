"""Stage 3+4 on the GPU (through the C-ABI) vs the reference's golden trajectories and the C oracle."""
import numpy as np
import pytest

from oracle import oracle_c as OC
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

CASES = ["n0_m1", "n0_m3", "n0_m0", "d45_m1", "d270_m2", "n0_nu05", "n0_nu0"]


@pytest.mark.parametrize("name", CASES)
def test_golden_step_for_step(golden, name):
    """Verification mode: the reference's pre-drawn uniforms -> identical positions at every step,
    identical lengths, and bit-exact presence counts."""
    from ssrs_b200 import movmodel as mm
    g = golden("tracks")
    dirn, mem, nu = g[f"{name}_params"]
    lens = g[f"{name}_len"]
    cap = g[f"{name}_traj"].shape[1]
    starts = g[f"{name}_starts"]
    res = mm.simulate_tracks_batch(float(dirn), starts[:, 0], starts[:, 1], g["U32"].shape, int(mem), float(nu),
                                   updraft_field=g["U32"], potential_field=g["P32"], uniforms=g[f"{name}_uni"],
                                   record=True, traj_cap=cap)
    got_len = res.traj_len.cpu().numpy()
    traj = res.traj.permute(1, 0, 2).cpu().numpy()
    complete = lens <= cap
    assert complete.any()
    for t in np.flatnonzero(complete):
        assert got_len[t] == lens[t]
        assert np.array_equal(traj[t, :lens[t]], g[f"{name}_traj"][t, :lens[t]])
    if complete.all():
        assert np.array_equal(res.presence.cpu().numpy(), g[f"{name}_presence"].astype(np.int32))
        assert res.total_steps == int((lens - 1).sum())
        # compute_presence_counts from stored trajectories gives the same raster
        assert np.array_equal(mm.compute_presence_counts(res.tracks(), g["U32"].shape), res.presence.cpu().numpy())


def test_drw_philox():
    from ssrs_b200 import movmodel as mm
    n = 500
    rng = np.random.RandomState(1)
    starts = np.stack([rng.randint(2, 20, n), rng.randint(2, 98, n)], 1).astype(np.int32)
    for dirn in (0.0, 30.0, 200.0):
        ref = OC.step_tracks(None, None, (80, 100), starts, dirn, 1, 1.0, seed=5, fast=True)
        res = mm.simulate_tracks_batch(dirn, starts[:, 0], starts[:, 1], (80, 100), seed=5)
        assert np.array_equal(res.presence.cpu().numpy(), ref["presence"]) and res.total_steps == ref["total_steps"]


def test_drw_and_serial_api(golden):
    from ssrs_b200 import movmodel as mm
    g = golden("tracks")
    tj = g["drw_traj"]
    res = mm.simulate_tracks_batch(30.0, [5], [30], g["U32"].shape, 1, 1.0, uniforms=g["drw_uni"][None, :], record=True,
                                   traj_cap=len(tj) + 1)
    assert np.array_equal(res.tracks()[0], tj)
    # reference-signature call consuming numpy's global stream
    name = "n0_m1"
    t = 2
    np.random.seed(1000 + t)
    out = mm.generate_simulated_tracks(0.0, list(g[f"{name}_starts"][t]), g["U32"].shape, 1, 1.0, g["U32"], g["P32"])
    L = int(g[f"{name}_len"][t])
    assert out.dtype == np.int16 and np.array_equal(out, g[f"{name}_traj"][t, :L])
    after = np.random.random_sample()
    np.random.seed(1000 + t)
    np.random.random_sample(L - 1)
    assert after == np.random.random_sample()        # the global stream was advanced exactly like the reference


def _fields(rows, cols, res, seed=1):
    from ssrs_b200.synth import synthetic_dem
    z = synthetic_dem(rows, cols, res, seed=seed)
    _, _, _, K = O.updraft_pipeline(z, res, 10.0, 270.0, 0.75)
    return K.astype(np.float32)


@pytest.mark.parametrize("mem,nu,dirn", [(1, 1.0, 0.0), (2, 1.0, 315.0), (0, 1.0, 0.0), (1, 2.0, 90.0), (1, 0.0, 0.0)])
def test_philox_matches_c_oracle(mem, nu, dirn):
    """Production mode: Philox streams keyed by (seed, track id, step) -> the C oracle reproduces every
    trajectory and the presence raster bit for bit (3000 tracks on a 200x240 grid), in the production
    arithmetic and in the reference's exact operation order, and the two orders agree with each other."""
    from ssrs_b200 import movmodel as mm
    rows, cols = 200, 240
    U = _fields(rows, cols, 100.0)
    P = O.solve_potential(U.astype(np.float64), dirn)
    rng = np.random.RandomState(3)
    n = 3000
    starts = np.stack([rng.randint(2, 30, n), rng.randint(2, cols - 2, n)], 1).astype(np.int32)
    cap = 4 * max(rows, cols)
    ref = OC.step_tracks(U, P, (rows, cols), starts, dirn, mem, nu, seed=1234, track_id0=17, traj_cap=cap, nthreads=8,
                         fast=True)
    res = mm.simulate_tracks_batch(dirn, starts[:, 0], starts[:, 1], (rows, cols), mem, nu, updraft_field=U,
                                   potential_field=P, seed=1234, track_id0=17, record=True, traj_cap=cap)
    # the reference's exact operation order (numpy bit for bit) picks the same moves from the same streams
    ref_exact = OC.step_tracks(U, P, (rows, cols), starts, dirn, mem, nu, seed=1234, track_id0=17, nthreads=8)
    res_exact = mm.simulate_tracks_batch(dirn, starts[:, 0], starts[:, 1], (rows, cols), mem, nu, updraft_field=U,
                                         potential_field=P, seed=1234, track_id0=17, exact=True)
    assert np.array_equal(res_exact.presence.cpu().numpy(), ref_exact["presence"])
    assert np.array_equal(res_exact.traj_len.cpu().numpy(), ref_exact["traj_len"])
    assert np.array_equal(ref_exact["presence"], ref["presence"])
    assert res.total_steps == ref["total_steps"]
    assert np.array_equal(res.traj_len.cpu().numpy(), ref["traj_len"])
    assert np.array_equal(res.presence.cpu().numpy(), ref["presence"])
    traj = res.traj.permute(1, 0, 2).cpu().numpy()
    for t in range(0, n, 37):
        L = min(ref["traj_len"][t], cap)
        assert np.array_equal(traj[t, :L], ref["traj"][t, :L])
    # counts-only launch (no trajectory store): with memory 1 and nu 1 this takes the kernel's fast lane
    # (pairs of steps per Philox block); it must reproduce the same tracks
    res2 = mm.simulate_tracks_batch(dirn, starts[:, 0], starts[:, 1], (rows, cols), mem, nu, updraft_field=U,
                                    potential_field=P, seed=1234, track_id0=17)
    assert res2.total_steps == ref["total_steps"]
    assert np.array_equal(res2.traj_len.cpu().numpy(), ref["traj_len"])
    assert np.array_equal(res2.presence.cpu().numpy(), ref["presence"])


def test_track_queue_matches_c_oracle():
    """More tracks than the kernel has threads (148 SMs x 6 CTAs x 128 = 113 664): tracks beyond the first per thread
    are drawn from the device queue by whichever lane is free, and lanes re-enter the fast lane with fresh budgets all
    the time (a 120 x 160 grid keeps every track within a few cells of the border).  Per-track lengths, the presence
    raster and the step total still equal the C oracle's bit for bit, for two directions (northbound: ordinary
    candidates; 135 degrees: tracks start heading away from the direction, so the unmasked directional fallback of
    movmodel.py:239-240 is taken often)."""
    from ssrs_b200 import movmodel as mm
    rows, cols = 120, 160
    U = _fields(rows, cols, 100.0, seed=4)
    n = 300_000
    rng = np.random.RandomState(11)
    starts = np.stack([rng.randint(2, rows - 2, n), rng.randint(2, cols - 2, n)], 1).astype(np.int32)
    for dirn in (0.0, 135.0):
        P = O.solve_potential(U.astype(np.float64), dirn)
        ref = OC.step_tracks(U, P, (rows, cols), starts, dirn, 1, 1.0, seed=77, track_id0=5, nthreads=8, fast=True)
        res = mm.simulate_tracks_batch(dirn, starts[:, 0], starts[:, 1], (rows, cols), 1, 1.0, updraft_field=U,
                                       potential_field=P, seed=77, track_id0=5)
        assert res.total_steps == ref["total_steps"]
        assert np.array_equal(res.traj_len.cpu().numpy(), ref["traj_len"])
        assert np.array_equal(res.presence.cpu().numpy(), ref["presence"])


def test_move_limit_and_tiny_inputs():
    """Tracks that never leave: a bowl-shaped potential keeps many tracks circling until max_moves = rows/2 * cols/2
    (movmodel.py:277, :285) — the fast lane's budget has to stop them at exactly that step.  Lengths, presence and the
    step total equal the C oracle's; then the degenerate inputs: no tracks at all, one track, the smallest grid."""
    from ssrs_b200 import movmodel as mm
    rows, cols = 48, 64
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    P = (((yy - rows / 2) ** 2 + (xx - cols / 2) ** 2) * 0.5).astype(np.float32)
    rng = np.random.RandomState(2)
    U = (0.2 + rng.rand(rows, cols)).astype(np.float32)
    n = 4000
    starts = np.stack([rng.randint(2, rows - 2, n), rng.randint(2, cols - 2, n)], 1).astype(np.int32)
    ref = OC.step_tracks(U, P, (rows, cols), starts, 0.0, 1, 1.0, seed=3, nthreads=8, fast=True)
    kmax = int(np.ceil(rows / 2 * cols / 2))
    assert (ref["traj_len"] - 1 >= kmax).sum() > 100                   # the case does exercise the limit
    res = mm.simulate_tracks_batch(0.0, starts[:, 0], starts[:, 1], (rows, cols), 1, 1.0, updraft_field=U, potential_field=P,
                                   seed=3)
    assert res.total_steps == ref["total_steps"]
    assert np.array_equal(res.traj_len.cpu().numpy(), ref["traj_len"])
    assert np.array_equal(res.presence.cpu().numpy(), ref["presence"])
    assert int(res.traj_len.max().item()) == kmax + 1
    # no tracks: nothing happens, nothing fails
    empty = mm.simulate_tracks_batch(0.0, starts[:0, 0], starts[:0, 1], (rows, cols), 1, 1.0, updraft_field=U,
                                     potential_field=P, seed=3)
    assert empty.total_steps == 0 and int(empty.presence.sum().item()) == 0
    # one track; and the smallest grid the ABI accepts (5 x 5: every cell is next to the border)
    one = mm.simulate_tracks_batch(0.0, starts[:1, 0], starts[:1, 1], (rows, cols), 1, 1.0, updraft_field=U, potential_field=P,
                                   seed=3)
    assert int(one.traj_len[0].item()) == int(ref["traj_len"][0])
    U5, P5 = U[:5, :5].copy(), np.ascontiguousarray(P[:5, :5])
    s5 = np.array([[2, 2], [1, 3], [3, 1]], dtype=np.int32)
    ref5 = OC.step_tracks(U5, P5, (5, 5), s5, 0.0, 1, 1.0, seed=9, nthreads=1, fast=True)
    got5 = mm.simulate_tracks_batch(0.0, s5[:, 0], s5[:, 1], (5, 5), 1, 1.0, updraft_field=U5, potential_field=P5, seed=9)
    assert np.array_equal(got5.traj_len.cpu().numpy(), ref5["traj_len"])
    assert np.array_equal(got5.presence.cpu().numpy(), ref5["presence"])


def test_sharding_invariance():
    """Tracks block-partitioned over 1/2/4/8 shards (what each GPU of a box would run) give bit-identical
    summed presence and per-track lengths: the RNG is keyed by the global track id."""
    from ssrs_b200 import movmodel as mm
    rows, cols = 200, 240
    U = _fields(rows, cols, 100.0, seed=2)
    P = O.solve_potential(U.astype(np.float64), 0.0)
    f = mm.interleave_fields(U, P)
    rng = np.random.RandomState(5)
    n = 4096
    sr, sc = rng.randint(2, 30, n), rng.randint(2, cols - 2, n)
    whole = mm.simulate_tracks_batch(0.0, sr, sc, (rows, cols), fields=f, seed=7)
    base_p, base_l = whole.presence.cpu().numpy(), whole.traj_len.cpu().numpy()
    for shards in (2, 4, 8):
        pres, lens = None, []
        per = n // shards
        for s in range(shards):
            sl = slice(s * per, (s + 1) * per)
            r = mm.simulate_tracks_batch(0.0, sr[sl], sc[sl], (rows, cols), fields=f, seed=7, track_id0=s * per,
                                         presence=pres)
            pres = r.presence
            lens.append(r.traj_len.cpu().numpy())
        assert np.array_equal(pres.cpu().numpy(), base_p)
        assert np.array_equal(np.concatenate(lens), base_l)


def test_full_size_properties():
    """BASELINE config 2 shape: 100k tracks on a (5000, 6000) grid.  Properties that do not need the
    oracle at full size: sum(presence) == total_steps + n_tracks (every appended point is counted once);
    a 512-track sample reproduces the C oracle exactly; every track ends on the border or at max_moves."""
    import torch
    from ssrs_b200 import layers, movmodel as mm
    from ssrs_b200.synth import synthetic_dem
    rows, cols, res = 5000, 6000, 10.0
    z = torch.from_numpy(synthetic_dem(rows, cols, res)).cuda()
    up = layers.updraft_fields(z, res, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
    # a smooth stand-in potential (north = 0, south = 1000 plus relief); stage 2 has its own tests
    yy = torch.linspace(1000.0, 0.0, rows, device="cuda")[:, None]
    pot = (yy + 5.0 * torch.sin(torch.arange(cols, device="cuda")[None, :] / 97.0)).float().contiguous()
    f = mm.interleave_fields(up, pot)
    n = 100_000
    rng = np.random.RandomState(1)
    sr, sc = rng.randint(99, 200, n), rng.randint(506, 5489, n)
    res_gpu = mm.simulate_tracks_batch(0.0, sr, sc, (rows, cols), fields=f, seed=99)
    total = res_gpu.total_steps
    assert int(res_gpu.presence.sum(dtype=torch.int64).item()) == total + n
    lens = res_gpu.traj_len.cpu().numpy()
    assert lens.min() > 500 and total == int((lens.astype(np.int64) - 1).sum())
    m = 512
    ref = OC.step_tracks(up.cpu().numpy(), pot.cpu().numpy(), (rows, cols), np.stack([sr[:m], sc[:m]], 1), 0.0, 1, 1.0,
                         seed=99, want_presence=False, nthreads=8, fast=True)
    assert np.array_equal(lens[:m], ref["traj_len"])


def test_errors():
    from ssrs_b200 import movmodel as mm
    from ssrs_b200._native import NativeError
    U = np.ones((50, 60), np.float32)
    with pytest.raises(ValueError):
        mm.simulate_tracks_batch(0.0, [70], [5], (50, 60), updraft_field=U, potential_field=U)
    with pytest.raises(NativeError):
        mm.simulate_tracks_batch(0.0, [7], [5], (50, 60), memory_parameter=99, updraft_field=U, potential_field=U)
    with pytest.raises(ValueError):
        mm.get_starting_indices(10, (60, 5, 1, 2), "random", (60., 50.), 100.)
    with pytest.raises(ValueError):
        mm.get_starting_indices(10, (5, 55, 1, 2), "spiral", (60., 50.), 100.)
    r = mm.simulate_tracks_batch(0.0, [], [], (50, 60), updraft_field=U, potential_field=U)
    assert r.total_steps == 0 and int(r.presence.sum()) == 0


@pytest.mark.parametrize("mem,nu,dirn,exact", [(1, 1.0, 0.0, False), (1, 1.0, 135.0, False), (1, 2.0, 90.0, False),
                                               (1, 1.0, 0.0, True), (2, 1.0, 315.0, False)])
def test_phased_launch_is_bit_identical(mem, nu, dirn, exact):
    """ssrs_step_tracks_phased (survivors compacted every few steps — a cut every 32 steps here, dozens of phases — and
    the default schedule) against the C oracle and the single launch: same lengths, presence, step total, and with
    recording the same trajectories.  memory 2 runs as one launch behind the same entry point."""
    from ssrs_b200 import movmodel as mm
    rows, cols = 200, 240
    U = _fields(rows, cols, 100.0)
    P = O.solve_potential(U.astype(np.float64), dirn)
    rng = np.random.RandomState(3)
    n = 3000
    starts = np.stack([rng.randint(0, rows, n), rng.randint(0, cols, n)], 1).astype(np.int32)
    ref = OC.step_tracks(U, P, (rows, cols), starts, dirn, mem, nu, seed=1234, track_id0=17, nthreads=8, fast=not exact)
    f = mm.interleave_fields(U, P)
    for first in (32, 0):
        res = mm.simulate_tracks_batch(dirn, starts[:, 0], starts[:, 1], (rows, cols), mem, nu, fields=f, seed=1234,
                                       track_id0=17, exact=exact, phased=True, first_phase_steps=first)
        assert res.total_steps == ref["total_steps"]
        assert np.array_equal(res.traj_len.cpu().numpy(), ref["traj_len"])
        assert np.array_equal(res.presence.cpu().numpy(), ref["presence"])
    cap = int(ref["traj_len"].max())
    rec = mm.simulate_tracks_batch(dirn, starts[:, 0], starts[:, 1], (rows, cols), mem, nu, fields=f, seed=1234,
                                   track_id0=17, exact=exact, phased=True, first_phase_steps=32, record=True, traj_cap=cap)
    one = mm.simulate_tracks_batch(dirn, starts[:, 0], starts[:, 1], (rows, cols), mem, nu, fields=f, seed=1234,
                                   track_id0=17, exact=exact, record=True, traj_cap=cap)
    assert np.array_equal(rec.traj.cpu().numpy(), one.traj.cpu().numpy())
