"""Transition-table walk (ssrs_transition_table + ssrs_walk_tracks, ssrs_b200/csrc/walk.cu) through the C-ABI against
its CPU restatement (oracle/ssrs_oracle.c, mode "table": same thresholds, same word-to-step mapping) — bit-exact
lengths, presence and step totals — and against the reference's distribution at config-1 size."""
import numpy as np
import pytest

from oracle import oracle_c as OC
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


def _fields(rows, cols, res, seed=1):
    from ssrs_b200.synth import synthetic_dem
    z = synthetic_dem(rows, cols, res, seed=seed)
    _, _, _, K = O.updraft_pipeline(z, res, 10.0, 270.0, 0.75)
    return K.astype(np.float32)


def _same(res, ref):
    assert res.total_steps == ref["total_steps"]
    assert np.array_equal(res.traj_len.cpu().numpy(), ref["traj_len"])
    assert np.array_equal(res.presence.cpu().numpy(), ref["presence"])


@pytest.mark.parametrize("dirn", [0.0, 135.0, 270.0, 45.0])
def test_walk_matches_c_oracle(dirn):
    """3000 tracks on 200 x 240 cells, four directions (135: tracks start heading away from the direction, so the
    unmasked directional fallback of movmodel.py:239-240 is drawn often), default phase schedule and a schedule that
    cuts every 32 steps (dozens of compaction passes): identical to the oracle, and to each other."""
    from ssrs_b200 import movmodel as mm
    rows, cols = 200, 240
    U = _fields(rows, cols, 100.0)
    P = O.solve_potential(U.astype(np.float64), dirn)
    rng = np.random.RandomState(3)
    n = 3000
    starts = np.stack([rng.randint(0, rows, n), rng.randint(0, cols, n)], 1).astype(np.int32)     # border starts included
    ref = OC.step_tracks(U, P, (rows, cols), starts, dirn, 1, 1.0, seed=1234, track_id0=17, nthreads=8, fast="table")
    f = mm.interleave_fields(U, P)
    for first in (0, 32):
        res = mm.simulate_tracks_batch(dirn, starts[:, 0], starts[:, 1], (rows, cols), fields=f, seed=1234, track_id0=17,
                                       walk=True, first_phase_steps=first)
        _same(res, ref)


def test_table_entries_match_oracle_rule():
    """Spot check of the table itself: border strip marked, interior thresholds monotone, a flat patch (no candidate
    lower) carries the directional weights of its three candidates."""
    import torch
    from ssrs_b200 import movmodel as mm
    rows, cols = 40, 50
    rng = np.random.RandomState(0)
    U = (0.1 + rng.rand(rows, cols)).astype(np.float32)
    P = np.linspace(1000.0, 0.0, rows, dtype=np.float32)[:, None] + rng.rand(rows, cols).astype(np.float32)
    P[10:16, 10:20] = 500.0                                           # plateau: potential differences are exactly zero
    f = mm.interleave_fields(U, P)
    tab = mm.build_transition_table(f, 0.0).cpu().numpy().view(np.uint32).reshape(rows, cols, 8, 2)
    BORDER, UNMASKED, ONE = 0xFFFFFFFE, 0xFFFFFFFF, 0x80000000
    assert (tab[:2, :, :, 0] == BORDER).all() and (tab[-2:, :, :, 0] == BORDER).all()
    assert (tab[:, 0, :, 0] == BORDER).all() and (tab[:, -2:, :, 0] == BORDER).all()
    inner = tab[2:-2, 1:-2]
    ok = inner[..., 0] <= ONE
    assert ((inner[..., 0] == UNMASKED) | ok).all()
    assert (inner[..., 1][ok] >= inner[..., 0][ok]).all() and (inner[..., 1][ok] <= ONE).all()
    # plateau interior, previous move north (flat 7 -> slot 6): candidates NW, N, NE with directional weights
    # cos(45), 1, cos(45): thresholds 0.7071 / 2.4142 and 1.7071 / 2.4142 on the 2^-31 lattice
    e = tab[12, 14, 6]
    w = np.array([np.cos(np.pi / 4), 1.0, np.cos(np.pi / 4)])
    assert abs(int(e[0]) - w[0] / w.sum() * 2 ** 31) <= 1 and abs(int(e[1]) - (w[0] + w[1]) / w.sum() * 2 ** 31) <= 1
    # ... previous move south (flat 1 -> slot 1): candidates SW, S, SE have no directional weight northbound
    assert tab[12, 14, 1, 0] == UNMASKED


def test_walk_queue_limit_and_tiny_inputs():
    """More tracks than resident lanes (entries taken from the input counter by whichever lane is free) on a small grid
    where every track lives near the border strip (table mode entered and left all the time); tracks circling in a bowl
    until max_moves; no tracks, one track, the smallest grid."""
    from ssrs_b200 import movmodel as mm
    rows, cols = 120, 160
    U = _fields(rows, cols, 100.0, seed=4)
    n = 300_000
    rng = np.random.RandomState(11)
    starts = np.stack([rng.randint(2, rows - 2, n), rng.randint(2, cols - 2, n)], 1).astype(np.int32)
    P = O.solve_potential(U.astype(np.float64), 0.0)
    ref = OC.step_tracks(U, P, (rows, cols), starts, 0.0, 1, 1.0, seed=77, track_id0=5, nthreads=8, fast="table")
    _same(mm.simulate_tracks_batch(0.0, starts[:, 0], starts[:, 1], (rows, cols), updraft_field=U, potential_field=P,
                                   seed=77, track_id0=5, walk=True), ref)
    # bowl: max_moves
    rows, cols = 48, 64
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    P = (((yy - rows / 2) ** 2 + (xx - cols / 2) ** 2) * 0.5).astype(np.float32)
    rng = np.random.RandomState(2)
    U = (0.2 + rng.rand(rows, cols)).astype(np.float32)
    n = 4000
    starts = np.stack([rng.randint(2, rows - 2, n), rng.randint(2, cols - 2, n)], 1).astype(np.int32)
    ref = OC.step_tracks(U, P, (rows, cols), starts, 0.0, 1, 1.0, seed=3, nthreads=8, fast="table")
    kmax = int(np.ceil(rows / 2 * cols / 2))
    assert (ref["traj_len"] - 1 >= kmax).sum() > 100
    for first in (0, 100):
        res = mm.simulate_tracks_batch(0.0, starts[:, 0], starts[:, 1], (rows, cols), updraft_field=U, potential_field=P,
                                       seed=3, walk=True, first_phase_steps=first)
        _same(res, ref)
        assert int(res.traj_len.max().item()) == kmax + 1
    empty = mm.simulate_tracks_batch(0.0, starts[:0, 0], starts[:0, 1], (rows, cols), updraft_field=U, potential_field=P,
                                     seed=3, walk=True)
    assert empty.total_steps == 0 and int(empty.presence.sum().item()) == 0
    U5, P5 = U[:5, :5].copy(), np.ascontiguousarray(P[:5, :5])
    s5 = np.array([[2, 2], [1, 3], [3, 1]], dtype=np.int32)
    ref5 = OC.step_tracks(U5, P5, (5, 5), s5, 0.0, 1, 1.0, seed=9, nthreads=1, fast="table")
    _same(mm.simulate_tracks_batch(0.0, s5[:, 0], s5[:, 1], (5, 5), updraft_field=U5, potential_field=P5, seed=9, walk=True),
          ref5)
    with pytest.raises(ValueError):
        mm.simulate_tracks_batch(0.0, s5[:, 0], s5[:, 1], (5, 5), 3, 1.0, updraft_field=U5, potential_field=P5, walk=True)


def test_walk_sharding_invariance():
    from ssrs_b200 import movmodel as mm
    rows, cols = 200, 240
    U = _fields(rows, cols, 100.0, seed=2)
    P = O.solve_potential(U.astype(np.float64), 0.0)
    f = mm.interleave_fields(U, P)
    tab = mm.build_transition_table(f, 0.0)
    rng = np.random.RandomState(5)
    n = 4096
    sr, sc = rng.randint(2, 30, n), rng.randint(2, cols - 2, n)
    whole = mm.simulate_tracks_batch(0.0, sr, sc, (rows, cols), fields=f, seed=7, walk=True, table=tab)
    base_p, base_l = whole.presence.cpu().numpy(), whole.traj_len.cpu().numpy()
    for shards in (2, 8):
        pres, lens = None, []
        per = n // shards
        for s in range(shards):
            sl = slice(s * per, (s + 1) * per)
            r = mm.simulate_tracks_batch(0.0, sr[sl], sc[sl], (rows, cols), fields=f, seed=7, track_id0=s * per,
                                         presence=pres, walk=True, table=tab)
            pres = r.presence
            lens.append(r.traj_len.cpu().numpy())
        assert np.array_equal(pres.cpu().numpy(), base_p)
        assert np.array_equal(np.concatenate(lens), base_l)


def test_walk_config1_distribution(golden):
    """Config-1 fields from the unmodified reference: the walk reproduces its oracle bit for bit and the reference's
    presence distribution within the bounds of tests/test_config1_parity.py."""
    from ssrs_b200 import movmodel as mm
    from test_config1_parity import _check_distribution, SHAPE
    g = golden("config1")
    c1 = {k: g[k] for k in g.files}
    starts = np.stack([c1["start_rows"], c1["start_cols"]], 1).astype(np.int32)
    runs = []
    for seed in (101, 202):
        res = mm.simulate_tracks_batch(0.0, c1["start_rows"], c1["start_cols"], SHAPE, updraft_field=c1["U32"],
                                       potential_field=c1["P32"], seed=seed, walk=True)
        ref = OC.step_tracks(c1["U32"], c1["P32"], SHAPE, starts, 0.0, 1, 1.0, seed=seed, nthreads=8, fast="table")
        _same(res, ref)
        runs.append((res.presence.cpu().numpy(), res.traj_len.cpu().numpy()))
    _check_distribution(c1, runs)


def test_walk_full_size_properties():
    """BASELINE config 2 shape (5000 x 6000, 100k tracks): sum(presence) = steps + tracks, a 512-track sample equals
    the oracle, and the walk and the gather-and-evaluate stepper (different realisations of the same distribution)
    agree on the mean track length within 1 %."""
    import torch
    from ssrs_b200 import layers, movmodel as mm
    from ssrs_b200.synth import synthetic_dem
    rows, cols, res = 5000, 6000, 10.0
    z = torch.from_numpy(synthetic_dem(rows, cols, res)).cuda()
    up = layers.updraft_fields(z, res, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
    yy = torch.linspace(1000.0, 0.0, rows, device="cuda")[:, None]
    pot = (yy + 5.0 * torch.sin(torch.arange(cols, device="cuda")[None, :] / 97.0)).float().contiguous()
    f = mm.interleave_fields(up, pot)
    n = 100_000
    rng = np.random.RandomState(1)
    sr, sc = rng.randint(99, 200, n), rng.randint(506, 5489, n)
    assert mm.walk_pays_off(n, (rows, cols)) and not mm.walk_pays_off(1000, (500, 600))
    res_w = mm.simulate_tracks_batch(0.0, sr, sc, (rows, cols), fields=f, seed=99, walk=True)
    total = res_w.total_steps
    assert int(res_w.presence.sum(dtype=torch.int64).item()) == total + n
    lens = res_w.traj_len.cpu().numpy()
    assert lens.min() > 500 and total == int((lens.astype(np.int64) - 1).sum())
    m = 512
    ref = OC.step_tracks(up.cpu().numpy(), pot.cpu().numpy(), (rows, cols), np.stack([sr[:m], sc[:m]], 1), 0.0, 1, 1.0,
                         seed=99, want_presence=False, nthreads=8, fast="table")
    assert np.array_equal(lens[:m], ref["traj_len"])
    res_g = mm.simulate_tracks_batch(0.0, sr, sc, (rows, cols), fields=f, seed=99)
    assert abs(res_g.total_steps - total) <= 0.01 * total
