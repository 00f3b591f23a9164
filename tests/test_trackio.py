"""On-disk track formats: the reference's pickled list and the packed offsets/points form."""
import pickle

import numpy as np
import pytest

from ssrs_b200 import trackio


def _tracks(rng, n):
    return [rng.randint(0, 500, (rng.randint(1, 40), 2)).astype(np.int16) for _ in range(n)]


def test_pack_roundtrip_and_formats(tmp_path):
    rng = np.random.RandomState(0)
    tracks = _tracks(rng, 57)
    off, pts = trackio.pack_tracks(tracks)
    assert off.dtype == np.int64 and pts.dtype == np.int16 and off[0] == 0 and off[-1] == sum(len(t) for t in tracks)
    back = trackio.unpack_tracks(off, pts)
    assert len(back) == len(tracks) and all(np.array_equal(a, b) for a, b in zip(back, tracks))
    # packed file
    p = trackio.save_tracks_packed(str(tmp_path / "a_tracks"), off, pts)
    assert p.endswith(".npz")
    got = trackio.load_tracks(str(tmp_path / "a_tracks"))
    assert all(np.array_equal(a, b) for a, b in zip(got, tracks))
    # the reference's pickle: a plain list of int16 arrays that the reference's own readers unpickle
    q = trackio.save_tracks_pickle(str(tmp_path / "b_tracks"), tracks)
    with open(q, "rb") as f:
        raw = pickle.load(f)
    assert isinstance(raw, list) and raw[0].dtype == np.int16 and raw[0].shape[1] == 2
    assert all(np.array_equal(a, b) for a, b in zip(trackio.load_tracks(str(tmp_path / "b_tracks.pkl")), tracks))
    # empty list and ragged extremes
    off0, pts0 = trackio.pack_tracks([])
    assert off0.tolist() == [0] and pts0.shape == (0, 2) and trackio.unpack_tracks(off0, pts0) == []
    with pytest.raises(ValueError):
        trackio.unpack_tracks(np.array([0, 5]), np.zeros((4, 2), np.int16))
    with pytest.raises(ValueError):
        trackio.pack_tracks([np.zeros((3, 3), np.int16)])
    with pytest.raises(FileNotFoundError):
        trackio.load_tracks(str(tmp_path / "missing"))


@pytest.mark.gpu
def test_packed_from_device_matches_list():
    from ssrs_b200.movmodel import simulate_tracks_batch
    rng = np.random.RandomState(1)
    rows, cols, n = 90, 70, 300
    U = (rng.rand(rows, cols) * (rng.rand(rows, cols) > 0.4)).astype(np.float32)
    P = (np.linspace(1000, 0, rows, dtype=np.float32)[:, None] + rng.rand(rows, cols).astype(np.float32))
    res = simulate_tracks_batch(0.0, rng.randint(2, 10, n), rng.randint(2, cols - 2, n), (rows, cols), updraft_field=U,
                                potential_field=P, seed=4, record=True, traj_cap=2000)
    off, pts = res.packed()
    lst = res.tracks()
    back = trackio.unpack_tracks(off, pts)
    assert len(back) == n and all(np.array_equal(a, b) for a, b in zip(back, lst))


def test_multi_rank_parts(tmp_path):
    """`<id>_tracks_part<r>of<w>.npz` written by the ranks of a sharded run load as one list in global-id order."""
    rng = np.random.RandomState(3)
    tracks = _tracks(rng, 41)
    base = str(tmp_path / "c_tracks")
    bounds = [0, 14, 28, 41]
    for r in range(3):
        off, pts = trackio.pack_tracks(tracks[bounds[r]:bounds[r + 1]])
        trackio.save_tracks_packed(f"{base}_part{r}of3", off, pts)
    got = trackio.load_tracks(base)
    assert len(got) == 41 and all(np.array_equal(a, b) for a, b in zip(got, tracks))
    import os
    os.remove(f"{base}_part1of3.npz")
    with pytest.raises(FileNotFoundError):
        trackio.load_tracks(base)
