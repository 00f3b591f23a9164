"""On-disk track formats ("next" row f-4).

The reference pickles a Python list with one `int16 [L, 2]` array per track (`ssrs/simulator.py:383-386`) and every
consumer unpickles the whole list (`:407-408`, `:527-528`).  That is kept for small runs (same file name, same
content), but a million pickled arrays are impractical, so large runs use a packed form:

    <id>_tracks.npz :  offsets int64 [n_tracks + 1],  points int16 [offsets[-1], 2]   (row, col)

track t is `points[offsets[t]:offsets[t + 1]]`.  A multi-rank run writes one packed file per rank,
`<id>_tracks_part<r>of<w>.npz`, holding that rank's contiguous block of global track ids.  `load_tracks` reads any of
the three and returns the reference's list (parts concatenated in rank order).
"""
from __future__ import annotations

import os
import pickle
from typing import List, Sequence, Tuple

import numpy as np


def pack_tracks(tracks: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """list of int16 [L, 2] -> (offsets int64 [n + 1], points int16 [sum L, 2])."""
    lens = np.fromiter((len(t) for t in tracks), dtype=np.int64, count=len(tracks))
    offsets = np.zeros(len(tracks) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    points = np.empty((int(offsets[-1]), 2), dtype=np.int16)
    for t, tr in enumerate(tracks):
        a = np.asarray(tr)
        if a.ndim != 2 or a.shape[1] != 2:
            raise ValueError("every track must be an [L, 2] array of (row, col)")
        points[offsets[t]:offsets[t + 1]] = a
    return offsets, points


def unpack_tracks(offsets: np.ndarray, points: np.ndarray) -> List[np.ndarray]:
    """(offsets, points) -> list of int16 [L, 2] views, the reference's in-memory form."""
    offsets = np.asarray(offsets, dtype=np.int64)
    if offsets.ndim != 1 or offsets.size < 1 or offsets[0] != 0 or (np.diff(offsets) < 0).any():
        raise ValueError("offsets must start at 0 and be non-decreasing")
    if int(offsets[-1]) != len(points):
        raise ValueError("offsets[-1] does not match the number of points")
    return [points[offsets[t]:offsets[t + 1]] for t in range(offsets.size - 1)]


def save_tracks_packed(fname: str, offsets, points) -> str:
    path = fname if fname.endswith(".npz") else f"{fname}.npz"
    np.savez(path, offsets=np.asarray(offsets, dtype=np.int64), points=np.asarray(points, dtype=np.int16))
    return path


def save_tracks_pickle(fname: str, tracks: Sequence[np.ndarray]) -> str:
    """The reference's file: pickled list of int16 arrays (`simulator.py:385-386`)."""
    path = fname if fname.endswith(".pkl") else f"{fname}.pkl"
    with open(path, "wb") as fobj:
        pickle.dump([np.asarray(t, dtype=np.int16) for t in tracks], fobj)
    return path


def load_tracks(fname: str) -> List[np.ndarray]:
    """Reads `<fname>.pkl` (reference format) or `<fname>.npz` (packed); `fname` may carry either extension."""
    base = fname[:-4] if fname.endswith((".pkl", ".npz")) else fname
    if os.path.exists(f"{base}.pkl") and not fname.endswith(".npz"):
        with open(f"{base}.pkl", "rb") as fobj:
            return pickle.load(fobj)
    if os.path.exists(f"{base}.npz"):
        with np.load(f"{base}.npz") as z:
            return unpack_tracks(z["offsets"], z["points"])
    parts = part_files(base)
    if parts:
        out: List[np.ndarray] = []
        for path in parts:                          # rank order = global track-id order (block partition, dist.py)
            with np.load(path) as z:
                out.extend(unpack_tracks(z["offsets"], z["points"]))
        return out
    raise FileNotFoundError(f"no track file {base}.pkl, {base}.npz or {base}_part*of*.npz")


def part_files(base: str) -> List[str]:
    """The `<base>_part<r>of<w>.npz` files a multi-rank run writes (one per rank, Simulator.simulate_tracks), in rank
    order.  Raises if the set is incomplete or mixes world sizes; [] if there is none."""
    import glob
    import re
    found = {}
    for path in glob.glob(f"{glob.escape(base)}_part*of*.npz"):
        m = re.fullmatch(r"_part(\d+)of(\d+)\.npz", path[len(base):])
        if m:
            found[(int(m.group(1)), int(m.group(2)))] = path
    if not found:
        return []
    worlds = {w for _, w in found}
    if len(worlds) != 1:
        raise ValueError(f"track parts of different runs under {base}_part*: world sizes {sorted(worlds)}")
    world = worlds.pop()
    missing = [r for r in range(world) if (r, world) not in found]
    if missing:
        raise FileNotFoundError(f"track parts {missing} of {world} are missing under {base}_part*")
    return [found[(r, world)] for r in range(world)]
