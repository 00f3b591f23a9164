"""TEST INFRASTRUCTURE ONLY — writes tests/golden/config1.npz by running the UNMODIFIED reference at the size of
BASELINE.json configs[0] (500 x 600 cells at 100 m, wind 10 m/s from 270 deg, 1000 northbound tracks).

Run in the authoring container (needs /root/reference):  python -m oracle.make_golden_config1
Everything stored is an output of `/root/reference/ssrs/{layers,movmodel}.py` (loaded by oracle/ref_loader.py); the
glue restates `ssrs/simulator.py:189-198,230-243,259-288,339-369` (float32 save/reload of the orograph, threshold,
MovModel calls, `get_starting_indices` after `np.random.seed`).  Tracks are generated with per-track seeding
(`np.random.seed(base + t)` before each `generate_simulated_tracks` call — SURVEY.md §0 finding 5: the reference's
fork pool replays identical streams, so only serial/per-track seeding is reproducible); the per-step uniforms are
NOT stored: a test re-draws them with `np.random.RandomState(base + t).random_sample(len - 1)`, the same stream.

Two independent realisations A (base 10_000) and B (base 20_000) of the same 1000 start cells give the
reference-vs-reference noise floor of the distributional parity metric.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import reference_fields  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402
from ssrs_b200.synth import synthetic_dem  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "config1.npz")
ROWS, COLS, RES = 500, 600, 100.0
NTRACKS = 1000
BASE = {"A": 10_000, "B": 20_000}
KEEP_TRAJ = 128            # full reference trajectories stored for the first tracks of realisation A
_G = {}


def _one(args):
    base, t, r, c = args
    M, U, phi = _G["M"], _G["U"], _G["phi"]
    np.random.seed(base + t)
    return M.generate_simulated_tracks(0.0, [int(r), int(c)], U.shape, 1, 1.0, U, phi)


def main():
    import multiprocessing as mp
    L, M = load_reference()
    t0 = time.time()
    z = synthetic_dem(ROWS, COLS, RES)
    slope, aspect, oro, K, phi = reference_fields(L, M, z, RES, 10.0, 270.0, 0.75, 0.0)
    print(f"reference fields: {time.time() - t0:.1f} s; K dtype {K.dtype}, zero fraction {(K == 0).mean():.3f}")
    U32 = np.asarray(K, dtype=np.float32)
    U = U32.astype(np.float64)                  # both sides step on exactly these values (benign)
    np.random.seed(1)
    sr, sc = M.get_starting_indices(NTRACKS, (5, 55, 1, 2), 'random', (60., 50.), RES)      # simulator.py:339-345
    _G.update(M=M, U=U, phi=phi)
    out = dict(U32=U32, P32=phi, start_rows=sr.astype(np.int32), start_cols=sc.astype(np.int32),
               bases=np.array([BASE["A"], BASE["B"]]))
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        for name, base in BASE.items():
            t0 = time.time()
            tracks = pool.map(_one, [(base, t, sr[t], sc[t]) for t in range(NTRACKS)], chunksize=8)
            lens = np.array([len(t) for t in tracks], dtype=np.int32)
            print(f"realisation {name}: {lens.sum() - NTRACKS} track-steps in {time.time() - t0:.1f} s, "
                  f"length min/mean/max {lens.min()}/{lens.mean():.1f}/{lens.max()}")
            out[f"{name}_len"] = lens
            out[f"{name}_presence"] = M.compute_presence_counts(tracks, U.shape)                # int16, movmodel.py:410-419
            if name == "A":
                cap = int(lens[:KEEP_TRAJ].max())
                tr = np.zeros((KEEP_TRAJ, cap, 2), dtype=np.int16)
                for t in range(KEEP_TRAJ):
                    tr[t, :lens[t]] = tracks[t]
                out["A_traj"] = tr
                out["A_smooth10"] = M.compute_smooth_presence_counts(tracks, U.shape, 10).astype(np.float32)   # :422-439
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
