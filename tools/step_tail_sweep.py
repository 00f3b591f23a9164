"""Times the stepping kernel on the bench grid for several settings of SSRS_STEP_TAIL_LANES (tail-mode L1 prefetch) and
track counts; checks that the presence raster is identical for every setting."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--tracks", type=int, nargs="+", default=[32, 1024, 100_000, 1_000_000])
ap.add_argument("--lanes", type=int, nargs="+", default=[0, 1, 2, 4, 8, 16, 32])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
import bench
from ssrs_b200 import movmodel as mm
class A: rows, cols, resolution, seed, no_solve = 5000, 6000, 10.0, 2021, False
up, pot, info = bench.build_fields_gpu(A, torch)
fields = mm.interleave_fields(up, pot)
shape = (A.rows, A.cols)
for n in a.tracks:
    A.tracks_per_gpu = n
    sr, sc = bench.start_cells(A, n)
    ref = None
    for lanes in a.lanes:
        os.environ["SSRS_STEP_TAIL_LANES"] = str(lanes)
        best = 1e30
        for rep in range(a.reps):
            presence = torch.zeros(shape, dtype=torch.int32, device="cuda"); total = torch.zeros(1, dtype=torch.int64, device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            mm.simulate_tracks_batch(0.0, sr, sc, shape, fields=fields, seed=A.seed, track_id0=0, presence=presence, total_steps=total)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if ref is None: ref = presence.clone()
        same = bool((presence == ref).all().item())
        print(f"tracks {n} tail_lanes {lanes}: {best:.2f} ms, {int(total.item()) / best / 1e6:.2f} G track-steps/s, identical {same}", flush=True)
