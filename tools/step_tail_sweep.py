"""Times the stepping kernel on the bench grid for several settings of the tail-prefetch / presence-hint knobs
(SSRS_STEP_PF_K, SSRS_STEP_PF_MINK, SSRS_STEP_RED_HINT) and track counts; checks that the presence raster is identical."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--tracks", type=int, nargs="+", default=[1024, 100_000, 1_000_000])
ap.add_argument("--cfg", type=str, nargs="+", default=["0,0,0", "4,0,0", "8,0,0", "16,0,0", "32,0,0", "8,15000,0", "16,15000,0", "0,0,1", "8,0,1", "16,15000,1"])
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
    for cfg in a.cfg:
        pfk, mink, hint = cfg.split(",")
        os.environ["SSRS_STEP_PF_K"] = pfk; os.environ["SSRS_STEP_PF_MINK"] = mink; os.environ["SSRS_STEP_RED_HINT"] = hint
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
        print(f"tracks {n} pf_k {pfk} mink {mink} hint {hint}: {best:.2f} ms, {int(total.item()) / best / 1e6:.2f} G track-steps/s, identical {same}", flush=True)
