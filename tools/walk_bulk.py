"""One transition-table walk of many tracks (bulk regime) for ncu; prints its time.  Usage: walk_bulk.py [n_tracks]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import movmodel as mm
class A: rows, cols, resolution, seed, no_solve, tracks_per_gpu = 5000, 6000, 10.0, 2021, False, 100000
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sr, sc = bench.start_cells(A, n)
up, pot, info = bench.build_fields_gpu(A, torch)
f = mm.interleave_fields(up, pot)
tab = mm.build_transition_table(f, 0.0)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = mm.simulate_tracks_batch(0.0, sr, sc, (A.rows, A.cols), fields=f, seed=2021, walk=True, table=tab)
    e1.record(); torch.cuda.synchronize()
print(f"n={n} ms={e0.elapsed_time(e1):.3f} steps={res.total_steps} steps/s={res.total_steps/e0.elapsed_time(e1)*1e3:.3e}")
