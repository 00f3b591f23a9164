"""Two stepping launches with many tracks per thread (bulk regime: lanes refill) for ncu; prints the second one's time."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import movmodel as mm
class A: rows, cols, resolution, seed, no_solve, tracks_per_gpu = 5000, 6000, 10.0, 2021, False, 100000
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
sr, sc = bench.start_cells(A, n)
up, pot, info = bench.build_fields_gpu(A, torch)
f = mm.interleave_fields(up, pot)
for _ in range(2):
    presence = torch.zeros((A.rows, A.cols), dtype=torch.int32, device="cuda"); total = torch.zeros(1, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mm.simulate_tracks_batch(0.0, sr, sc, (A.rows, A.cols), fields=f, seed=2021, presence=presence, total_steps=total)
    e1.record(); torch.cuda.synchronize()
print(f"n={n} ms={e0.elapsed_time(e1):.3f} steps={int(total.item())} steps/s={int(total.item())/e0.elapsed_time(e1)*1e3:.3e}")
