"""One stepping launch with few tracks (pure per-step latency regime) for ncu source-level profiling."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import movmodel as mm
class A: rows, cols, resolution, seed, no_solve, tracks_per_gpu = 5000, 6000, 10.0, 2021, False, 100000
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sr, sc = bench.start_cells(A, 100000)
up, pot, info = bench.build_fields_gpu(A, torch)
f = mm.interleave_fields(up, pot)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = mm.simulate_tracks_batch(0.0, sr[:n], sc[:n], (5000, 6000), fields=f, seed=2021)
    e1.record(); torch.cuda.synchronize()
L = res.traj_len.cpu().numpy().astype(np.int64) - 1
print(f"n={n} ms={e0.elapsed_time(e1):.3f} steps={L.sum()} maxlen={L.max()} us/step(longest)={e0.elapsed_time(e1)*1e3/L.max():.3f}")
