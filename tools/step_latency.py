"""Per-step latency of the stepping kernel (kernel time / longest track) and bulk throughput, on the bench fields."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import movmodel as mm

class A: pass
a = A(); a.rows = 5000; a.cols = 6000; a.resolution = 10.0; a.tracks_per_gpu = 100000; a.seed = 2021; a.no_solve = False
sr, sc = bench.start_cells(a, 1_000_000)
up, pot, info = bench.build_fields_gpu(a, torch)
f = mm.interleave_fields(up, pot)
def run(n, exact=False, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = mm.simulate_tracks_batch(0.0, sr[:n], sc[:n], (5000, 6000), fields=f, seed=2021, exact=exact)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    L = res.traj_len.cpu().numpy().astype(np.int64) - 1
    print(f"n={n:8d} exact={exact} ms={best:9.3f} steps={L.sum():12d} maxlen={L.max():7d} us/step(longest)={best*1e3/L.max():6.3f} "
          f"steps/s={L.sum()/best*1e3:.3e} p99={int(np.percentile(L,99))}")
for n in (32, 1024, 100_000, 1_000_000):
    run(n)
run(100_000, exact=True, reps=1)
