"""Why is the ring of 100k-track batches (5.2e10 track-steps/s) slower than one 1M-track batch (6.9e10)?  Diagnostics on
the config-2 fields: K batches over S streams with (a) one presence raster per slot (what bench.py does), (b) ONE raster
shared by all slots (wrong maps, same work: isolates the footprint of the rasters), (c) batches issued in waves of S that
start together.  Usage: python tools/step_pipe_diag.py K S"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import movmodel as mm
class A: rows, cols, resolution, seed = 5000, 6000, 10.0, 2021
K, S = int(sys.argv[1]), int(sys.argv[2])
w = {"rows": A.rows, "cols": A.cols, "res": A.resolution, "wind": "uniform", "start_km": (5, 55, 1, 2)}
up, pot, info = bench.build_fields_gpu(A, torch, 1, w)
f = mm.interleave_fields(up, pot)
shape = (A.rows, A.cols); n = 100_000
rng = np.random.RandomState(1)
sr, sc = rng.randint(99, 200, n), rng.randint(506, 5489, n)
start = torch.from_numpy(np.stack([sr, sc], 1).astype(np.int32)).cuda()
wsb = int(mm.N.load().ssrs_walk_workspace_bytes(n))
def ev(): return torch.cuda.Event(enable_timing=True)
streams = [torch.cuda.Stream() for _ in range(S)]
ws = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(S)]
total = torch.zeros(1, dtype=torch.int64, device="cuda")
for name in ("own", "shared", "own_waves", "own"):
    pres = [torch.zeros(shape, dtype=torch.int32, device="cuda") for _ in range(S if name != "shared" else 1)]
    for rep in range(2):
        total.zero_(); torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for s in streams: s.wait_event(e0)
        for i in range(K):
            if name == "own_waves" and i % S == 0 and i:
                wave = ev(); 
                for s in streams: torch.cuda.current_stream().wait_stream(s)
                wave.record()
                for s in streams: s.wait_event(wave)
            with torch.cuda.stream(streams[i % S]):
                p = pres[i % len(pres)]
                if name != "shared": p.zero_()
                mm.simulate_tracks_batch(0.0, start, None, shape, fields=f, seed=3000 + i, presence=p, total_steps=total,
                                         phased=True, workspace=ws[i % S])
        for s in streams: torch.cuda.current_stream().wait_stream(s)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name:10s} K={K} S={S}: {ms:.1f} ms total, {ms/K:.2f} ms/batch, {int(total.item())/ms*1e3:.3e} steps/s", flush=True)
    del pres
