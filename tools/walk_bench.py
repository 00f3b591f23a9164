"""Transition-table walk on the bench fields: table build time, one launch alone (100k / 1M tracks), and K batches of
100k tracks pipelined over S streams (what bench.py times).  Usage: python tools/walk_bench.py [K] [S]"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import movmodel as mm

class A: rows, cols, resolution, seed, no_solve, tracks_per_gpu = 5000, 6000, 10.0, 2021, False, 100000
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sr, sc = bench.start_cells(A, 1_000_000)
up, pot, info = bench.build_fields_gpu(A, torch)
print("potential ms", info.get("potential_ms"), info.get("potential_stats", {}).get("iterations"))
f = mm.interleave_fields(up, pot)
shape = (A.rows, A.cols)
def ev(): return torch.cuda.Event(enable_timing=True)
# table build
tab = mm.build_transition_table(f, 0.0)
torch.cuda.synchronize()
e0, e1 = ev(), ev(); e0.record()
for _ in range(5): mm.build_transition_table(f, 0.0, out=tab)
e1.record(); torch.cuda.synchronize()
print(f"table build {e0.elapsed_time(e1)/5:.3f} ms  ({tab.numel()/1e9:.2f} GB)")
def one(n, walk, seed=2021, reps=2):
    best = 1e9
    for _ in range(reps):
        e0, e1 = ev(), ev(); e0.record()
        res = mm.simulate_tracks_batch(0.0, sr[:n], sc[:n], shape, fields=f, seed=seed, walk=walk, table=tab if walk else None)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    L = res.traj_len.cpu().numpy().astype(np.int64) - 1
    print(f"n={n:8d} walk={walk} ms={best:9.3f} steps={L.sum():12d} maxlen={L.max():7d} us/step(longest)={best*1e3/L.max():6.3f} steps/s={L.sum()/best*1e3:.3e}")
    return res
for n in (32, 1024, 100_000, 1_000_000):
    one(n, True)
one(100_000, False); one(1_000_000, False, reps=1)
# pipelined batches
n = 100_000
for S in (1, 2, 3, 4, 6):
    streams = [torch.cuda.Stream() for _ in range(S)]
    pres = [torch.zeros(shape, dtype=torch.int32, device="cuda") for _ in range(S)]
    ws = [torch.empty(int(mm.N.load().ssrs_walk_workspace_bytes(n)), dtype=torch.uint8, device="cuda") for _ in range(S)]
    tabs = [torch.empty_like(tab) for _ in range(min(S, 3))]
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for with_table in (False, True):
        total.zero_(); torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for s in streams: s.wait_event(e0)
        for i in range(K):
            s = streams[i % S]
            with torch.cuda.stream(s):
                pres[i % S].zero_()
                t = tabs[i % len(tabs)] if with_table else tab
                if with_table: mm.build_transition_table(f, 0.0, out=t)
                mm.simulate_tracks_batch(0.0, sr[:n], sc[:n], shape, fields=f, seed=3000 + i, walk=True, table=t,
                                         workspace=ws[i % S], presence=pres[i % S], total_steps=total)
        for s in streams: torch.cuda.current_stream().wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"pipelined K={K} S={S} table_per_batch={with_table}: {ms:.1f} ms total, {ms/K:.2f} ms/batch, {int(total.item())/ms*1e3:.3e} steps/s")
