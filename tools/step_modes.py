"""Stepping on the bench fields in its three forms — one launch per batch (gather), phased gather, transition-table walk —
alone and with K batches of 100k tracks pipelined over S streams; plus how far apart same-age tracks are (row spread of a
cohort), which decides what stays L2-resident.  Usage: python tools/step_modes.py [K]"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import movmodel as mm

class A: rows, cols, resolution, seed, no_solve, tracks_per_gpu = 5000, 6000, 10.0, 2021, False, 100000
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sr, sc = bench.start_cells(A, 1_000_000)
up, pot, info = bench.build_fields_gpu(A, torch)
f = mm.interleave_fields(up, pot)
shape = (A.rows, A.cols)
lib = mm.N.load()
def ev(): return torch.cuda.Event(enable_timing=True)
tab = mm.build_transition_table(f, 0.0)
# ---- cohort spread: rows of 4000 tracks at fixed ages
m = 4000
lens = mm.simulate_tracks_batch(0.0, sr[:m], sc[:m], shape, fields=f, seed=7).traj_len.cpu().numpy()
cap = 30000
rec = mm.simulate_tracks_batch(0.0, sr[:m], sc[:m], shape, fields=f, seed=7, record=True, traj_cap=cap)
tr = rec.traj.cpu().numpy()          # [cap, m, 2]
for k in (1000, 2000, 4000, 6000, 8000, 10000, 14000, 20000, 29000):
    alive = lens > k
    rows_k = tr[k, alive, 0].astype(np.int64)
    if alive.sum() < 10: break
    q = np.percentile(rows_k, [1, 10, 50, 90, 99])
    print(f"age {k:6d}: alive {alive.mean()*100:5.1f} %  rows p1/p10/p50/p90/p99 = {q.astype(int).tolist()}  band(p1..p99) {int(q[4]-q[0])} rows")
del rec, tr
def kw(mode):
    return dict(gather=dict(), phased=dict(phased=True), walk=dict(walk=True, table=tab))[mode]
def one(n, mode, seed=2021, reps=2):
    best = 1e9
    for _ in range(reps):
        e0, e1 = ev(), ev(); e0.record()
        res = mm.simulate_tracks_batch(0.0, sr[:n], sc[:n], shape, fields=f, seed=seed, **kw(mode))
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    L = res.traj_len.cpu().numpy().astype(np.int64) - 1
    print(f"n={n:8d} {mode:7s} ms={best:9.3f} steps={L.sum():12d} maxlen={L.max():7d} us/step(longest)={best*1e3/L.max():6.3f} steps/s={L.sum()/best*1e3:.3e}", flush=True)
for mode in ("gather", "phased", "walk"):
    for n in (1024, 100_000, 1_000_000):
        one(n, mode, reps=1 if n == 1_000_000 else 2)
n = 100_000
wsb = int(lib.ssrs_walk_workspace_bytes(n))
for mode in ("gather", "phased", "walk"):
    for S in (1, 2, 3, 4, 6, 8):
        streams = [torch.cuda.Stream() for _ in range(S)]
        pres = [torch.zeros(shape, dtype=torch.int32, device="cuda") for _ in range(S)]
        ws = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(S)]
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for s in streams: s.wait_event(e0)
        for i in range(K):
            s = streams[i % S]
            with torch.cuda.stream(s):
                pres[i % S].zero_()
                extra = dict(kw(mode))
                if mode != "gather": extra["workspace"] = ws[i % S]
                mm.simulate_tracks_batch(0.0, sr[:n], sc[:n], shape, fields=f, seed=3000 + i, presence=pres[i % S],
                                         total_steps=total, **extra)
        for s in streams: torch.cuda.current_stream().wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"pipelined {mode:7s} K={K} S={S}: {ms:.1f} ms total, {ms/K:.2f} ms/batch, {int(total.item())/ms*1e3:.3e} steps/s", flush=True)
