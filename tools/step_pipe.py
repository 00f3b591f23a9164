"""K batches of 100k tracks pipelined over S streams, phased gather stepping: sweep of S (and whatever the environment
selects: SSRS_B200_LIB, SSRS_X_PHASE_PCT).  Usage: python tools/step_pipe.py K S1,S2,... [mode]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import movmodel as mm
class A: rows, cols, resolution, seed, no_solve, tracks_per_gpu = 5000, 6000, 10.0, 2021, False, 100000
K = int(sys.argv[1]); SS = [int(x) for x in sys.argv[2].split(",")]; mode = sys.argv[3] if len(sys.argv) > 3 else "phased"
sr, sc = bench.start_cells(A, 1_000_000)
up, pot, info = bench.build_fields_gpu(A, torch)
f = mm.interleave_fields(up, pot)
shape = (A.rows, A.cols); n = 100_000
wsb = int(mm.N.load().ssrs_walk_workspace_bytes(1_000_000))
def ev(): return torch.cuda.Event(enable_timing=True)
tag = f"lib={os.path.basename(os.environ.get('SSRS_B200_LIB', 'default'))} pct={os.environ.get('SSRS_X_PHASE_PCT', '25')}"
for nn in (100_000, 1_000_000):
    e0, e1 = ev(), ev(); e0.record()
    res = mm.simulate_tracks_batch(0.0, sr[:nn], sc[:nn], shape, fields=f, seed=2021, phased=(mode == "phased"))
    e1.record(); torch.cuda.synchronize()
    print(f"{tag} single n={nn}: {e0.elapsed_time(e1):.2f} ms  {res.total_steps/e0.elapsed_time(e1)*1e3:.3e} steps/s", flush=True)
for S in SS:
    streams = [torch.cuda.Stream() for _ in range(S)]
    pres = [torch.zeros(shape, dtype=torch.int32, device="cuda") for _ in range(S)]
    ws = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(S)]
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    for rep in range(2):
        total.zero_(); torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for s in streams: s.wait_event(e0)
        for i in range(K):
            with torch.cuda.stream(streams[i % S]):
                pres[i % S].zero_()
                mm.simulate_tracks_batch(0.0, sr[:n], sc[:n], shape, fields=f, seed=3000 + i, presence=pres[i % S], total_steps=total,
                                         phased=(mode == "phased"), workspace=ws[i % S] if mode == "phased" else None)
        for s in streams: torch.cuda.current_stream().wait_stream(s)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{tag} pipelined {mode} K={K} S={S}: {ms:.1f} ms total, {ms/K:.2f} ms/batch, {int(total.item())/ms*1e3:.3e} steps/s", flush=True)
