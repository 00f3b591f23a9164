"""BASELINE configs[0] (600 x 500 DEM at 100 m, wind 10 m/s from 270 deg, 1000 northbound tracks) on the same box, both
ways: the UNMODIFIED reference on the host cores — stencil, np.vectorize threshold, assembly, SuperLU
solve, pooled stepping (oracle/ref_cpu.config1_full; test infrastructure, staged under oracle/_ref/) — and the drop-in
`Simulator` on the GPU (second of two runs, so CUDA context, workspace arena and module loading are paid by the first).
Prints one JSON line with seconds per stage and track-steps/s for both.  usage: config1_compare.py [--cpu-only]"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_cpu
from ssrs_b200.movmodel import get_starting_indices
from ssrs_b200.synth import synthetic_dem

rows, cols, res = 500, 600, 100.0
z = synthetic_dem(rows, cols, res)
np.random.seed(7)
sr, sc = get_starting_indices(1000, (5, 55, 1, 2), "random", (60., 50.), res)
out = {"config": "BASELINE configs[0]: uniform mode, 600x500 DEM at 100 m, wind 10 m/s from 270 deg, 1000 northbound tracks",
       "host_cores": os.cpu_count()}
t0 = time.perf_counter()
ref = ref_cpu.config1_full(z, res, 10.0, 270.0, 0.75, sr, sc, 0.0)
ref["total_s"] = time.perf_counter() - t0
out["reference_cpu"] = ref
if "--cpu-only" not in sys.argv:
    import torch
    from ssrs_b200 import Config, Simulator
    real_stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")          # the Simulator prints like the reference does
    runs = []
    for k in range(2):
        with tempfile.TemporaryDirectory(prefix="ssrs_cfg1_") as d:
            cfg = Config(run_name=f"cfg1_{k}", out_dir=d, sim_seed=7, region_width_km=(60., 50.), resolution=res,
                         sim_mode="uniform", uniform_windspeed=10., uniform_winddirn=270., track_direction=0.,
                         track_count=1000, track_start_region=(5, 55, 1, 2))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            sim = Simulator(cfg, elevation=z)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            sim.simulate_tracks()                # potential solve + stepping + presence + trajectory recording + artefacts
            torch.cuda.synchronize(); t2 = time.perf_counter()
            sim.compute_presence_map(radius=1000.)
            torch.cuda.synchronize(); t3 = time.perf_counter()
            runs.append({"constructor_s": t1 - t0, "simulate_tracks_s": t2 - t1, "presence_map_s": t3 - t2, "total_s": t3 - t0,
                         "potential_s": sim.timings.get("potential_s"), "stepping_launches_s": sim.timings.get("tracks_s"),
                         "solve_iterations": sim.solve_stats["iterations"], "track_steps": sim.total_track_steps,
                         "track_steps_per_s_whole_call": sim.total_track_steps / (t2 - t1)})
    sys.stdout = real_stdout
    out["ssrs_b200_gpu_first_run"], out["ssrs_b200_gpu"] = runs
    out["speedup_total"] = ref["total_s"] / runs[1]["total_s"]
print(json.dumps(out))
