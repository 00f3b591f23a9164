"""Seasonal-mode run on N GPUs (under torch.distributed.run): `cases` wind conditions on a rows x cols grid, once with
every case sharded over the ranks (tracks block-partitioned, row-sharded potential solve, presence all-reduce) and once
with the cases distributed over the ranks (case_parallel: no solver/presence communication).  Prints both wall times
and compares them: the potentials agree to float32 rounding (the row-sharded solve uses a different hierarchy), so the
stochastic tracks are different realisations of the same process and the presence maps agree statistically.
usage: seasonal_run.py rows cols res cases tracks"""
import json, os, sys, time, tempfile, shutil
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssrs_b200 import Config, Simulator, dist as D
from ssrs_b200.synth import seasonal_wind_conditions, synthetic_dem

rows, cols, res, ncases, ntracks = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = synthetic_dem(rows, cols, res)
spd, drn = seasonal_wind_conditions(ncases, seed=11)
cases = {f"case{i:02d}": (float(spd[i]), float(drn[i])) for i in range(ncases)}
km = (cols * res / 1000.0, rows * res / 1000.0)
box = [tempfile.mkdtemp(prefix="ssrs_seasonal_") if rank == 0 else None]
if world > 1:
    dist.broadcast_object_list(box, src=0)
out_dir = box[0]
result = {"grid": [rows, cols], "world": world, "cases": ncases, "tracks_per_case": ntracks}
maps = {}
if world > 1:
    D.native_comm()                          # one-time NCCL communicator creation stays outside the timed runs
    D.presence_allreduce(torch.zeros(8, dtype=torch.int32, device="cuda"))
sys.stdout = open(os.devnull, "w")          # the Simulator prints like the reference does
for mode in ("sharded", "case_parallel"):
    cfg = Config(run_name=f"seas_{mode}", out_dir=out_dir, sim_seed=5, sim_mode="seasonal", region_width_km=km, resolution=res,
                 track_count=ntracks, track_start_region=(km[0] * 5 / 60, km[0] * 55 / 60, km[1] / 50, km[1] * 2 / 50),
                 track_direction=0.)
    D.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    sim = Simulator(cfg, elevation=z, wind_cases=cases, case_parallel=(mode == "case_parallel"))
    sim.simulate_tracks(save_tracks=False)
    summ = sim.compute_presence_map(radius=200.0 if res <= 10 else 1000.0)
    D.barrier(); torch.cuda.synchronize()
    result[mode] = {"wall_s": time.perf_counter() - t0, "track_steps_last_case": sim.total_track_steps}
    maps[mode] = (summ, sim.mode_data_dir)
sys.stdout = sys.__stdout__
if rank == 0:
    ulp = float(np.spacing(np.float32(1000.0)))
    worst = 0.0
    for cid in cases:
        a = np.load(os.path.join(maps["sharded"][1], f"{cid}_d0_t75_fluidflow_r0_potential.npy")).astype(np.float64)
        b = np.load(os.path.join(maps["case_parallel"][1], f"{cid}_d0_t75_fluidflow_r0_potential.npy")).astype(np.float64)
        worst = max(worst, float(np.abs(a - b).max()) / ulp)
    p, q = maps["sharded"][0].astype(np.float64), maps["case_parallel"][0].astype(np.float64)
    result["potential_max_diff_ulp"] = worst
    result["summary_l1"] = float(np.abs(p / p.sum() - q / q.sum()).sum() / 2)      # normalised-L1 distance
if rank == 0:
    print(json.dumps(result), flush=True)
    shutil.rmtree(out_dir, ignore_errors=True)
if world > 1:
    D.destroy_native_comm()
    dist.destroy_process_group()
