"""Row-sharded potential solve on N GPUs (run under torch.distributed.run): checks it against the single-GPU solve
of the same grid and prints both times.  usage: sharded_solve_run.py rows cols res [repeat]"""
import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssrs_b200 import layers, dist as D
from ssrs_b200.potential import solve_potential_device
from ssrs_b200.synth import synthetic_dem

rows, cols, res = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
repeat = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = torch.from_numpy(synthetic_dem(rows, cols, res)).cuda()
K = layers.updraft_fields(z, res, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
out = {"grid": [rows, cols], "world": world}
for mode in ("single", "sharded"):
    if mode == "sharded" and world == 1:
        continue
    best = None
    for it in range(repeat):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        phi, st = solve_potential_device(K, 0.0, strict=False, sharded=(mode == "sharded"))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, st)
    out[mode] = {"wall_ms": best[0] * 1e3, "setup_ms": best[1]["setup_ms"], "solve_ms": best[1]["solve_ms"],
                 "iterations": best[1]["iterations"], "restarts": best[1]["restarts"], "converged": best[1]["converged"],
                 "rel_residual": best[1]["rel_residual"], "levels": best[1]["levels"]}
    if mode == "single":
        ref = phi.clone()
    else:
        d = (phi - ref).abs().max().item()
        out["max_abs_diff_vs_single"] = d
        out["ulp_1000"] = float(np.spacing(np.float32(1000.0)))
        # every rank must hold the same raster
        chk = phi.double().sum()
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        out["identical_on_all_ranks"] = bool(all(float(x) == float(lst[0]) for x in lst))
if world > 1:
    # presence all-reduce through the library's communicator
    pres = torch.full((rows, cols), rank + 1, dtype=torch.int32, device="cuda")
    D.presence_allreduce(pres)
    out["presence_allreduce_ok"] = bool((pres == world * (world + 1) // 2).all().item())
    out["halo_mode"] = D.halo_mode()
    dist.barrier()
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    D.destroy_native_comm()
    dist.destroy_process_group()
