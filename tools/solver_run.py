"""Runs the stage-2 solver once on the bench grid (after a small warm-up solve) and prints its stats."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssrs_b200 import layers
from ssrs_b200.potential import solve_potential_device
from ssrs_b200.synth import synthetic_dem
rows, cols, res = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (5000, 6000, 10.0)
zs = torch.from_numpy(synthetic_dem(256, 320, res, seed=1)).cuda()
ks = layers.updraft_fields(zs, res, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
if not os.environ.get('SSRS_NO_WARMUP'):
    solve_potential_device(ks, 0.0, strict=False)
z = torch.from_numpy(synthetic_dem(rows, cols, res)).cuda()
K = layers.updraft_fields(z, res, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
torch.cuda.synchronize(); t0 = time.time()
phi, st = solve_potential_device(K, 0.0, strict=False)
torch.cuda.synchronize(); print("wall_s", time.time() - t0, st)
