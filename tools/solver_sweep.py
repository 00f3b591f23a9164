"""Sweeps the solver's experimental knobs (env SSRS_X_*) on the bench grid; prints iterations and times."""
import os, sys, time, itertools, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssrs_b200 import layers
from ssrs_b200.potential import solve_potential_device
from ssrs_b200.synth import synthetic_dem
rows, cols, res = 5000, 6000, 10.0
z = torch.from_numpy(synthetic_dem(rows, cols, res)).cuda()
K = layers.updraft_fields(z, res, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
solve_potential_device(K[:512, :512].contiguous(), 0.0, strict=False)
base = None
for spec in sys.argv[1:]:
    for kv in spec.split(","):
        k, v = kv.split("=")
        os.environ[k] = v
    torch.cuda.synchronize(); t0 = time.time()
    phi, st = solve_potential_device(K, 0.0, strict=False)
    torch.cuda.synchronize(); dt = time.time() - t0
    if base is None:
        base = phi.clone()
    d = (phi - base).abs().max().item()
    print(spec, "wall_s %.3f it %d restarts %d conv %d rel %.2e setup %.0f solve %.0f maxdiff_vs_first %.3g" % (
        dt, st["iterations"], st["restarts"], st["converged"], st["rel_residual"], st["setup_ms"], st["solve_ms"], d), flush=True)
