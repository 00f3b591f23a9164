"""Stopping rule vs accuracy: error of the GPU potential against the refined truth at 1000 x 1200 / 10 m (float32 ulp at 1000)
and solve time at 5000 x 6000, for several settings of the two experiment knobs of the stopping rule."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssrs_b200 import layers
from ssrs_b200.potential import solve_potential_device
from ssrs_b200.synth import synthetic_dem
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "potential_truth10m.npz"))
K1 = torch.from_numpy(g["K32"]).cuda(); truth = g["phi_truth32"].astype(np.float64)
ULP = float(np.spacing(np.float32(1000.0)))
z = torch.from_numpy(synthetic_dem(5000, 6000, 10.0)).cuda()
K5 = layers.updraft_fields(z, 10.0, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
solve_potential_device(K5, 0.0)
for ff, ac in ((0.5, 0.5), (0.25, 0.5), (0.25, 0.25), (0.125, 0.25), (0.125, 0.125), (0.0625, 0.0625), (0.03, 0.03)):
    os.environ["SSRS_X_FLOORFRAC"] = str(ff); os.environ["SSRS_X_ACCEPT"] = str(ac)
    phi, st = solve_potential_device(K1, 0.0, strict=False)
    d = np.abs(phi.cpu().numpy().astype(np.float64) - truth)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, st5 = solve_potential_device(K5, 0.0, strict=False)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    print(f"floorfrac {ff:6.3f} accept {ac:6.3f}: 1000x1200 err {d.max()/ULP:5.2f} ulp, {100*(d>0).mean():5.2f} % cells differ, it {st['iterations']} conv {st['converged']} res {st['rel_residual']:.2e} | "
          f"5000x6000 {ms:6.1f} ms it {st5['iterations']} conv {st5['converged']} res {st5['rel_residual']:.2e} restarts {st5['restarts']}", flush=True)
