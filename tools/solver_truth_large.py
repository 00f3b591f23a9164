"""Stopping rule vs accuracy at BASELINE's two large grids: the float32 potential of the default rule (floor/16) against
self-truths — the same solver driven far below its default stopping point (explicit rtol, float64 iterate) — in float32
ulp at 1000.  No refined reference exists at these sizes (SuperLU cannot factor 3e7 unknowns here)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssrs_b200 import layers
from ssrs_b200.potential import solve_potential_device
from ssrs_b200.synth import synthetic_dem
ULP = float(np.spacing(np.float32(1000.0)))
for rows, cols in ((5000, 6000), (10000, 12000)):
    z = torch.from_numpy(synthetic_dem(rows, cols, 10.0)).cuda()
    K = layers.updraft_fields(z, 10.0, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
    del z
    solve_potential_device(K, 0.0)
    truths = {}
    for rt in (1e-10, 1e-12):
        os.environ["SSRS_X_FLOORFRAC"] = os.environ["SSRS_X_ACCEPT"] = "1e-9"      # the rule's floor out of the way: rtol decides
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _, st = solve_potential_device(K, 0.0, rtol=rt, max_iter=400, strict=False, want_f64=True)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        truths[rt] = st.pop("potential_f64")
        print(f"{rows}x{cols} truth rtol {rt:g}: {ms:7.1f} ms it {st['iterations']} conv {st['converged']} res {st['rel_residual']:.2e} restarts {st['restarts']}", flush=True)
    d = (truths[1e-10] - truths[1e-12]).abs().max().item()
    print(f"{rows}x{cols} truths differ by {d / ULP:.4f} ulp (float64 iterates)", flush=True)
    T = truths[1e-12]; T32 = T.float()
    del truths
    for ff in (0.5, 0.0625, 1 / 64, 1 / 256, 1 / 1024):
        os.environ["SSRS_X_FLOORFRAC"] = str(ff); os.environ["SSRS_X_ACCEPT"] = str(ff)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        phi, st = solve_potential_device(K, 0.0, strict=False)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        e = (phi.double() - T).abs().max().item() / ULP
        nd = (phi != T32).float().mean().item()
        print(f"{rows}x{cols} floor/{1 / ff:6.0f}: {ms:7.1f} ms it {st['iterations']} conv {st['converged']} res {st['rel_residual']:.2e} restarts {st['restarts']} | "
              f"err {e:6.2f} ulp vs float64 truth, {100 * nd:6.3f} % cells differ from round(truth)", flush=True)
    del os.environ["SSRS_X_FLOORFRAC"], os.environ["SSRS_X_ACCEPT"]
    del K, T, T32, phi
    torch.cuda.empty_cache()
