"""Track-length statistics against grid size at 10 m resolution: is the 2 x nrow mean at 5000 x 6000
physics of the float32 potential or solver error?  For 1000 x 1200 (GPU potential AND the refined-truth potential of
tests/golden/potential_truth10m.npz), 2000 x 2400 and 5000 x 6000: 4000 northbound tracks x 2 seeds, lengths in units
of nrow, plus the float32 plateau fraction (cells without a strictly lower neighbour) of each potential."""
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import layers, movmodel as mm
from ssrs_b200.potential import solve_potential_device
from ssrs_b200.synth import synthetic_dem

def plateau(p):
    rows, cols = p.shape
    c = p[1:-1, 1:-1]
    has_lower = torch.zeros_like(c, dtype=torch.bool)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr or dc:
                has_lower |= p[1 + dr:rows - 1 + dr, 1 + dc:cols - 1 + dc] < c
    return float((~has_lower).float().mean().item())

def lengths(up, pot, rows, cols, res, label):
    w = {"rows": rows, "cols": cols, "res": res}
    f = mm.interleave_fields(up, pot)
    out = {"grid": [rows, cols], "potential": label, "plateau_fraction_f32": plateau(pot)}
    for seed in (1, 2):
        sr, sc = bench.start_cells(w, 4000, seed)
        L = mm.simulate_tracks_batch(0.0, sr, sc, (rows, cols), fields=f, seed=seed).traj_len.cpu().numpy().astype(np.int64) - 1
        out[f"seed{seed}"] = {"mean_over_nrow": float(L.mean() / rows), "p50": float(np.median(L) / rows),
                             "p99": float(np.percentile(L, 99) / rows), "max": float(L.max() / rows)}
    print(json.dumps(out), flush=True)

for rows, cols in ((1000, 1200), (2000, 2400), (5000, 6000)):
    res = 10.0
    z = torch.from_numpy(synthetic_dem(rows, cols, res)).cuda()
    up = layers.updraft_fields(z, res, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
    pot, st = solve_potential_device(up, 0.0)
    lengths(up, pot, rows, cols, res, f"ssrs_b200 GPU solve ({st['iterations']} iterations)")
    if rows == 1000:
        g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "potential_truth10m.npz"))
        K = torch.from_numpy(g["K32"]).cuda()
        print(json.dumps({"K_equals_fixture": bool(torch.equal(K, up))}))
        lengths(K, torch.from_numpy(g["phi_truth32"]).cuda(), rows, cols, res, "refined truth (SuperLU + long-double refinement), float32")
        ulp = np.spacing(np.abs(g["phi_truth32"])).astype(np.float64)
        slu = (g["phi_truth32"].astype(np.float64) + g["superlu_minus_truth_ulp"] * ulp).astype(np.float32)
        lengths(K, torch.from_numpy(slu).cuda(), rows, cols, res, "reference answer: unrefined SuperLU, float32")
