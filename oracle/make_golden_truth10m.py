"""TEST INFRASTRUCTURE ONLY — writes tests/golden/potential_truth10m.npz: a refined-truth potential at 10 m resolution.

Run in the authoring container (needs /root/reference):  python -m oracle.make_golden_truth10m

Every large BASELINE config runs at 10 m; there the reference's own answer — SuperLU on the row-normalised system,
rounded to float32 (`ssrs/movmodel.py:110-128`) — is itself ~11 float32 ulp away from the exact solution of its
linear system (the system is that ill-conditioned: conductances span 1e-10..1).  A solver can therefore not be
judged against unrefined SuperLU at this resolution.  This script produces the truth instead:
  1. stage 1 by the UNMODIFIED reference (`layers.py`) on the 1000 x 1200 synthetic DEM at 10 m -> K (float32);
  2. the reference's linear system restated in [row, col] ids (oracle_np.edge_weights; equal to
     the reference's matrix entry by entry, tests/test_oracle_golden.py), factorised once by SuperLU;
  3. iterative refinement: residual of the UN-normalised free-node equations in 80-bit long double, correction by the
     same LU factors, until the long-double residual stops falling;
  4. stored: K (float32), the refined potential rounded to float32, and the unrefined SuperLU float32 potential as
     its difference from the truth in float32 ulps (int8).
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as O  # noqa: E402
from oracle.make_golden import reference_fields  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402
from ssrs_b200.synth import synthetic_dem  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "potential_truth10m.npz")
ROWS, COLS, RES = 1000, 1200, 10.0


def apply_operator_ld(g_ld, phi_ld):
    """oracle_np.apply_operator in long double."""
    nrow, ncol = phi_ld.shape
    y = np.zeros_like(phi_ld)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr == 0 and dc == 0:
                continue
            rs, rn = O._slices(nrow, dr)
            cs, cn = O._slices(ncol, dc)
            d = 3 * (dr + 1) + (dc + 1)
            y[rs, cs] += g_ld[d][rs, cs] * (phi_ld[rs, cs] - phi_ld[rn, cn])
    return y


def main():
    L, M = load_reference()
    z = synthetic_dem(ROWS, COLS, RES)
    _, _, _, K, _ = reference_fields(L, M, z, RES, 10.0, 270.0, 0.75, None)
    K32 = np.asarray(K, dtype=np.float32)
    K64 = K32.astype(np.float64)
    nrow, ncol = K64.shape
    n = nrow * ncol
    t0 = time.time()
    g = O.edge_weights(K64)
    idx = np.arange(n).reshape(nrow, ncol)
    rows_l, cols_l, vals_l = [], [], []
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr == 0 and dc == 0:
                continue
            rs, rn = O._slices(nrow, dr)
            cs, cn = O._slices(ncol, dc)
            d = 3 * (dr + 1) + (dc + 1)
            rows_l.append(idx[rs, cs].ravel()); cols_l.append(idx[rn, cn].ravel()); vals_l.append(g[d][rs, cs].ravel())
    G = sp.coo_matrix((np.concatenate(vals_l), (np.concatenate(rows_l), np.concatenate(cols_l))), shape=(n, n)).tocsr()
    rowsum = np.asarray(G.sum(axis=1)).ravel()
    Gn = sp.diags(1.0 / rowsum) @ G                                  # movmodel.py:110-112
    bmask, bval = O.boundary_grid(0.0, nrow, ncol)
    bm = bmask.ravel()
    inner, bnd = np.flatnonzero(~bm), np.flatnonzero(bm)
    Gi = Gn[inner, :].tocsc()
    A = (sp.eye(inner.size, format="csc") - Gi[:, inner]).tocsc()   # movmodel.py:119-120
    b = Gi[:, bnd] @ bval.ravel()[bnd]
    lu = spla.splu(A)
    x = lu.solve(b)                                                  # movmodel.py:121 (spsolve = this factorisation)
    print(f"SuperLU factorisation + solve: {time.time() - t0:.1f} s")
    phi = np.empty(n); phi[inner] = x; phi[bnd] = bval.ravel()[bnd]
    phi_superlu32 = phi.reshape(nrow, ncol).astype(np.float32)       # movmodel.py:128: the reference's answer
    # refinement on the un-normalised equations  sum_j g_ij (phi_i - phi_j) = 0  (free nodes), long-double residuals
    g_ld = g.astype(np.longdouble)
    phi_ld = phi.reshape(nrow, ncol).astype(np.longdouble)
    rs_ld = rowsum.reshape(nrow, ncol).astype(np.longdouble)
    hist = []
    for it in range(6):
        r = -apply_operator_ld(g_ld, phi_ld)
        r[bmask] = 0
        rn = r / rs_ld                                               # residual of the row-normalised rows = A's rows
        hist.append(float(np.abs(rn[~bmask]).max()))
        d = lu.solve(np.asarray(rn.ravel()[inner], dtype=np.float64))
        upd = np.zeros(n); upd[inner] = d
        phi_ld = phi_ld + upd.reshape(nrow, ncol).astype(np.longdouble)
        print(f"refinement {it}: max normalised residual {hist[-1]:.3e}, max correction {np.abs(d).max():.3e}")
        if np.abs(d).max() < 1e-9:
            break
    truth32 = np.asarray(phi_ld, dtype=np.float64).astype(np.float32)
    ulp = np.spacing(np.abs(truth32).astype(np.float32)).astype(np.float64)
    dev = np.abs(phi_superlu32.astype(np.float64) - np.asarray(phi_ld, dtype=np.float64)) / np.maximum(ulp, 1e-30)
    print(f"unrefined SuperLU float32 vs truth: max {dev.max():.2f} ulp, {100 * (dev > 0.5).mean():.1f} % of cells off by > 0.5 ulp")
    # the reference's own (unrefined SuperLU, float32) answer is kept as its distance from the truth in float32 ulps
    delta = np.rint((phi_superlu32.astype(np.float64) - truth32) / ulp).astype(np.int8)
    np.savez_compressed(OUT, K32=K32, phi_truth32=truth32, superlu_minus_truth_ulp=delta, residual_history=np.array(hist),
                        shape=np.array([ROWS, COLS]), res=RES)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
