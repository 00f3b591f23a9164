"""TEST INFRASTRUCTURE ONLY — numpy restatement of the SSRS hot path (the parity oracle).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module;
the product (`ssrs_b200/`) never does.  Every function cites the reference lines it restates
(paths relative to `/root/reference/`).  Pinning: `oracle/make_golden.py` runs the *unmodified*
reference (loaded by `oracle/ref_loader.py`) in the authoring container and stores its outputs under
`tests/golden/`; `tests/test_oracle_golden.py` checks this restatement against those files, so the
oracle is pinned by outputs of the reference itself (the reference ships no tests of its own).

Conventions: arrays are `[row, col]`, row index grows northward; the flat 3x3 index is
`3*(dr+1)+(dc+1)` (`ssrs/movmodel.py:131-141`).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SQRT2_F32 = float(np.float32(math.sqrt(2.0)))       # facs are stored as 'f4', movmodel.py:82
HM_FLOOR = 1e-08                                    # movmodel.py:104-105


# --------------------------------------------------------------------------------------------
# stage 1: slope / aspect / orographic updraft / threshold            ssrs/layers.py
# --------------------------------------------------------------------------------------------
def horn_gradients(z: np.ndarray, res: float):
    """`dz_dx` (derivative along axis 0) and `dz_dy` (along axis 1) on the interior.
    layers.py:80-90 (same expressions again at :113-123)."""
    z1 = z[:-2, 2:]; z2 = z[1:-1, 2:]; z3 = z[2:, 2:]
    z4 = z[:-2, 1:-1]; z6 = z[2:, 1:-1]
    z7 = z[:-2, :-2]; z8 = z[1:-1, :-2]; z9 = z[2:, :-2]
    dz_dx = ((z3 + 2 * z6 + z9) - (z1 + 2 * z4 + z7)) / (8 * res)
    dz_dy = ((z1 + 2 * z2 + z3) - (z7 + 2 * z8 + z9)) / (8 * res)
    return dz_dx, dz_dy


def compute_slope_degrees(z: np.ndarray, res: float) -> np.ndarray:
    """layers.py:63-93 — border cells are NaN -> nan_to_num -> 0."""
    out = np.zeros_like(z, dtype=np.float64)
    gx, gy = horn_gradients(np.asarray(z, dtype=np.float64), res)
    out[1:-1, 1:-1] = np.degrees(np.arctan(np.sqrt(gx ** 2 + gy ** 2)))
    return out


def compute_aspect_degrees(z: np.ndarray, res: float) -> np.ndarray:
    """layers.py:96-128 — `dz_dx == 0 -> 1e-10` (:124), `180 - atan(dy/dx) + 90*sign(dx)` (:125-127)."""
    out = np.zeros_like(z, dtype=np.float64)
    gx, gy = horn_gradients(np.asarray(z, dtype=np.float64), res)
    gx = np.where(gx == 0.0, 1e-10, gx)
    out[1:-1, 1:-1] = 180.0 - np.degrees(np.arctan(gy / gx)) + 90.0 * (gx / np.abs(gx))
    return out


def compute_orographic_updraft(wspeed, wdirn, slope, aspect, min_updraft_val: float = 0.0):
    """layers.py:11-22."""
    aspect_diff = np.maximum(0.0, np.cos((aspect - wdirn) * np.pi / 180.0))
    return np.maximum(min_updraft_val, wspeed * (np.sin(slope * np.pi / 180.0) * aspect_diff))


def get_above_threshold_speed(arr: np.ndarray, thr: float) -> np.ndarray:
    """layers.py:171-185.  `np.vectorize` hands each element to the scalar function as a Python float,
    so the arithmetic is float64 whatever the input dtype; output float64 (the usual case: element
    [0,0] is a border zero)."""
    a = np.asarray(arr).astype(np.float64)
    with np.errstate(over="ignore"):
        mid = thr * (np.exp((a / thr) ** 5) - 1.0) / (np.exp(1) - 1.0)
    return np.where(a > 1e-02, np.where(a > thr, a, mid), 0.0)


def updraft_pipeline(z32: np.ndarray, res: float, wspeed, wdirn, thr: float):
    """Glue of simulator.py:189-198 + :230-243: float64 stencil -> orograph saved as float32 ->
    reloaded -> threshold.  Returns (slope, aspect, orograph_f32, updraft_f64)."""
    z = np.asarray(z32, dtype=np.float64)
    slope = compute_slope_degrees(z, res)
    aspect = compute_aspect_degrees(z, res)
    ws = wspeed * np.ones(z.shape) if np.isscalar(wspeed) else np.asarray(wspeed, dtype=np.float64)
    wd = wdirn * np.ones(z.shape) if np.isscalar(wdirn) else np.asarray(wdirn, dtype=np.float64)
    oro = compute_orographic_updraft(ws, wd, slope, aspect).astype(np.float32)
    return slope, aspect, oro, get_above_threshold_speed(oro, thr)


# --------------------------------------------------------------------------------------------
# stage 2: Dirichlet sets, 8-neighbour operator, direct solve          ssrs/movmodel.py:21-128
# --------------------------------------------------------------------------------------------
def boundary_nodes(move_dirn: float, nrow: int, ncol: int):
    """movmodel.py:21-57, in column-major node ids `i = col*nrow + row` exactly as the reference."""
    north = np.array([nrow * (x + 1) - 1 for x in range(ncol)], dtype=np.int64)
    south = np.array([nrow * x for x in range(ncol)], dtype=np.int64)
    west = np.arange(1, nrow - 1, dtype=np.int64)
    east = (ncol - 1) * nrow + np.arange(1, nrow - 1, dtype=np.int64)
    ang = move_dirn % 90.0
    quad = (move_dirn % 360) // 90.0
    cl = round(ncol * ang / 90.0)
    rl = round(nrow * ang / 90.0)
    if quad == 0:
        low = np.concatenate((north[cl:], east[nrow - rl:]))
        high = np.concatenate((south[:ncol - cl], west[:rl]))
    elif quad == 1:
        low = np.concatenate((south[ncol - cl:], east[:nrow - rl]))
        high = np.concatenate((north[:cl], west[rl:]))
    elif quad == 2:
        low = np.concatenate((south[:ncol - cl], west[:rl]))
        high = np.concatenate((north[cl:], east[nrow - rl:]))
    else:
        high = np.concatenate((south[ncol - cl:], east[:nrow - rl]))
        low = np.concatenate((north[:cl], west[rl:]))
    nodes = np.concatenate((low, high))
    vals = np.zeros(nodes.size)
    vals[nodes.size // 2:] = 1000.0           # movmodel.py:54-56 (split by position, not by set)
    return nodes, vals


def boundary_grid(move_dirn: float, nrow: int, ncol: int):
    """Same sets scattered to `[row, col]`: (bool mask, float64 values)."""
    nodes, vals = boundary_nodes(move_dirn, nrow, ncol)
    mask = np.zeros((nrow, ncol), dtype=bool)
    val = np.zeros((nrow, ncol), dtype=np.float64)
    mask[nodes % nrow, nodes // nrow] = True
    val[nodes % nrow, nodes // nrow] = vals
    return mask, val


def harmonic_mean_floor(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """movmodel.py:442-447 with minval=1e-8 (:104-105)."""
    with np.errstate(divide="ignore"):
        hm = 2.0 / (1.0 / a + 1.0 / b)
    return np.where((a != 0) & (b != 0), hm, HM_FLOOR)


def _slices(n: int, d: int):
    """(source slice, neighbour slice) along one axis for offset d in {-1,0,1}."""
    if d == 0:
        return slice(0, n), slice(0, n)
    if d == 1:
        return slice(0, n - 1), slice(1, n)
    return slice(1, n), slice(0, n - 1)


def edge_weights(K: np.ndarray) -> np.ndarray:
    """Un-normalised link weights g[d, r, c] of movmodel.py:59-84 + :100-105,
    d = 3*(dr+1)+(dc+1); zero where the neighbour is outside the grid; g[4] = 0.
    Includes the last-column quirk: interior rows of column ncol-1 get sqrt2 on S and 1 on SW."""
    K = np.asarray(K, dtype=np.float64)
    nrow, ncol = K.shape
    g = np.zeros((9, nrow, ncol))
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr == 0 and dc == 0:
                continue
            rs, rn = _slices(nrow, dr)
            cs, cn = _slices(ncol, dc)
            fac = np.full((nrow, ncol), SQRT2_F32 if dr * dc != 0 else 1.0)
            if nrow > 2:
                if (dr, dc) == (-1, 0):
                    fac[1:nrow - 1, ncol - 1] = SQRT2_F32
                elif (dr, dc) == (-1, -1):
                    fac[1:nrow - 1, ncol - 1] = 1.0
            d = 3 * (dr + 1) + (dc + 1)
            g[d][rs, cs] = harmonic_mean_floor(K[rs, cs], K[rn, cn]) / fac[rs, cs]
    return g


def apply_operator(g: np.ndarray, phi: np.ndarray) -> np.ndarray:
    """y[r,c] = sum_d g[d,r,c] * (phi[r,c] - phi[r+dr,c+dc])  (free-node equation, Appendix B)."""
    nrow, ncol = phi.shape
    y = np.zeros_like(phi, dtype=np.float64)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr == 0 and dc == 0:
                continue
            rs, rn = _slices(nrow, dr)
            cs, cn = _slices(ncol, dc)
            d = 3 * (dr + 1) + (dc + 1)
            y[rs, cs] += g[d][rs, cs] * (phi[rs, cs] - phi[rn, cn])
    return y


def solve_potential(K: np.ndarray, move_dirn: float) -> np.ndarray:
    """movmodel.py:86-128: row-normalised `(I - G_ii) phi_i = G_ib phi_b`, SuperLU, float32 result.
    Built in `[row, col]` (row-major) ids; the permutation to the reference's column-major ids does
    not change the solution of the direct solve beyond rounding."""
    K = np.asarray(K, dtype=np.float64)
    nrow, ncol = K.shape
    n = nrow * ncol
    g = edge_weights(K)
    idx = np.arange(n).reshape(nrow, ncol)
    rows_l, cols_l, vals_l = [], [], []
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr == 0 and dc == 0:
                continue
            rs, rn = _slices(nrow, dr)
            cs, cn = _slices(ncol, dc)
            d = 3 * (dr + 1) + (dc + 1)
            rows_l.append(idx[rs, cs].ravel())
            cols_l.append(idx[rn, cn].ravel())
            vals_l.append(g[d][rs, cs].ravel())
    G = sp.coo_matrix((np.concatenate(vals_l), (np.concatenate(rows_l), np.concatenate(cols_l))),
                      shape=(n, n)).tocsr()
    rowsum = np.asarray(G.sum(axis=1)).ravel()
    G = sp.diags(1.0 / rowsum) @ G                       # movmodel.py:110-112
    bmask, bval = boundary_grid(move_dirn, nrow, ncol)
    bm = bmask.ravel()
    inner = np.flatnonzero(~bm)
    bnd = np.flatnonzero(bm)
    Gi = G[inner, :].tocsc()
    A = sp.eye(inner.size, format="csc") - Gi[:, inner]  # movmodel.py:119-120
    b = Gi[:, bnd] @ bval.ravel()[bnd]
    x = spla.spsolve(A.tocsc(), b)                       # movmodel.py:121
    phi = np.empty(n)
    phi[inner] = x
    phi[bnd] = bval.ravel()[bnd]
    return phi.reshape(nrow, ncol).astype(np.float32)    # movmodel.py:128


# --------------------------------------------------------------------------------------------
# stage 3: track stepping                                             ssrs/movmodel.py:131-318
# --------------------------------------------------------------------------------------------
NEIGHBOUR_DELTAS = [(r - 1, c - 1) for r in range(3) for c in range(3)]          # movmodel.py:131-141
NORMS_INV = np.array([[1 / np.sqrt(2), 1, 1 / np.sqrt(2)], [1, 0, 1], [1 / np.sqrt(2), 1, 1 / np.sqrt(2)]],
                     dtype=np.float32)


def track_restrictions(dr: int, dc: int) -> np.ndarray:
    """movmodel.py:185-202 as a table: moves within 45 deg of the previous move; (0,0) -> all but centre."""
    a = np.zeros((3, 3), dtype=int)
    if dr == 0 and dc == 0:
        a[:, :] = 1
    else:
        for r in (-1, 0, 1):
            for c in (-1, 0, 1):
                if (r, c) == (0, 0):
                    continue
                # angle between (dr,dc) and (r,c) <= 45 deg
                dot = (dr * r + dc * c) / (math.hypot(dr, dc) * math.hypot(r, c))
                if dot > 0.7:
                    a[r + 1, c + 1] = 1
    a[1, 1] = 0
    return a.flatten()


def directional_probs(theta: float) -> np.ndarray:
    """movmodel.py:247-257 (theta in radians, clockwise from north)."""
    m = np.zeros((3, 3))
    m[0, :] = [np.cos(np.pi / 4 + theta), np.cos(theta), np.cos(7 * np.pi / 4 + theta)]
    m[1, :] = [np.cos(np.pi / 2 + theta), 0, np.cos(3 * np.pi / 2 + theta)]
    m[2, :] = [np.cos(3 * np.pi / 4 + theta), np.cos(np.pi + theta), np.cos(5 * np.pi / 4 + theta)]
    m[m < 0.01] = 0.0
    return np.flipud(m.clip(min=0.0)).flatten()


def move_away_from_boundary(row, col, nr, nc):
    """movmodel.py:205-217."""
    nrow_, ncol_ = row, col
    if row <= 1:
        nrow_ = row + 2
    elif row >= nr - 2:
        nrow_ = row - 2
    if col <= 0:
        ncol_ = col + 2
    elif col >= nc - 2:
        ncol_ = col - 2
    return nrow_, ncol_


def pairwise9(p) -> float:
    """numpy's float64 add.reduce over 9 contiguous elements (pairwise-sum kernel, 8 accumulators
    then the tail): ((p0+p1)+(p2+p3)) + ((p4+p5)+(p6+p7)) + p8."""
    return (((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]))) + p[8]


def move_probabilities(p_in, dirvec, nu: float, mask):
    """movmodel.py:220-244 with the directional vector precomputed."""
    p = np.array(p_in, dtype=np.float64)
    if np.isnan(p).any():
        p = dirvec.copy()
    p = np.clip(p, 0.0, None)
    p[4] = 0.0
    p = p * mask
    if np.count_nonzero(p) == 0:
        p = dirvec.copy()
    p[4] = 0.0
    p = p * mask
    if np.count_nonzero(p) == 0:
        p = dirvec.copy()
    p = p / pairwise9(p)
    p = np.power(p, nu)
    p = p / pairwise9(p)
    return p


def simulate_track(move_dirn, start, shape, mem, nu, U, P, uniforms):
    """movmodel.py:264-318 for one track, consuming `uniforms[k]` at step k in place of the global
    numpy stream (`np.random.choice` draws exactly one `random_sample()` per call, :312).
    `U` float64 (thresholded updraft) or None ('drw'), `P` float32 or None.  Returns int16 [L,2]."""
    nr, nc = shape
    burnin = int(min(nr, nc) / 10)
    max_moves = nr / 2 * nc / 2
    dirvec = directional_probs(move_dirn * np.pi / 180.0)
    row0, col0 = int(start[0]), int(start[1])
    traj = [(row0, col0)]
    dirs = [(0, 0)]
    pos = (row0, col0)
    k = 0
    while k < max_moves:
        row, col = pos
        if k > burnin:
            if not (0 < row < nr - 1 and 0 < col < nc - 1):
                break
        else:
            row, col = move_away_from_boundary(row, col, nr, nc)
        probs = np.ones((3, 3), dtype=np.float32)
        if U is not None:
            lu = np.clip(U[row - 1:row + 2, col - 1:col + 2], 1e-06, None)
            probs = probs * (2.0 / (1.0 / lu[1, 1] + 1.0 / lu))
        else:
            probs = dirvec.reshape(3, 3)
        if P is not None:
            lp = P[row - 1:row + 2, col - 1:col + 2]
            probs = probs * ((lp[1, 1] - lp) * NORMS_INV)
        mask = track_restrictions(0, 0)
        for d in dirs[-mem:]:
            mask = mask & track_restrictions(*d)
        p = move_probabilities(probs.flatten(), dirvec, nu, mask)
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        idx = int(np.searchsorted(cdf, uniforms[k], side="right"))
        dr, dc = NEIGHBOUR_DELTAS[idx]
        pos = (row + dr, col + dc)
        traj.append(pos)
        dirs.append((dr, dc))
        k += 1
    return np.array(traj, dtype=np.int16)


def starting_indices(ntracks, sbounds, stype, twidth, tres, rng=None):
    """movmodel.py:144-182; `rng` is a `np.random.RandomState` standing in for the global stream."""
    from math import ceil, floor
    if (sbounds[1] < sbounds[0] or sbounds[3] < sbounds[2] or sbounds[0] < 0.0 or sbounds[2] < 0.0
            or sbounds[1] > twidth[0] or sbounds[3] > twidth[1]):
        raise ValueError("track_start_region incompatible with terrain_width!")
    res_km = tres / 1000.0
    xmax = ceil(twidth[0] / res_km)
    ymax = ceil(twidth[1] / res_km)
    xlo = min(max(floor(sbounds[0] / res_km) - 1, 1), xmax - 2)
    xup = max(min(ceil(sbounds[1] / res_km), xmax - 1), 2)
    ylo = min(max(floor(sbounds[2] / res_km) - 1, 1), ymax - 2)
    yup = max(min(ceil(sbounds[3] / res_km), ymax - 1), 2)
    xm, ym = np.mgrid[xlo:xup, ylo:yup]
    base = np.vstack((np.ravel(ym), np.ravel(xm)))
    nb = base.shape[1]
    if stype == "structured":
        idx = np.round(np.linspace(0, nb - 1, ntracks % nb))
        if ntracks > nb:
            s = np.tile(base, (1, ntracks // nb))
            s = np.hstack((s, s[:, idx.astype(int)]))
        else:
            s = base[:, idx.astype(int)]
    elif stype == "random":
        rng = np.random if rng is None else rng
        s = base[:, rng.randint(0, nb, ntracks)]
    else:
        raise ValueError(f"Model:Invalid sim_start_type of {stype}\nOptions: structured, random")
    s = s.astype(int)
    return s[0, :], s[1, :]


# --------------------------------------------------------------------------------------------
# stage 4: presence                                                   ssrs/movmodel.py:410-439
# --------------------------------------------------------------------------------------------
def presence_counts(tracks, shape) -> np.ndarray:
    """movmodel.py:410-419 with int64 accumulation (the reference's int16 wraps above 32767)."""
    cnt = np.zeros(shape, dtype=np.int64)
    for t in tracks:
        np.add.at(cnt, (t[:, 0].astype(np.int64), t[:, 1].astype(np.int64)), 1)
    return cnt


def smooth_presence(counts: np.ndarray, radius: float) -> np.ndarray:
    """movmodel.py:422-439 (disk kernel, `convolve2d(mode='same')`, float32)."""
    import scipy.signal as ssg
    krad = int(radius)
    kernel = np.zeros((2 * krad + 1, 2 * krad + 1))
    y, x = np.ogrid[-krad:krad + 1, -krad:krad + 1]
    kernel[x ** 2 + y ** 2 <= krad ** 2] = 1
    kernel /= np.sum(kernel)
    return ssg.convolve2d(counts, kernel, mode="same").astype(np.float32)


# ---- "next" rows f-2 / f-3 ------------------------------------------------------------
def interpolated_wind_conditions(xlocs, ylocs, wspeed, wdirn, xgrid, ygrid, method: str = 'linear'):
    """ssrs/simulator.py:765-792 (`_interpolate_wtk_vardata` + `_get_interpolated_wind_conditions`): u/v components
    through scipy.interpolate.griddata — the third-party arithmetic the reference itself calls — then speed and
    direction.  float64 [len(ygrid), len(xgrid)], NaN outside the convex hull of the sites."""
    from scipy.interpolate import griddata
    points = np.array([xlocs, ylocs]).T                                  # :770
    xmesh, ymesh = np.meshgrid(xgrid, ygrid)                             # :771
    easterly = np.multiply(wspeed, np.sin(np.asarray(wdirn) * np.pi / 180.))     # :784
    northerly = np.multiply(wspeed, np.cos(np.asarray(wdirn) * np.pi / 180.))    # :785
    ie = griddata(points, easterly, (xmesh, ymesh), method=method)       # :772-773
    inn = griddata(points, northerly, (xmesh, ymesh), method=method)
    ws = np.sqrt(np.square(ie) + np.square(inn))                         # :788-789
    wd = np.arctan2(ie, inn)                                             # :790
    wd = np.mod(wd + 2. * np.pi, 2. * np.pi)                             # :791
    return ws, wd * 180. / np.pi


def thermal_hit_probability(aspect: np.ndarray) -> np.ndarray:
    """ssrs/layers.py:194-202: probability that a cell seeds a thermal: inside the 10 % border,
    np.random.randint(1, int(wtfactor)) == 5 with wtfactor = 1000 + |aspect - 180| / 180 * 2000."""
    ysize, xsize = aspect.shape
    bx, by = int(0.1 * xsize), int(0.1 * ysize)
    p = np.zeros(aspect.shape)
    wt = (1000 + (np.abs(aspect - 180.) / 180.) * 2000.).astype(np.int64)        # int(wtfactor)
    p[by:ysize - by, bx:xsize - bx] = 1.0 / (wt[by:ysize - by, bx:xsize - bx] - 1)
    return p


def smooth_thermals(wt_init: np.ndarray) -> np.ndarray:
    """ssrs/layers.py:211: scipy.ndimage.gaussian_filter(wt_init, sigma=4, mode='constant')."""
    from scipy import ndimage
    return ndimage.gaussian_filter(wt_init, sigma=4, mode='constant')
