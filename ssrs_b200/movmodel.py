"""Host-side mirror of the reference's `ssrs/movmodel.py`, backed by the CUDA library.

Names, argument meaning and error behaviour follow `/root/reference/ssrs/movmodel.py`:
  MovModel(move_dirn, grid_shape).get_boundary_nodes()          :21-57
  MovModel.assemble_sparse_linear_system()                      :59-84   (API parity only)
  MovModel.solve_sparse_linear_system(K, bnodes, benergy, ...)  :86-128  -> GPU solver (stage 2)
  get_starting_indices(ntracks, sbounds, stype, twidth, tres)   :144-182 (host; consumes np.random like the reference)
  generate_simulated_tracks(...)                                :264-318 -> one track through the GPU stepper
  compute_presence_counts(tracks, gridshape)                    :410-419
and the batched entry point the Simulator really uses, `simulate_tracks_batch`, which replaces the
process pool of `ssrs/simulator.py:360-369` with one kernel launch.
"""
from __future__ import annotations

import ctypes as C
from math import ceil, floor, sqrt
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N

# flat move index 3*(dr+1)+(dc+1)  (reference constants, movmodel.py:131-141)
neighbour_deltas = [np.array([i // 3 - 1, i % 3 - 1]) for i in range(9)]
neighbour_delta_norms_inv = np.array(
    [[0.0 if (r, c) == (1, 1) else 1.0 / np.hypot(r - 1, c - 1) for c in range(3)] for r in range(3)], dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# small host helpers kept for API parity
# ---------------------------------------------------------------------------------------------
# multiples of pi/4 added to the heading for each flat move index (centre has no weight)
_COMPASS_K = (3, 4, 5, 2, None, 6, 1, 0, 7)


def get_directional_probs(theta: float) -> np.ndarray:
    """Cosine weights of the 8 compass moves about heading `theta` (radians, clockwise from north);
    weights below 0.01 are dropped; flat layout with row+1 = north (reference :247-257)."""
    out = np.zeros(9)
    for i, k in enumerate(_COMPASS_K):
        if k is None:
            continue
        v = np.cos(k * np.pi / 4 + theta)
        out[i] = v if v >= 0.01 else 0.0
    return out


def get_track_restrictions(dr: int, dc: int) -> np.ndarray:
    """9 flags: moves within 45 degrees of the previous move (dr, dc); (0,0) allows all but staying (reference :185-202)."""
    out = np.zeros(9, dtype=int)
    for i in range(9):
        r, c = i // 3 - 1, i % 3 - 1
        if (r, c) == (0, 0):
            continue
        if dr == 0 and dc == 0:
            out[i] = 1
        else:
            out[i] = int((dr * r + dc * c) / (np.hypot(dr, dc) * np.hypot(r, c)) > 0.7)
    return out


def move_away_from_boundary(row, col, num_rows, num_cols):
    """Burn-in relocation (reference :205-217); note the row/col asymmetry (row <= 1 but col <= 0)."""
    if row <= 1:
        row += 2
    elif row >= num_rows - 2:
        row -= 2
    if col <= 0:
        col += 2
    elif col >= num_cols - 2:
        col -= 2
    return row, col


def get_harmonic_mean(in_first, in_second):
    return 2.0 / (1.0 / in_first + 1.0 / in_second)


def harmonic_mean(aval: float, bval: float, minval: float = 1e-10) -> float:
    """Reference :442-447."""
    return 2.0 / (1.0 / aval + 1 / bval) if (aval != 0 and bval != 0) else minval


def generate_move_probabilities(in_probs, move_dirn: float, nu_par: float, dir_bool) -> np.ndarray:
    """Reference :220-244 (host version, for API parity and small checks)."""
    dirvec = get_directional_probs(move_dirn * np.pi / 180.0)
    p = np.array(in_probs, dtype=float)
    mask = np.asarray(dir_bool, dtype=float)
    if np.isnan(p).any():
        print('NANs in move probabilities!')
        p = dirvec.copy()
    p = np.clip(p, 0.0, None)
    for _ in range(2):
        p[4] = 0.0
        p = p * mask
        if np.count_nonzero(p) == 0:
            p = dirvec.copy()
    p = p / np.sum(p)
    p = np.power(p, nu_par)
    return p / np.sum(p)


def get_starting_indices(ntracks: int, sbounds, stype: str, twidth, tres: float):
    """Start cells inside `track_start_region` (km); 'random' draws `np.random.randint` from the global
    stream exactly as the reference (:144-182), so seeded runs start from the same cells."""
    x0, x1, y0, y1 = sbounds
    if x1 < x0 or y1 < y0 or x0 < 0.0 or y0 < 0.0 or x1 > twidth[0] or y1 > twidth[1]:
        raise ValueError('track_start_region incompatible with terrain_width!')
    km = tres / 1000.0
    nx, ny = ceil(twidth[0] / km), ceil(twidth[1] / km)
    xs = range(min(max(floor(x0 / km) - 1, 1), nx - 2), max(min(ceil(x1 / km), nx - 1), 2))
    ys = range(min(max(floor(y0 / km) - 1, 1), ny - 2), max(min(ceil(y1 / km), ny - 1), 2))
    # candidate list ordered x-major (the reference's np.mgrid ravel order)
    cand_rows = np.tile(np.asarray(ys), len(xs))
    cand_cols = np.repeat(np.asarray(xs), len(ys))
    nb = cand_rows.size
    if stype == 'structured':
        pick = np.round(np.linspace(0, nb - 1, ntracks % nb)).astype(int)
        if ntracks > nb:
            reps = ntracks // nb
            sel = np.concatenate([np.tile(np.arange(nb), reps), pick])
        else:
            sel = pick
    elif stype == 'random':
        sel = np.random.randint(0, nb, ntracks)
    else:
        raise ValueError((f'Model:Invalid sim_start_type of {stype}\n'
                          'Options: structured, random'))
    return cand_rows[sel].astype(int), cand_cols[sel].astype(int)


# ---------------------------------------------------------------------------------------------
# stage 3+4 on the GPU
# ---------------------------------------------------------------------------------------------
class TrackBatchResult:
    """Device-resident outcome of one batched stepping launch."""

    def __init__(self, n_tracks, shape, traj, traj_len, presence, total_steps, traj_cap):
        self.n_tracks = n_tracks
        self.shape = shape
        self.traj = traj              # int16 [traj_cap, n_tracks, 2] step-major, or None
        self.traj_len = traj_len      # int32 [n_tracks]
        self.presence = presence      # int32 [rows, cols] (bit pattern of the kernel's uint32 counts) or None
        self._total = total_steps     # int64 [1] tensor
        self.traj_cap = traj_cap

    @property
    def total_steps(self) -> int:
        return int(self._total.item())

    def _checked_lengths(self):
        if self.traj is None:
            raise ValueError("trajectories were not recorded (record=False)")
        lens = self.traj_len.cpu().numpy()
        if (lens < 0).any():
            raise ValueError("the supplied uniforms were shorter than the longest track (negative traj_len)")
        if (lens > self.traj_cap).any():
            raise ValueError("traj_cap was smaller than the longest track; leave traj_cap=None (two-pass recording) "
                             "or use record_tracks_packed")
        return lens

    def tracks(self) -> List[np.ndarray]:
        """List of int16 [L, 2] arrays like the reference's pool.map result (:318)."""
        lens = self._checked_lengths()
        tr = self.traj.permute(1, 0, 2).contiguous().cpu().numpy()
        return [tr[i, :lens[i]].copy() for i in range(self.n_tracks)]

    def packed(self):
        """(offsets int64 [n + 1], points int16 [total, 2]) on the host — the packed on-disk form (trackio.py),
        gathered on the device from the step-major trajectory buffer (`ssrs_pack_trajectories`)."""
        torch = N.require_cuda()
        lens_h = self._checked_lengths().astype(np.int64)
        offsets_h = np.zeros(self.n_tracks + 1, dtype=np.int64)
        np.cumsum(lens_h, out=offsets_h[1:])
        total = int(offsets_h[-1])
        if total == 0:
            return offsets_h, np.zeros((0, 2), dtype=np.int16)
        offsets = torch.from_numpy(offsets_h[:-1].copy()).to("cuda")
        points = torch.empty((total, 2), dtype=torch.int16, device="cuda")
        N.check(N.load().ssrs_pack_trajectories(N.ptr(self.traj), self.traj_cap, N.ptr(self.traj_len), N.ptr(offsets),
                                                self.n_tracks, N.ptr(points), N.current_stream()), "ssrs_pack_trajectories")
        return offsets_h, points.cpu().numpy()


def interleave_fields(updraft, potential):
    """{updraft, potential} -> float32 [rows, cols, 2] on the device (one 8-byte gather per cell)."""
    torch = N.require_cuda()
    u = _f32_cuda(updraft, torch)
    p = _f32_cuda(potential, torch)
    if u.shape != p.shape or u.dim() != 2:
        raise ValueError("updraft and potential must be 2-D rasters of the same shape")
    out = torch.empty(u.shape + (2,), dtype=torch.float32, device="cuda")
    N.check(N.load().ssrs_interleave_fields(N.ptr(u), N.ptr(p), N.ptr(out), u.numel(), N.current_stream()),
            "ssrs_interleave_fields")
    return out


def _f32_cuda(a, torch):
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to("cuda")


TRAJ_BUFFER_BYTES = 2 << 30        # device budget of one step-major trajectory buffer (4 B x cap x tracks)


def _launch_steps(fields, rows, cols, start, n, track_id0, dirp_c, memory, nu, seed, u_t, ustride, traj, cap, traj_len,
                  presence, total_steps, exact, workspace=None, first_phase_steps=0):
    lib = N.load()
    args = (N.ptr(fields), rows, cols, N.ptr(start), n, int(track_id0), dirp_c, int(memory), float(nu),
            int(seed) & (2 ** 64 - 1), N.ptr(u_t), ustride, N.ptr(traj), cap, N.ptr(traj_len), N.ptr(presence),
            N.ptr(total_steps), 1 if exact else 0)
    if workspace is None:
        N.check(lib.ssrs_step_tracks(*args, N.current_stream()), "ssrs_step_tracks")
    else:
        N.check(lib.ssrs_step_tracks_phased(*args, N.ptr(workspace), workspace.numel() * workspace.element_size(),
                                            int(first_phase_steps), N.current_stream()), "ssrs_step_tracks_phased")


def walk_pays_off(track_count: int, grid_shape, memory_parameter: int = 1, scaling_parameter: float = 1.0) -> bool:
    """Policy for the transition-table walk (`simulate_tracks_batch(walk=True)`): the table costs about as much as
    stepping every (cell, previous move) pair twice, so it pays when the batch is expected to take several times more
    steps than the grid has pairs.  Decided from the GLOBAL track count and the grid only, never from a rank's share:
    the two stepping paths map random words to steps differently, so every rank of a sharded run must take the same
    one for the result to be independent of the number of GPUs."""
    rows, cols = int(grid_shape[0]), int(grid_shape[1])
    if memory_parameter != 1 or scaling_parameter != 1.0:
        return False
    return int(track_count) * (rows + cols) // 2 >= 8 * rows * cols


def build_transition_table(fields, move_dirn: float, out=None):
    """Transition table of one (fields, direction): uint8 CUDA tensor of `ssrs_walk_table_bytes` bytes (64 B per cell),
    written by `ssrs_transition_table` on the current stream.  Valid for track_dirn_restrict = 1, nu = 1."""
    torch = N.require_cuda()
    lib = N.load()
    if fields is None or fields.dim() != 3 or fields.shape[2] != 2:
        raise ValueError("fields must be the interleaved [rows, cols, 2] tensor of interleave_fields")
    rows, cols = int(fields.shape[0]), int(fields.shape[1])
    nbytes = int(lib.ssrs_walk_table_bytes(rows, cols))
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    elif out.numel() * out.element_size() < nbytes or not out.is_cuda:
        raise ValueError("out is too small for the transition table")
    dirp = get_directional_probs(move_dirn * np.pi / 180.0)
    N.check(lib.ssrs_transition_table(N.ptr(fields), rows, cols, (C.c_double * 9)(*dirp.tolist()), N.ptr(out),
                                      N.current_stream()), "ssrs_transition_table")
    return out


def simulate_tracks_batch(move_dirn: float, start_rows, start_cols, grid_shape, memory_parameter: int = 1,
                          scaling_parameter: float = 1.0, fields=None, updraft_field=None, potential_field=None,
                          seed: int = 0, track_id0: int = 0, uniforms=None, record: bool = False,
                          traj_cap: Optional[int] = None, presence=None, total_steps=None,
                          exact: bool = False, walk: bool = False, table=None, workspace=None,
                          first_phase_steps: int = 0, phased: bool = False) -> TrackBatchResult:
    """All tracks of one (case, realisation) in a single launch.

    `start_rows`, `start_cols`: host arrays of start cells, or `start_rows` = a CUDA int32 tensor [n, 2] of (row, col)
    with `start_cols=None` (no upload per launch).
    `fields` is a pre-interleaved device tensor from `interleave_fields`; alternatively give
    `updraft_field` and `potential_field`; with neither the 'drw' model runs (reference :298-299).
    `uniforms` ([n_tracks, stride] float64) switches on verification mode; otherwise Philox keyed by
    (seed, track_id0 + i, step).  `exact=True` forces the reference's exact operation order in production mode
    (verification mode always uses it).  `presence` (int32 CUDA tensor [rows, cols]) is accumulated into if given,
    else a fresh raster is created.
    `phased=True` runs the launch in phases with survivor compaction (`ssrs_step_tracks_phased`: bit-identical results,
    full warps; `workspace` as for the walk is created when not given).
    `walk=True` takes the transition-table walk (`ssrs_walk_tracks`; needs fields, memory 1, nu 1, no uniforms, no
    recording): `table` from `build_transition_table` is built here when not given, `workspace` (uint8 CUDA tensor of
    `ssrs_walk_workspace_bytes(n)` bytes) likewise.  See `walk_pays_off` for when it is worth it.
    `record=True` stores the trajectories (step-major int16 [cap, n, 2]).  Track lengths are heavy-tailed (the longest
    of 100k tracks on 5000 x 6000 cells has ~1e5 points against a mean of 1e4), so without an explicit `traj_cap` the
    recording is two-pass: a first launch without trajectory or presence output yields the exact lengths (the random
    streams are counter-based or caller-supplied, so the second launch repeats it step for step), the second records
    with cap = longest track.  For batches whose buffer would exceed TRAJ_BUFFER_BYTES use `record_tracks_packed`.
    """
    torch = N.require_cuda()
    N.load()
    rows, cols = int(grid_shape[0]), int(grid_shape[1])
    start = None
    if isinstance(start_rows, torch.Tensor) and start_cols is None:
        # already on the device: int32 [n, 2] (row, col), validated by whoever built it (no host round trip per launch)
        start = start_rows
        if not start.is_cuda or start.dtype != torch.int32 or start.dim() != 2 or start.shape[1] != 2 or not start.is_contiguous():
            raise ValueError("a device start array must be a contiguous CUDA int32 tensor [n, 2]")
        n = int(start.shape[0])
    else:
        sr = np.asarray(start_rows).astype(np.int64).ravel()
        sc = np.asarray(start_cols).astype(np.int64).ravel()
        if sr.shape != sc.shape:
            raise ValueError("start_rows and start_cols differ in length")
        n = sr.size
        if n and (sr.min() < 0 or sr.max() >= rows or sc.min() < 0 or sc.max() >= cols):
            raise ValueError("start location outside the grid")
    if fields is None and updraft_field is not None:
        if potential_field is None:
            raise ValueError("updraft_field without potential_field is not supported")
        fields = interleave_fields(updraft_field, potential_field)
    if fields is not None and tuple(fields.shape) != (rows, cols, 2):
        raise ValueError(f"fields shape {tuple(fields.shape)} does not match grid {(rows, cols)}")
    if start is None:
        start = torch.from_numpy(np.stack([sr, sc], axis=1).astype(np.int32)).to("cuda")
    dirp = get_directional_probs(move_dirn * np.pi / 180.0)
    dirp_c = (C.c_double * 9)(*dirp.tolist())
    u_t, ustride = None, 0
    if uniforms is not None:
        un = np.ascontiguousarray(uniforms, dtype=np.float64)
        if un.ndim != 2 or un.shape[0] != n:
            raise ValueError("uniforms must be [n_tracks, stride]")
        u_t, ustride = torch.from_numpy(un).to("cuda"), un.shape[1]
    traj_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    if walk:
        if fields is None or memory_parameter != 1 or scaling_parameter != 1.0 or uniforms is not None or record or exact:
            raise ValueError("walk=True needs fields, memory_parameter 1, scaling_parameter 1, Philox streams, "
                             "record=False and exact=False")
        lib = N.load()
        if table is None:
            table = build_transition_table(fields, move_dirn)
        ws_bytes = int(lib.ssrs_walk_workspace_bytes(n))
        if workspace is None:
            workspace = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        if presence is None:
            presence = torch.zeros((rows, cols), dtype=torch.int32, device="cuda")
        if total_steps is None:
            total_steps = torch.zeros(1, dtype=torch.int64, device="cuda")
        N.check(lib.ssrs_walk_tracks(N.ptr(table), N.ptr(fields), rows, cols, N.ptr(start), n, int(track_id0), dirp_c,
                                     int(seed) & (2 ** 64 - 1), N.ptr(traj_len), N.ptr(presence), N.ptr(total_steps),
                                     N.ptr(workspace), workspace.numel() * workspace.element_size(),
                                     int(first_phase_steps), N.current_stream()), "ssrs_walk_tracks")
        res = TrackBatchResult(n, (rows, cols), None, traj_len, presence, total_steps, 0)
        res._keepalive = (table, workspace, start)          # the launches are asynchronous
        return res
    common = (fields, rows, cols, start, n, track_id0, dirp_c, memory_parameter, scaling_parameter, seed, u_t, ustride)
    traj = None
    cap = 0
    if record:
        if traj_cap:
            cap = int(traj_cap)
        else:
            _launch_steps(*common, None, 0, traj_len, None, None, exact)            # pass 1: lengths only
            cap = max(1, int(traj_len.abs().max().item())) if n else 1
            if 4 * cap * n > TRAJ_BUFFER_BYTES:
                raise ValueError(f"recording {n} tracks (longest {cap} points) needs a {4 * cap * n / 2 ** 30:.1f} GiB "
                                 f"trajectory buffer; use record_tracks_packed (chunked) or pass traj_cap explicitly")
        traj = torch.zeros((cap, n, 2), dtype=torch.int16, device="cuda")
    if presence is None:
        presence = torch.zeros((rows, cols), dtype=torch.int32, device="cuda")
    if total_steps is None:
        total_steps = torch.zeros(1, dtype=torch.int64, device="cuda")
    if phased and workspace is None:
        workspace = torch.empty(int(N.load().ssrs_walk_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
    _launch_steps(*common, traj, cap, traj_len, presence, total_steps, exact, workspace if phased else None,
                  first_phase_steps)
    res = TrackBatchResult(n, (rows, cols), traj, traj_len, presence, total_steps, cap)
    res._keepalive = (workspace, start, u_t)                # the launches are asynchronous
    return res


def record_tracks_packed(move_dirn: float, start_rows, start_cols, grid_shape, memory_parameter: int = 1,
                         scaling_parameter: float = 1.0, fields=None, seed: int = 0, track_id0: int = 0,
                         lengths=None, exact: bool = False, buffer_bytes: int = TRAJ_BUFFER_BYTES):
    """Trajectories of a large batch as (offsets int64 [n + 1], points int16 [total, 2]) on the host, recorded in
    contiguous chunks of track ids sized so that every chunk's step-major buffer (4 B x its longest track x its
    tracks) stays below `buffer_bytes`.  `lengths` (int array [n], points per track) from a previous
    `simulate_tracks_batch(...).traj_len` of the same (seed, track_id0) saves the measuring launch.  Presence is not
    accumulated here (the counting launch already did); Philox streams make every relaunch identical."""
    sr = np.asarray(start_rows).astype(np.int64).ravel()
    sc = np.asarray(start_cols).astype(np.int64).ravel()
    n = sr.size
    if lengths is None:
        res = simulate_tracks_batch(move_dirn, sr, sc, grid_shape, memory_parameter, scaling_parameter, fields=fields,
                                    seed=seed, track_id0=track_id0, exact=exact)
        lengths = res.traj_len.cpu().numpy()
    lengths = np.asarray(lengths, dtype=np.int64)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    points = np.empty((int(offsets[-1]), 2), dtype=np.int16)
    lo = 0
    while lo < n:
        hi, cap = lo, 0
        while hi < n and 4 * max(cap, int(lengths[hi])) * (hi - lo + 1) <= buffer_bytes:
            cap = max(cap, int(lengths[hi]))
            hi += 1
        if hi == lo:                                    # one track longer than the whole budget: record it alone
            cap, hi = int(lengths[lo]), lo + 1
        res = simulate_tracks_batch(move_dirn, sr[lo:hi], sc[lo:hi], grid_shape, memory_parameter, scaling_parameter,
                                    fields=fields, seed=seed, track_id0=track_id0 + lo, record=True, traj_cap=cap,
                                    exact=exact)
        off_c, pts_c = res.packed()
        if not np.array_equal(off_c[1:] - off_c[:-1], lengths[lo:hi]):
            raise RuntimeError("record_tracks_packed: relaunch produced different track lengths")
        points[offsets[lo]:offsets[hi]] = pts_c
        lo = hi
    return offsets, points


def generate_simulated_tracks(move_dirn: float, start_location, grid_shape, memory_parameter: int = 1,
                              scaling_parameter: float = 1.0, updraft_field=None, potential_field=None):
    """One track, reference signature (:264-318).  Consumes numpy's *global* random stream exactly like the
    reference (one `random_sample()` per step), so `np.random.seed(s)` followed by this call reproduces the
    reference's trajectory when given the same float32 fields."""
    rows, cols = int(grid_shape[0]), int(grid_shape[1])
    chunk = int(4 * max(rows, cols))
    state = np.random.get_state()
    while True:
        np.random.set_state(state)
        u = np.random.random_sample(chunk)[None, :]
        # a track that needs more than `chunk` uniforms stops at the end of the stream and reports a negative length
        res = simulate_tracks_batch(move_dirn, [start_location[0]], [start_location[1]], grid_shape, memory_parameter,
                                    scaling_parameter, updraft_field=updraft_field, potential_field=potential_field,
                                    uniforms=u, record=True, traj_cap=chunk + 1)
        length = int(res.traj_len.cpu().numpy()[0])
        if length > 0:
            break
        chunk *= 4
    np.random.set_state(state)
    if length > 1:
        np.random.random_sample(length - 1)      # leave the global stream where the reference would
    return res.tracks()[0]


def compute_presence_counts(tracks: Sequence[np.ndarray], gridshape: Tuple[int, int]) -> np.ndarray:
    """Reference :410-419 on the GPU.  Returns int32 counts (the reference's int16 wraps above 32767)."""
    torch = N.require_cuda()
    lib = N.load()
    rows, cols = int(gridshape[0]), int(gridshape[1])
    n = len(tracks)
    presence = torch.zeros((rows, cols), dtype=torch.int32, device="cuda")
    if n == 0:
        return presence.cpu().numpy()
    lens = np.array([len(t) for t in tracks], dtype=np.int32)
    cap = int(max(1, lens.max()))
    host = np.zeros((cap, n, 2), dtype=np.int16)
    for i, t in enumerate(tracks):
        host[:lens[i], i, :] = np.asarray(t, dtype=np.int16).reshape(-1, 2)
    traj = torch.from_numpy(host).to("cuda")
    tl = torch.from_numpy(lens).to("cuda")
    N.check(lib.ssrs_presence_counts(N.ptr(traj), cap, N.ptr(tl), n, rows, cols, N.ptr(presence),
                                     N.current_stream()), "ssrs_presence_counts")
    return presence.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# stage 2 boundary: MovModel
# ---------------------------------------------------------------------------------------------
class MovModel:
    """Fluid-flow movement model (reference :10-128)."""

    def __init__(self, move_dirn: float, grid_shape: Tuple[int, int]):
        self.move_dirn = move_dirn
        self.grid_shape = grid_shape

    def get_boundary_nodes(self):
        """Dirichlet node ids (column-major `col*nrow + row`) and values (reference :21-57): the low (0)
        and high (1000) edges depend on the quadrant of `move_dirn`; the concatenated list is split by
        position at `size // 2`, which for oblique directions mislabels a node or two — kept."""
        nrow, ncol = self.grid_shape
        cols_all = np.arange(ncol, dtype=np.int64)
        rows_in = np.arange(1, nrow - 1, dtype=np.int64)
        edge = {'N': cols_all * nrow + (nrow - 1), 'S': cols_all * nrow,
                'W': rows_in, 'E': (ncol - 1) * nrow + rows_in}
        frac = (self.move_dirn % 90.0) / 90.0
        quad = int((self.move_dirn % 360) // 90.0)
        cl, rl = round(ncol * frac), round(nrow * frac)
        N_, S_, W_, E_ = edge['N'], edge['S'], edge['W'], edge['E']
        a = (N_[cl:], E_[nrow - rl:])
        b = (S_[:ncol - cl], W_[:rl])
        c = (S_[ncol - cl:], E_[:nrow - rl])
        d = (N_[:cl], W_[rl:])
        low, high = {0: (a, b), 1: (c, d), 2: (b, a), 3: (d, c)}[quad]
        nodes = np.concatenate(low + high)
        vals = np.zeros(nodes.size)
        vals[nodes.size // 2:] = 1000.0
        return nodes, vals

    def assemble_sparse_linear_system(self):
        """Adjacency of the 8-neighbour graph in the reference's COO form (:59-84): (row_inds uint32,
        col_inds uint32, facs float32).  The GPU solver is matrix-free and does not need these arrays;
        they are produced (vectorised) for callers that inspect them.  Entry order within a row follows
        the reference's neighbour order, including the last-column factor quirk."""
        nrow, ncol = self.grid_shape
        n = nrow * ncol
        i = np.arange(n, dtype=np.int64)
        r = i % nrow
        # neighbour id offsets in reference order: W NW N NE E SE S SW ; N/S boundary rows use 5-long lists
        full = np.array([-nrow, -nrow + 1, 1, nrow + 1, nrow, nrow - 1, -1, -nrow - 1])
        north = np.array([nrow, nrow - 1, -1, -nrow - 1, -nrow])
        south = np.array([-nrow, -nrow + 1, 1, nrow + 1, nrow])
        rows_l, cols_l, facs_l = [], [], []
        for sel, offs in ((r == nrow - 1, north), ((r == 0) & (nrow > 1), south), ((r > 0) & (r < nrow - 1), full)):
            ids = i[sel]
            if ids.size == 0:
                continue
            nb = ids[:, None] + offs[None, :]
            ok = (nb >= 0) & (nb < n)
            pos = np.cumsum(ok, axis=1) - 1                      # position within the filtered list
            fac = np.where(pos % 2 == 1, sqrt(2.0), 1.0)
            rows_l.append(np.repeat(ids, ok.sum(axis=1)))
            cols_l.append(nb[ok])
            facs_l.append(fac[ok])
        order = np.argsort(np.concatenate(rows_l), kind='stable')
        return (np.concatenate(rows_l)[order].astype('u4'), np.concatenate(cols_l)[order].astype('u4'),
                np.concatenate(facs_l)[order].astype('f4'))

    @classmethod
    def solve_sparse_linear_system(cls, conductivity, bnodes, benergy, row_inds=None, col_inds=None, facs=None,
                                   **solver_opts) -> np.ndarray:
        """Potential at all nodes, float32 [nrow, ncol] (reference :86-128), solved on the GPU by the
        matrix-free AMG-preconditioned Krylov solver.  `row_inds/col_inds/facs` are accepted for signature
        compatibility and ignored: the operator they describe is built into the kernels."""
        from .potential import solve_potential_nodes
        return solve_potential_nodes(conductivity, bnodes, benergy, **solver_opts)
