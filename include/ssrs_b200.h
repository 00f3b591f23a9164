/* ssrs_b200 — C-ABI of the B200-native SSRS hot path.
 *
 * The reference (NREL/SSRS) has no FFI layer: its hot path is the set of Python functions in
 * ssrs/layers.py and ssrs/movmodel.py that ssrs/simulator.py:21-28 imports.  Each entry point below
 * replaces one of those functions (cited per function as file:line under /root/reference) and is what
 * a ctypes binding inside the reference would call (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - rasters are C-contiguous [row][col], row 0 = south, as the reference's flipped rasters
 *     (ssrs/raster.py:49); gridsize = (rows, cols) = (ysize, xsize) (ssrs/simulator.py:71-73);
 *   - the caller owns all memory; nothing here allocates persistent device memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     asynchronous on it unless stated otherwise;
 *   - return value: 0 on success, negative ssrs_status on failure; ssrs_last_error() gives the text.
 *     Nothing throws across this boundary and there is no CPU fallback: without a CUDA device every
 *     compute entry point returns SSRS_ERR_CUDA.
 */
#ifndef SSRS_B200_H
#define SSRS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSRS_ABI_VERSION 1

#if defined(__GNUC__)
#define SSRS_API __attribute__((visibility("default")))
#else
#define SSRS_API
#endif

typedef enum ssrs_status {
    SSRS_OK = 0,
    SSRS_ERR_INVALID = -1,     /* bad argument (the Python layer raises ValueError) */
    SSRS_ERR_CUDA = -2,        /* CUDA runtime / driver error, or no device */
    SSRS_ERR_UNSUPPORTED = -3, /* valid request outside the implemented range */
    SSRS_ERR_NOT_CONVERGED = -4
} ssrs_status;

SSRS_API int ssrs_abi_version(void);
SSRS_API const char* ssrs_last_error(void);
/* sm_count / compute capability of the current device */
SSRS_API int ssrs_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Stage 1 — orographic updraft.  Replaces, fused in one pass over the DEM:
 *   compute_slope_degrees      ssrs/layers.py:63-93
 *   compute_aspect_degrees     ssrs/layers.py:96-128
 *   compute_orographic_updraft ssrs/layers.py:11-22
 *   get_above_threshold_speed  ssrs/layers.py:171-185   (on the float32-rounded orograph, as
 *                              ssrs/simulator.py:198,233 round-trips it through a float32 .npy)
 * wspeed/wdirn: per-cell wind [rows][cols] (snapshot/seasonal modes) or both NULL to use the
 * uniform scalars (ssrs/simulator.py:194-195).  Any of the four outputs may be NULL.
 * Border cells of every output are 0 (nan_to_num of the reference's NaN border).
 */
SSRS_API int ssrs_updraft(const float* dem, int rows, int cols, float resolution,
                 const float* wspeed, const float* wdirn,
                 float uniform_wspeed, float uniform_wdirn_deg,
                 float threshold,
                 float* slope_deg, float* aspect_deg, float* orograph, float* updraft,
                 void* stream);

/* compute_orographic_updraft alone (ssrs/layers.py:11-22) on slope / aspect rasters in degrees — the reference's
 * signature for callers that hold those rasters (3DEP 'Slope' / 'Aspect' layers, ssrs/simulator.py:152-168);
 * wspeed / wdirn: per-cell rasters or NULL for the uniform values.  out = max(V sin(slope) max(cos(aspect - dirn), 0), min_updraft). */
SSRS_API int ssrs_orographic_updraft(const float* slope_deg, const float* aspect_deg, const float* wspeed, const float* wdirn,
                            float uniform_wspeed, float uniform_wdirn_deg, float min_updraft, float* out, int64_t n,
                            void* stream);

/* get_above_threshold_speed alone (ssrs/layers.py:171-185) for rasters that did not come from
 * ssrs_updraft (e.g. orograph + thermals, ssrs/simulator.py:236-242). */
SSRS_API int ssrs_threshold(const float* in, float* out, int64_t n, float threshold, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2 — directional potential.  Replaces
 *   MovModel.assemble_sparse_linear_system   ssrs/movmodel.py:59-84   (the operator is built into the kernels)
 *   MovModel.solve_sparse_linear_system      ssrs/movmodel.py:86-128  (SuperLU -> matrix-free AMG + BiCGStab)
 * conductivity: float32 [rows][cols] thresholded updraft (the `conductivity` argument of the reference).
 * bnodes_host / bvalues_host: HOST arrays exactly as MovModel.get_boundary_nodes() returns them
 *   (ssrs/movmodel.py:21-57): column-major node ids `col*rows + row` and their Dirichlet values.
 * rtol: relative 2-norm residual of the un-normalised system; <= 0 (default) iterates to the accuracy float64 can
 *   attain (the solver estimates the residual floor d_i*ulp(phi_i)/2 and stops within it); max_iter <= 0 -> 300.
 * potential: float32 [rows][cols] out (the reference returns float32, :128).
 * Workspace (~300 B/cell) comes from a per-device arena of large cudaMalloc blocks that stays cached between
 * solves (a seasonal run solves once per wind case); ssrs_release_workspace() returns it.  Synchronous.
 * Returns SSRS_ERR_NOT_CONVERGED (potential still written) if the tolerance was not reached.
 */
typedef struct ssrs_solve_stats {
    int32_t iterations;          /* BiCGStab iterations (two V-cycles + two operator applications each) */
    int32_t restarts;            /* true-residual restarts */
    int32_t levels;              /* AMG levels including the fine grid */
    int32_t converged;           /* 1: tolerance met; 2: stopped at the float64 attainable accuracy (true residual stagnated <= 1e-6) */
    double rel_residual;         /* final |b - A x|_2 / |b - A x0|_2, true residual */
    double setup_ms, solve_ms;   /* host wall clock around the synchronous phases */
    double operator_complexity;  /* sum of nnz over levels / fine nnz */
    int64_t level_rows[24];
    int64_t coarsest_rows;
    int64_t workspace_bytes;
} ssrs_solve_stats;

SSRS_API int ssrs_potential_solve(const float* conductivity, int rows, int cols,
                                  const int64_t* bnodes_host, const double* bvalues_host, int64_t n_bnodes,
                                  double rtol, int max_iter, float* potential, ssrs_solve_stats* stats,
                                  void* stream);

/* Grows the solver's cached workspace arena to what a solve of a rows x cols grid needs (~500 B/cell), so that the first
 * solve does not pay for the allocation (0.2-1.2 s at 5000 x 6000).  Optional; Simulator calls it in its constructor. */
SSRS_API int ssrs_reserve_workspace(int rows, int cols);

/* Frees the solver's cached workspace on the current device (not while a solve is running). */
SSRS_API int ssrs_release_workspace(void);

/* ssrs_potential_solve that also returns the float64 iterate the float32 potential was rounded from (potential64:
 * float64 [rows][cols]) — for diagnostics: the reference's tracks are steered by float32 rounding plateaus of the
 * potential (SURVEY.md §0 finding 4), so tests check the discrete maximum principle on the un-rounded solution. */
SSRS_API int ssrs_potential_solve_f64(const float* conductivity, int rows, int cols,
                             const int64_t* bnodes_host, const double* bvalues_host, int64_t n_bnodes,
                             double rtol, int max_iter, float* potential, double* potential64,
                             ssrs_solve_stats* stats, void* stream);

/* Row-sharded solve (BASELINE config 5): the grid's rows are split into `comm->size` contiguous
 * slabs; rank r owns slab r.  Every rank passes the SAME full conductivity raster (fields are replicated for the
 * stepping stage anyway) and receives the full potential.  While a level is distributed (>= 1e6 rows) every rank
 * builds and applies ITS rows of the hierarchy only — aggregates never straddle a slab boundary, so the setup needs
 * the neighbours' aggregate ids along the slab boundary and a few per-part integers, nothing else —; the solve phase
 * — V-cycles, operator applications, Krylov vectors — exchanges one-row halos (fine level) or the few
 * boundary-adjacent entries (coarse levels) with the two neighbouring ranks before each operator application,
 * plus one scalar all-reduce per inner product.  The first level below 1e6 rows is all-gathered and everything
 * from there down is computed by every rank.  The result is bit-identical to a setup in which every rank builds
 * everything (SSRS_X_REDUNDANT_SETUP=1 selects that, for comparison).
 *
 * ssrs_comm is the transport: plain function pointers, so the product uses NCCL over NVLink
 * (ssrs_comm_create_nccl) and the CPU test build drives the same solver code over gloo.
 * All three callbacks return 0 on success.  Offsets and sizes are in bytes. */
typedef struct ssrs_comm {
    int32_t rank, size;
    void* ctx;
    /* exchange contiguous ranges of ONE buffer with rank-1 ("up") and rank+1 ("down"); a size of 0 skips that
     * transfer; stream-ordered */
    int (*exchange)(void* ctx, void* base,
                    int64_t send_up_off, int64_t send_up_bytes, int64_t recv_up_off, int64_t recv_up_bytes,
                    int64_t send_dn_off, int64_t send_dn_bytes, int64_t recv_dn_off, int64_t recv_dn_bytes,
                    void* stream);
    /* in-place sum of `count` (<= 4 * SSRS_MAX_RANKS) HOST doubles over all ranks; blocking; bit-identical result on every rank */
    int (*allreduce_sum)(void* ctx, double* values_host, int32_t count, void* stream);
    /* in-place all-gather: rank r contributes bytes [offsets_host[r], offsets_host[r+1]) of `base` */
    int (*allgather)(void* ctx, void* base, const int64_t* offsets_host, void* stream);
    /* in-place sum of a device uint32 array over all ranks (presence maps); stream-ordered */
    int (*allreduce_u32)(void* ctx, uint32_t* values, int64_t count, void* stream);
} ssrs_comm;

#define SSRS_MAX_RANKS 16

SSRS_API int ssrs_potential_solve_sharded(const float* conductivity, int rows, int cols,
                                          const int64_t* bnodes_host, const double* bvalues_host, int64_t n_bnodes,
                                          double rtol, int max_iter, float* potential, ssrs_solve_stats* stats,
                                          const ssrs_comm* comm, void* stream);

/* NCCL transport (one process per GPU).  Rank 0 calls ssrs_nccl_unique_id and distributes the 128 bytes by any
 * means (the Python host uses torch.distributed); every rank then calls ssrs_comm_create_nccl collectively. */
SSRS_API int ssrs_nccl_unique_id(void* id128_host);
SSRS_API int ssrs_comm_create_nccl(const void* id128_host, int rank, int size, ssrs_comm** comm_out);
SSRS_API int ssrs_comm_destroy(ssrs_comm* comm);
/* How this communicator exchanges halos: 1 = over peer memory (each rank maps its neighbours' staging blocks through
 * CUDA IPC; one kernel per exchange stores the boundary ranges into the neighbours' memory over NVLink, releases a
 * sequence number there, waits for theirs and copies the arrived ranges into the ghost entries; chosen with
 * SSRS_COMM_HALO=peer in the environment of ssrs_comm_create_nccl, when every rank can map its neighbours),
 * 0 = grouped ncclSend/ncclRecv (the default), -1 = not a communicator of ssrs_comm_create_nccl. */
SSRS_API int ssrs_comm_halo_mode(const ssrs_comm* comm);

/* Presence-map reduction over the ranks that stepped disjoint blocks of tracks: in-place
 * sum of the uint32 count raster, one collective per (case, realisation) map.  Counts are integers, so the
 * result is independent of the number of ranks. */
SSRS_API int ssrs_presence_allreduce(uint32_t* presence, int64_t n, const ssrs_comm* comm, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 3+4 — batched track stepping with fused presence accumulation.  Replaces
 *   generate_simulated_tracks  ssrs/movmodel.py:264-318  (one call per track in the reference,
 *                              mapped over a process pool at ssrs/simulator.py:360-369)
 *   compute_presence_counts    ssrs/movmodel.py:410-419  (fused: every appended point is counted)
 *
 * fields: interleaved {updraft, potential} pairs [rows][cols][2] built by ssrs_interleave_fields,
 *         or NULL for the 'drw' movement model (no fields, ssrs/simulator.py:370-381).
 * start_rc: int32 [n_tracks][2] (row, col) from get_starting_indices (ssrs/movmodel.py:144-182).
 * dirprob9_host: HOST pointer to the 9 weights of get_directional_probs(move_dirn*pi/180)
 *         (ssrs/movmodel.py:247-257), computed by the caller exactly as the reference does.
 * memory, nu: track_dirn_restrict and track_stochastic_nu (ssrs/config.py:56-57).
 * Random numbers: if uniforms != NULL ("verification mode") step k of track t consumes
 *         uniforms[t*uniforms_stride + k] — the reference's pre-drawn np.random stream (uniforms_stride must cover
 *         the longest track: a track that needs step k >= uniforms_stride stops there and reports
 *         traj_len[t] = -(points so far), so the caller can retry with a longer stream); otherwise
 *         Philox4x32-10 keyed by seed with counter (track_id0 + t, k), so results do not depend on
 *         how tracks are sharded over GPUs.
 * flags: 0, or SSRS_STEP_EXACT to evaluate the probabilities in the reference's exact operation order
 *         (bit-for-bit numpy; always used in verification mode).  The default production arithmetic cancels
 *         the normalisations (same distribution, ~7x fewer float64 operations).
 * traj (optional): int16 [traj_cap][n_tracks][2] step-major (row, col); points beyond traj_cap are
 *         not stored but still stepped and counted.   traj_len (optional): int32 [n_tracks] number of
 *         trajectory points (= steps + 1).   presence (optional): uint32 [rows][cols], incremented
 *         atomically (not cleared).   total_steps (optional): one uint64, incremented by the number
 *         of track-steps taken (loop iterations at ssrs/movmodel.py:285-317).
 * Asynchronous on `stream`.  Tracks are handed to the lanes through a device counter (8 bytes per launch, taken
 * from an 8 KB per-device array the library allocates on first use and zeroes on `stream`): which lane steps which
 * track varies from run to run, the results do not (see above).
 */
#define SSRS_STEP_EXACT 1

SSRS_API int ssrs_step_tracks(const float* fields, int rows, int cols,
                     const int32_t* start_rc, int64_t n_tracks, int64_t track_id0,
                     const double* dirprob9_host, int memory, double nu,
                     uint64_t seed, const double* uniforms, int64_t uniforms_stride,
                     int16_t* traj, int64_t traj_cap, int32_t* traj_len,
                     uint32_t* presence, unsigned long long* total_steps,
                     int flags, void* stream);

/* ssrs_step_tracks as a PHASED launch (track_dirn_restrict = 1; other values run as one launch): track lengths are
 * heavy-tailed — at 5000 x 6000 the median track takes 9e3 steps, 1 % take more than 3e4 and the longest 1.2e5 — so a
 * launch that steps every track to its end runs most of its life with a few lanes per warp alive.  Here the stepping
 * is cut at fixed step counts (first cut after about one crossing of the grid's short side, then x1.25; or
 * first_phase_steps > 0): each phase's kernel appends the survivors' states (16 bytes) to a compact list, the next
 * phase's kernel packs them into full warps again and its surplus CTAs exit at once, which frees SMs for other
 * streams' launches.  Same arguments and bit-identical results as ssrs_step_tracks (the random stream is keyed by
 * track id and step); workspace = ssrs_walk_workspace_bytes(n_tracks) bytes, caller-owned, reusable once `stream`
 * has passed the call. */
/* Number of phases (= kernel launches) ssrs_step_tracks_phased / ssrs_walk_tracks use on a rows x cols grid. */
SSRS_API int ssrs_step_phase_count(int rows, int cols, int first_phase_steps);

SSRS_API int ssrs_step_tracks_phased(const float* fields, int rows, int cols,
                     const int32_t* start_rc, int64_t n_tracks, int64_t track_id0,
                     const double* dirprob9_host, int memory, double nu,
                     uint64_t seed, const double* uniforms, int64_t uniforms_stride,
                     int16_t* traj, int64_t traj_cap, int32_t* traj_len,
                     uint32_t* presence, unsigned long long* total_steps,
                     int flags, void* workspace, int64_t workspace_bytes, int first_phase_steps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 3+4 for large batches — transition-table walk (ssrs_b200/csrc/walk.cu).  Same reference functions as
 * ssrs_step_tracks (generate_simulated_tracks ssrs/movmodel.py:264-318 mapped over the pool at
 * ssrs/simulator.py:360-369; compute_presence_counts ssrs/movmodel.py:410-419), for the default movement settings
 * track_dirn_restrict = 1 and track_stochastic_nu = 1 (ssrs/config.py:56-57), where the move distribution of
 * movmodel.py:294-312 depends on the cell and the previous move only:
 *   ssrs_transition_table  evaluates it once per (cell, previous move) from the interleaved fields — production
 *                          arithmetic of ssrs_step_tracks, fallback chain of movmodel.py:228-240 included — as two 31-bit
 *                          cumulative thresholds; table = ssrs_walk_table_bytes(rows, cols) = 64 bytes per cell;
 *   ssrs_walk_tracks       steps every track with one 8-byte gather, two integer compares on a Philox word and one
 *                          presence increment per step.  Cells within two rows / one column of the border, a track's
 *                          first steps and steps near its step limit take ssrs_step_tracks' general step on `fields`.
 *                          The walk is cut into phases at fixed step counts; survivors are compacted between phases
 *                          (workspace = ssrs_walk_workspace_bytes(n_tracks), caller-owned, reusable once the stream
 *                          has passed the call).  first_phase_steps: 0 = default schedule (first cut after about one
 *                          crossing of the grid's short side, then x1.25), else the first cut (tests).
 * Random numbers: Philox4x32-10 keyed by seed with counter (track_id0 + t, step): results do not depend on sharding or
 * on the phase schedule.  The mapping of words to steps differs from ssrs_step_tracks (documented in walk.cu and
 * restated by oracle/ssrs_oracle.c), so the two entry points give different realisations of the same distribution;
 * probabilities are quantised to 2^-31.  traj_len, total_steps optional; presence required (uint32, not cleared).
 * Asynchronous on `stream`. */
SSRS_API int64_t ssrs_walk_table_bytes(int rows, int cols);
SSRS_API int64_t ssrs_walk_workspace_bytes(int64_t n_tracks);
SSRS_API int ssrs_transition_table(const float* fields, int rows, int cols, const double* dirprob9_host, void* table,
                          void* stream);
SSRS_API int ssrs_walk_tracks(const void* table, const float* fields, int rows, int cols,
                     const int32_t* start_rc, int64_t n_tracks, int64_t track_id0,
                     const double* dirprob9_host, uint64_t seed,
                     int32_t* traj_len, uint32_t* presence, unsigned long long* total_steps,
                     void* workspace, int64_t workspace_bytes, int first_phase_steps, void* stream);

/* {updraft, potential} -> interleaved pairs (one 8-byte gather per cell in the stepping kernel). */
SSRS_API int ssrs_interleave_fields(const float* updraft, const float* potential, float* fields,
                           int64_t n, void* stream);

/* compute_presence_counts (ssrs/movmodel.py:410-419) for stored trajectories:
 * traj int16 [traj_cap][n_tracks][2] step-major + traj_len as written by ssrs_step_tracks. */
SSRS_API int ssrs_presence_counts(const int16_t* traj, int64_t traj_cap, const int32_t* traj_len,
                         int64_t n_tracks, int rows, int cols, uint32_t* presence, void* stream);

/* Step-major trajectories as written by ssrs_step_tracks -> the packed on-disk form (ssrs_b200/trackio.py): point k of
 * track t goes to points[offsets[t] + k] for k < traj_len[t]; offsets int64 [n_tracks] = exclusive running sum of
 * traj_len (all lengths <= traj_cap); points int16 [sum traj_len][2]. */
SSRS_API int ssrs_pack_trajectories(const int16_t* traj, int64_t traj_cap, const int32_t* traj_len, const int64_t* offsets,
                           int64_t n_tracks, int16_t* points, void* stream);

/* Running sums of a uint32 count raster along each row, the input of ssrs_smooth_presence:
 * row_prefix int64 [rows][cols+1], row_prefix[r][0] = 0, row_prefix[r][c+1] = counts[r][0] + ... + counts[r][c]. */
SSRS_API int ssrs_row_prefix_sums(const uint32_t* counts, int rows, int cols, long long* row_prefix, void* stream);

/* compute_smooth_presence_counts (ssrs/movmodel.py:422-439): disk kernel of `radius` cells, zero padding
 * ('same'), normalised by the kernel's cell count.  row_prefix: int64 [rows][cols+1] running sums of the
 * counts along each row (row_prefix[r][c] = sum counts[r][0..c-1]); out: float32 [rows][cols]. */
SSRS_API int ssrs_smooth_presence(const long long* row_prefix, int rows, int cols, int radius, float* out,
                                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * "Next" rows: the producers in front of stage 1.
 *
 * ssrs_interp_wind replaces Simulator._get_interpolated_wind_conditions / _interpolate_wtk_vardata
 * (ssrs/simulator.py:765-792, scipy griddata(method='linear') on the u/v components): px/py/east/north are
 * float64 [npoints] (projected site coordinates; east = speed*sin(dirn), north = speed*cos(dirn), :784-785),
 * triangles int32 [ntriangles][3] is the Delaunay triangulation of the sites (host: scipy.spatial.Delaunay, the
 * same Qhull griddata uses).  Cell (r, c) sits at (x0 + c*resolution, y0 + r*resolution) (get_terrain_grid,
 * :177-185).  owner_scratch: int32 [rows][cols].  Outputs float32 [rows][cols]; NaN outside the convex hull. */
SSRS_API int ssrs_interp_wind(const double* px, const double* py, const double* east, const double* north, int npoints,
                              const int32_t* triangles, int ntriangles, double x0, double y0, double resolution,
                              int rows, int cols, int32_t* owner_scratch, float* wspeed, float* wdirn, void* stream);

/* The same for Config.wtk_interp_type = 'nearest' (ssrs/config.py:60; scipy griddata(method='nearest'), a k-d tree
 * query): every cell takes the u/v of its closest site (float64 Euclidean distance; equal distances go to the lower
 * site index), defined everywhere — no triangulation, no NaN. */
SSRS_API int ssrs_interp_wind_nearest(const double* px, const double* py, const double* east, const double* north,
                                      int npoints, double x0, double y0, double resolution, int rows, int cols,
                                      float* wspeed, float* wdirn, void* stream);

/* The same for Config.wtk_interp_type = 'cubic' (scipy griddata(method='cubic') = CloughTocher2DInterpolator with
 * tol 1e-6, maxiter 400, no rescaling: gradients at the sites by Gauss-Seidel sweeps that minimise the curvature along
 * the triangulation's edges, then a C1 piecewise cubic on every triangle; NaN outside the convex hull).  Besides the
 * arguments of ssrs_interp_wind: neighbors int32 [ntriangles][3] (scipy.spatial.Delaunay.neighbors: the triangle
 * opposite vertex k, -1 on the hull) and the sites' adjacency in CSR form (Delaunay.vertex_neighbor_vertices:
 * vertex_nb_indptr int32 [npoints+1], vertex_nb_indices int32 [indptr[npoints]]), all on the device.  ct_scratch:
 * ssrs_interp_wind_cubic_scratch_bytes(npoints, ntriangles) bytes on the device, 32-byte aligned; on return (stream
 * order) it holds the gradients float64 [2][npoints][2] (east, north), the Bezier ordinates float64
 * [ntriangles][2][20] and, last, two int32 sweep counts (0 = the 400 sweeps did not reach the tolerance: scipy warns
 * and uses the last iterate, and so does this entry point). */
SSRS_API int64_t ssrs_interp_wind_cubic_scratch_bytes(int npoints, int ntriangles);
SSRS_API int ssrs_interp_wind_cubic(const double* px, const double* py, const double* east, const double* north, int npoints,
                                    const int32_t* triangles, const int32_t* neighbors, int ntriangles,
                                    const int32_t* vertex_nb_indptr, const int32_t* vertex_nb_indices, double x0, double y0,
                                    double resolution, int rows, int cols, int32_t* owner_scratch, void* ct_scratch,
                                    float* wspeed, float* wdirn, void* stream);

/* compute_thermals (ssrs/layers.py:188-214) in two steps: the random seeds (Philox4x32-10 keyed by (seed, cell):
 * same distribution as the reference's np.random draws, not the same stream) and the deterministic
 * scipy.ndimage.gaussian_filter(sigma, mode='constant', truncate) smoothing.  tmp: float32 [rows][cols];
 * weights_scratch: float32 [2*int(truncate*sigma+0.5)+1] on the device. */
SSRS_API int ssrs_thermal_seeds(const float* aspect, int rows, int cols, float thermal_intensity_scale, uint64_t seed,
                                float* seeds, void* stream);
SSRS_API int ssrs_gaussian_blur(const float* in, float* out, float* tmp, int rows, int cols, float sigma, float truncate,
                                float* weights_scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSRS_B200_H */
