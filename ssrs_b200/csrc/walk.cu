// Stage 3+4, production path for large batches: transition-table walk with phase compaction (sm_100a).
//
// The gather-and-evaluate stepper (tracks.cu) spends ~125 instructions and three dependent gathers per track-step on
// weights that depend only on (cell, previous move) — with the default direction memory of one move
// (track_dirn_restrict = 1, ssrs/config.py:57) and nu = 1 the move distribution of ssrs/movmodel.py:294-312 is a
// function of the cell and of the previous move alone.  A batch of 1e5 tracks takes 1e9 steps on 3e7 cells, so every
// (cell, move) pair is evaluated ~4 times over; here it is evaluated ONCE:
//
//   transition_table_kernel   per cell and previous move: the three candidates' weights in the production arithmetic
//                             of tracks.cu (float32 potential differences, float64 products), including the fallback
//                             chain of generate_move_probabilities (movmodel.py:228-240), normalised and stored as two
//                             31-bit cumulative thresholds {T1, T2}: 8 bytes per (cell, move), 64 bytes per cell
//                             (1.92 GB at 5000 x 6000), written once per (case, realisation) at HBM speed;
//   walk_kernel               per track-step: ONE 8-byte gather, two integer compares against a 31-bit Philox word, one
//                             `red.global.add` on the presence raster — ~25 instructions instead of ~125, one dependent
//                             load instead of a gather-evaluate chain.
//
// Track lengths are heavy-tailed (median 9e3 steps, 1 % above 3e4, maximum 1.2e5 at 5000 x 6000), so a launch that
// steps every track to completion runs most of its life with a few lanes per warp alive.  The walk is therefore cut
// into PHASES at fixed step counts (geometric, x1.25): a phase kernel steps its tracks up to the phase's cap and appends
// the survivors' states (16 bytes) to a compact list; the next phase's kernel packs them into full warps again and
// surplus CTAs exit at once, which frees the SMs for the next batch's launch on another stream.  Phases do not change
// results: the random stream is keyed by (seed, global track id, step).
//
// Step types (restated by oracle/ssrs_oracle.c, mode "table", which reproduces this file bit for bit on the CPU):
//   * a step is a TABLE step when the previous move is known, the cell is tabulated (2 <= row <= rows-3,
//     1 <= col <= cols-3: no burn-in relocation or exit test can apply, movmodel.py:205-217,287-291), at least four more
//     steps are allowed, and the track entered table mode at the last step index divisible by four and has not left it
//     since; it consumes word (k & 3) of Philox4x32-10 block (track, k >> 2, 0);
//   * every other step is the general production step of tracks.cu on the fields, drawing its 52-bit uniform from the
//     first two words of block (track, k, 1).
// Probabilities are quantised to 2^-31 (4.7e-10): the same distribution as the float64 cumulative comparison to that
// resolution; distributional parity with the reference is tested in tests/test_config1_parity.py.
//
// Compiled with -fmad=false like tracks.cu.
#include "stepper.cuh"

#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

namespace cg = cooperative_groups;

namespace ssrs {
namespace {

constexpr unsigned W_ONE = 0x80000000u;        // probability 1 on the 2^-31 lattice
constexpr unsigned W_UNMASKED = 0xFFFFFFFFu;   // no candidate has any weight: unmasked directional draw (movmodel.py:239-240)
constexpr unsigned W_BORDER = 0xFFFFFFFEu;     // cell not tabulated (relocation / exit rules may apply): general step

// round(p * 2^31) for p in [0, 1], saturating; the cast truncates a non-negative value (same in C)
__device__ __forceinline__ unsigned prob31(double p) {
    const double x = p * 2147483648.0 + 0.5;
    return x >= 2147483648.0 ? W_ONE : (unsigned)x;
}

// {T1, T2} of one (cell, previous move): the three candidates' weights exactly as choose_fast3 / fast_step evaluate them
// (q_i = max.NaN(d_i, 0) u_i s_a s_b), the directional fallback when none is positive or a difference is NaN, then
// cumulative probabilities on the 2^-31 lattice.  pick = (r < T1) ? 0 : (r < T2) ? 1 : 2 for a 31-bit uniform r.
__device__ __forceinline__ uint2 table_entry(double dq0, double dq1, double dq2, double s0, double s1, double s2,
                                             double g0, double g1, double g2) {
    double q0 = dq0 * (s1 * s2), q1 = dq1 * (s0 * s2), q2 = dq2 * (s0 * s1);
    double c1 = q0 + q1, c2 = c1 + q2;
    if (!(c2 > 0.0 && c2 <= 1.7976931348623157e308)) {       // all zero, or NaN / infinity: movmodel.py:228-236
        q0 = g0; q1 = g1; q2 = g2;
        c1 = q0 + q1; c2 = c1 + q2;
        if (!(c2 > 0.0)) return make_uint2(W_UNMASKED, 0u);
    }
    const double inv = 1.0 / c2;
    unsigned t1 = (q0 > 0.0) ? prob31(q0 * inv) : 0u;
    unsigned t2 = (q2 > 0.0) ? prob31(c1 * inv) : W_ONE;
    if (!(q1 > 0.0) && !(q2 > 0.0)) t1 = W_ONE;
    if (t2 < t1) t2 = t1;
    return make_uint2(t1, t2);
}

struct TableParams {
    const float2* fields;
    uint2* table;
    int rows, cols;
    double dirp[9];
};

__global__ void __launch_bounds__(256) transition_table_kernel(const TableParams T) {
    const long long n = (long long)T.rows * T.cols;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int nc = T.cols, nr = T.rows;
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < n; cell += stride) {
        const int r = (int)(cell / nc), c = (int)(cell - (long long)r * nc);
        uint4* out = reinterpret_cast<uint4*>(T.table + cell * 8);
        if (r < 2 || r > nr - 3 || c < 1 || c > nc - 3) {
            const uint4 b = make_uint4(W_BORDER, 0u, W_BORDER, 0u);
            out[0] = b; out[1] = b; out[2] = b; out[3] = b;
            continue;
        }
        const float2* base = T.fields + cell;
        const float2 fc = __ldg(base);
        const double uc = clip_updraft(fc.x);
        double dq[9], s[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            if (i == 4) { dq[i] = 0.0; s[i] = 0.0; continue; }
            const int dr = i / 3 - 1, dc = i % 3 - 1;
            const float2 f = __ldg(base + dr * nc + dc);
            const float ninv = (dr != 0 && dc != 0) ? 0.70710677f : 1.0f;          // movmodel.py:139-141
            const float d = __fmul_rn(__fsub_rn(fc.y, f.y), ninv);                 // float32, movmodel.py:301-304
            const double u = clip_updraft(f.x);
            dq[i] = (double)fmax_nan(d, 0.0f) * u;
            s[i] = uc + u;
        }
        uint2 e[8];
#pragma unroll
        for (int last = 0; last < 9; ++last) {
            if (last == 4) continue;
            // candidates (ascending flat index) of every previous move, by the move's flat index; entry 4 unused
            constexpr int C3[9][3] = {{0, 1, 3}, {0, 1, 2}, {1, 2, 5}, {0, 3, 6}, {0, 0, 0}, {2, 5, 8}, {3, 6, 7}, {6, 7, 8}, {5, 7, 8}};
            const int i0 = C3[last][0], i1 = C3[last][1], i2 = C3[last][2];
            e[last - (last > 4)] = table_entry(dq[i0], dq[i1], dq[i2], s[i0], s[i1], s[i2], T.dirp[i0], T.dirp[i1], T.dirp[i2]);
        }
        out[0] = make_uint4(e[0].x, e[0].y, e[1].x, e[1].y);
        out[1] = make_uint4(e[2].x, e[2].y, e[3].x, e[3].y);
        out[2] = make_uint4(e[4].x, e[4].y, e[5].x, e[5].y);
        out[3] = make_uint4(e[6].x, e[6].y, e[7].x, e[7].y);
    }
}

struct WalkParams {
    TrackParams tp;              // fields, start, traj_len, presence, total_steps, n_tracks, track_id0, dirp, rows, cols, burnin, kmax, rk
    const uint2* table;
    const uint4* in_list;        // {track, row | col << 16, k, previous move} of the survivors of the previous phase; NULL: fresh tracks
    uint4* out_list;
    unsigned* in_head;           // entries of the input taken so far
    const unsigned* in_count;    // entries in in_list (device counter written by the previous phase); NULL: tp.n_tracks
    unsigned* out_count;
    int kcap;                    // the phase steps every track up to step kcap (a multiple of 4)
    unsigned dthr[9];            // unmasked directional fallback: 31-bit cumulative thresholds in flat order
};

__global__ void __launch_bounds__(128, 8) walk_kernel(const WalkParams W) {
    const TrackParams& P = W.tp;
    const unsigned n_in = W.in_count != nullptr ? *W.in_count : (unsigned)P.n_tracks;
    if ((unsigned)(blockIdx.x * blockDim.x) >= n_in) return;       // surplus CTA of a late phase: nothing to step
    __shared__ int4 s_cand9[9];      // general step: by flat move, element offsets of the candidates, .w = flat indices (4 bits each)
    __shared__ int4 s_cand[8];       // table step: by slot, element offsets, .w = slots of the candidates
    if (threadIdx.x < 9) {
        const unsigned last = threadIdx.x;
        const unsigned c3 = (unsigned)((last < 5 ? (C3_A >> (12 * last)) : (C3_B >> (12 * (last - 5)))) & 0xFFFu);
        const int i0 = c3 & 15, i1 = (c3 >> 4) & 15, i2 = (c3 >> 8) & 15;
        const int4 off = make_int4((i0 / 3 - 1) * P.cols + (i0 % 3 - 1), (i1 / 3 - 1) * P.cols + (i1 % 3 - 1),
                                   (i2 / 3 - 1) * P.cols + (i2 % 3 - 1), (int)c3);
        s_cand9[last] = off;
        if (last != 4u)
            s_cand[slot_of(last)] = make_int4(off.x, off.y, off.z, (int)(slot_of(i0) | (slot_of(i1) << 4) | (slot_of(i2) << 8)));
    }
    __syncthreads();
    const int nr = P.rows, nc = P.cols, kmax = P.kmax;
    unsigned long long steps_local = 0;
    bool have = false, tmode = false, done = false;
    int row = 0, col = 0, k = 0;          // in table mode `row` holds the linear cell index and `last` the move's slot
    unsigned last = 4, t = 0;

    while (true) {
        if (!tmode) {
            if (!have && !done) {
                // next entry of the phase's input: lanes that ask together take consecutive entries with one atomic
                const auto g = cg::coalesced_threads();
                unsigned base = 0;
                if (g.thread_rank() == 0) base = atomicAdd(W.in_head, (unsigned)g.size());
                const unsigned idx = g.shfl(base, 0) + g.thread_rank();
                if (idx >= n_in) done = true;
                else {
                    if (W.in_list != nullptr) {
                        const uint4 e = W.in_list[idx];
                        t = e.x; row = (int)(e.y & 0xFFFFu); col = (int)(e.y >> 16); k = (int)e.z; last = e.w;
                    } else {
                        t = idx;
                        const int2 s = __ldg(P.start + t);
                        row = s.x; col = s.y; k = 0; last = 4;
                        atomicAdd(P.presence + (long long)row * nc + col, 1u);     // the start point counts (movmodel.py:410-419)
                    }
                    have = true;
                }
            }
            if (have) {
                if (k >= W.kcap) {
                    // the phase is over for this track: its state goes to the next phase's input
                    const auto g = cg::coalesced_threads();
                    unsigned base = 0;
                    if (g.thread_rank() == 0) base = atomicAdd(W.out_count, (unsigned)g.size());
                    W.out_list[g.shfl(base, 0) + g.thread_rank()] = make_uint4(t, (unsigned)row | ((unsigned)col << 16), (unsigned)k, last);
                    have = false;
                } else if ((k & 3) == 0 && last != 4u && row >= 2 && row <= nr - 3 && col >= 1 && col <= nc - 3 && k + 4 <= kmax) {
                    tmode = true;
                    row = row * nc + col;
                    last = slot_of(last);
                } else {
                    // ---- one general step: tracks.cu's production step on the fields ----------------------------------
                    int r = row, c = col;
                    bool finish = k >= kmax;                                       // movmodel.py:285
                    if (!finish) {
                        if (k > P.burnin) finish = !(0 < r && r < nr - 1 && 0 < c && c < nc - 1);   // :287-289
                        else {                                                     // :290-291, :205-217
                            if (r <= 1) r += 2; else if (r >= nr - 2) r -= 2;
                            if (c <= 0) c += 2; else if (c >= nc - 2) c -= 2;
                        }
                    }
                    if (finish) {
                        if (P.traj_len != nullptr) P.traj_len[t] = k + 1;
                        steps_local += (unsigned long long)k;
                        have = false;
                    } else {
                        const int glin = r * nc + c;
                        const float2* base = P.fields + glin;
                        const unsigned long long gid = (unsigned long long)(P.track_id0 + (long long)t);
                        unsigned a, b, cc, dd;
                        philox4x32_10((unsigned)gid, (unsigned)(gid >> 32), (unsigned)k, 1u, P, a, b, cc, dd);
                        const double u = uniform52(a, b);
                        int idx;
                        if (last != 4u) {
                            const int4 cand = s_cand9[last];
                            const float2 fc = __ldg(base), f0 = __ldg(base + cand.x), f1 = __ldg(base + cand.y), f2 = __ldg(base + cand.z);
                            idx = choose_fast3<true, true>(P, base, nc, 0u, cand.w & 15, (cand.w >> 4) & 15, (cand.w >> 8) & 15,
                                                           fc, f0, f1, f2, u);
                        } else idx = choose_fast_general<true>(P, base, nc, 0x1EFu, u);
                        const int dr = ((idx * 11) >> 5) - 1, dc = idx - 3 * (dr + 1) - 1;
                        row = r + dr;                                              // :313-317
                        col = c + dc;
                        ++k;
                        last = (unsigned)idx;
                        atomicAdd(P.presence + (glin + dr * nc + dc), 1u);
                    }
                }
            }
        }
        if (done && !have) break;
        if (tmode) {
            const unsigned long long gid = (unsigned long long)(P.track_id0 + (long long)t);
            unsigned w[4];
            philox4x32_10((unsigned)gid, (unsigned)(gid >> 32), (unsigned)(k >> 2), 0u, P, w[0], w[1], w[2], w[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (tmode) {
                    const uint2 rec = __ldg(W.table + ((long long)row * 8 + last));
                    const unsigned r31 = w[j] >> 1;
                    if (rec.x <= W_ONE) {
                        const int4 cd = s_cand[last];
                        const bool a = r31 < rec.x, b = r31 < rec.y;
                        row += a ? cd.x : (b ? cd.y : cd.z);
                        last = ((unsigned)cd.w >> (a ? 0 : (b ? 4 : 8))) & 15u;
                        ++k;
                        red_add1(P.presence + row);
                    } else if (rec.x == W_UNMASKED) {
                        // a track heading away from track_direction in a potential minimum: all nine directional weights,
                        // mask dropped (movmodel.py:239-240) — first flat index whose cumulative threshold exceeds r
                        int cnt = 0;
#pragma unroll
                        for (int i = 0; i < 9; ++i) cnt += (r31 >= W.dthr[i]) ? 1 : 0;
                        const int idx = cnt < 8 ? cnt : 8;
                        const int dr = ((idx * 11) >> 5) - 1, dc = idx - 3 * (dr + 1) - 1;
                        row += dr * nc + dc;
                        last = slot_of((unsigned)idx);
                        ++k;
                        red_add1(P.presence + row);
                    } else {
                        tmode = false;                                             // border strip: back to (row, col) form
                        const int lin = row;
                        row = lin / nc; col = lin - row * nc;
                        last = flat_of(last);
                    }
                }
            }
            if (tmode && (k >= W.kcap || k + 4 > kmax)) {       // end of the phase, or too close to the step limit
                tmode = false;
                const int lin = row;
                row = lin / nc; col = lin - row * nc;
                last = flat_of(last);
            }
        }
    }
    if (P.total_steps != nullptr && steps_local) {
        const auto g = cg::coalesced_threads();
        const unsigned long long sum = cg::reduce(g, steps_local, cg::plus<unsigned long long>());
        if (g.thread_rank() == 0) atomicAdd(P.total_steps, sum);
    }
}

}  // namespace
}  // namespace ssrs

using namespace ssrs;

extern "C" int64_t ssrs_walk_table_bytes(int rows, int cols) {
    if (rows <= 0 || cols <= 0) return 0;
    return (int64_t)rows * cols * 8 * (int64_t)sizeof(uint2);
}

extern "C" int64_t ssrs_walk_workspace_bytes(int64_t n_tracks) {
    if (n_tracks < 0) return 0;
    return phase_workspace_bytes(n_tracks);
}

extern "C" int ssrs_transition_table(const float* fields, int rows, int cols, const double* dirprob9_host, void* table,
                                     void* stream) {
    SSRS_REQUIRE(fields != nullptr && table != nullptr && dirprob9_host != nullptr, "ssrs_transition_table: NULL argument");
    SSRS_REQUIRE(rows >= 5 && cols >= 5, "ssrs_transition_table: grid %dx%d too small", rows, cols);
    SSRS_REQUIRE((long long)rows * cols < 2147483647LL, "ssrs_transition_table: more than 2^31 cells");
    TableParams T;
    T.fields = reinterpret_cast<const float2*>(fields);
    T.table = reinterpret_cast<uint2*>(table);
    T.rows = rows; T.cols = cols;
    for (int i = 0; i < 9; ++i) T.dirp[i] = dirprob9_host[i];
    long long blocks = cdiv((long long)rows * cols, 256);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    transition_table_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(T);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int ssrs_walk_tracks(const void* table, const float* fields, int rows, int cols, const int32_t* start_rc,
                                int64_t n_tracks, int64_t track_id0, const double* dirprob9_host, uint64_t seed,
                                int32_t* traj_len, uint32_t* presence, unsigned long long* total_steps, void* workspace,
                                int64_t workspace_bytes, int first_phase_steps, void* stream) {
    SSRS_REQUIRE(table != nullptr && fields != nullptr && presence != nullptr, "ssrs_walk_tracks: table, fields and presence are required");
    SSRS_REQUIRE(rows >= 5 && cols >= 5 && rows <= 32767 && cols <= 32767, "ssrs_walk_tracks: grid %dx%d not supported", rows, cols);
    SSRS_REQUIRE((long long)rows * cols < 2147483647LL, "ssrs_walk_tracks: more than 2^31 cells");
    SSRS_REQUIRE(n_tracks >= 0 && n_tracks < 2147483647LL && track_id0 >= 0, "ssrs_walk_tracks: bad track count or id");
    SSRS_REQUIRE(start_rc != nullptr || n_tracks == 0, "ssrs_walk_tracks: start_rc is NULL");
    SSRS_REQUIRE(dirprob9_host != nullptr, "ssrs_walk_tracks: dirprob9_host is NULL");
    SSRS_REQUIRE(workspace != nullptr && workspace_bytes >= ssrs_walk_workspace_bytes(n_tracks),
                 "ssrs_walk_tracks: workspace smaller than ssrs_walk_workspace_bytes(n_tracks)");
    if (n_tracks == 0) return SSRS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    WalkParams W;
    TrackParams& P = W.tp;
    P.fields = reinterpret_cast<const float2*>(fields);
    P.start = reinterpret_cast<const int2*>(start_rc);
    P.uniforms = nullptr; P.traj = nullptr;
    P.traj_len = traj_len; P.presence = presence; P.total_steps = total_steps;
    P.in_head = nullptr; P.in_count = nullptr; P.in_list = nullptr; P.out_list = nullptr; P.out_count = nullptr; P.kcap = 0;
    P.n_tracks = n_tracks; P.track_id0 = track_id0; P.ustride = 0; P.traj_cap = 0;
    P.seed = seed;
    for (int i = 0; i < 9; ++i) P.dirp[i] = dirprob9_host[i];
    P.nu = 1.0; P.nu_is_one = 1;
    P.max_moves = (double)rows / 2 * (double)cols / 2;                       // movmodel.py:277
    {
        const double km = ceil(P.max_moves);
        P.kmax = km > 2147483647.0 ? 2147483647 : (int)km;
    }
    P.rows = rows; P.cols = cols;
    P.burnin = (int)((rows < cols ? rows : cols) / 10);                      // movmodel.py:276
    P.memory = 1;
    for (int i = 0; i < 10; ++i) {
        P.rk[2 * i] = (unsigned)seed + (unsigned)i * 0x9E3779B9u;
        P.rk[2 * i + 1] = (unsigned)(seed >> 32) + (unsigned)i * 0xBB67AE85u;
    }
    {   // unmasked directional fallback: cumulative thresholds of the nine weights, same order of additions as tracks.cu
        double run[9], acc = 0.0;
        int lastpos = -1;
        for (int i = 0; i < 9; ++i) { acc += P.dirp[i]; run[i] = acc; if (P.dirp[i] > 0.0) lastpos = i; }
        SSRS_REQUIRE(lastpos >= 0 && acc > 0.0, "ssrs_walk_tracks: directional weights are all zero");
        for (int i = 0; i < 9; ++i) {
            const double x = run[i] / run[8] * 2147483648.0 + 0.5;
            W.dthr[i] = (i >= lastpos || x >= 2147483648.0) ? 0x80000000u : (unsigned)x;
        }
    }
    W.table = reinterpret_cast<const uint2*>(table);
    uint4* lists[2] = {reinterpret_cast<uint4*>(workspace), reinterpret_cast<uint4*>(workspace) + n_tracks};
    unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<uint4*>(workspace) + 2 * n_tracks);   // [phase][head, count]
    int caps[kMaxPhases];
    const int n_phases = phase_caps(rows, cols, P.kmax, first_phase_steps, caps);
    SSRS_CUDA_TRY(cudaMemsetAsync(counters, 0, 2 * (kMaxPhases + 1) * sizeof(unsigned), st));
    int per_sm = 0;
    SSRS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, walk_kernel, 128, 0));
    if (per_sm < 1) per_sm = 1;
    long long blocks = cdiv(n_tracks, 128);
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    for (int p = 0; p < n_phases; ++p) {
        W.in_list = p == 0 ? nullptr : lists[(p - 1) & 1];
        W.out_list = lists[p & 1];
        W.in_head = counters + 2 * p;
        W.in_count = p == 0 ? nullptr : counters + 2 * p + 1;
        W.out_count = counters + 2 * (p + 1) + 1;
        W.kcap = caps[p];
        walk_kernel<<<(int)blocks, 128, 0, st>>>(W);
    }
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}
