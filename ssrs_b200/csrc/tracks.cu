// Stage 3+4 — batched stochastic track stepping with fused presence accumulation (sm_100a).
//
// Replaces generate_simulated_tracks (ssrs/movmodel.py:264-318, one Python call per track mapped over a
// fork pool at ssrs/simulator.py:360-369) and compute_presence_counts (ssrs/movmodel.py:410-419).
//
// One thread owns one track at a time and steps it to completion; when its track ends the lane takes the
// next unstarted track from a device queue (lane-level refill inside one flat loop, see the kernel), so
// warps stay as full as the heavy-tailed track lengths allow.  Per step a lane gathers the neighbours its
// direction-memory mask allows from the interleaved {updraft, potential} raster (one 8-byte read-only load
// each), evaluates the move probabilities — in the reference's exact arithmetic on request (float32
// potential differences, float64 everything else, numpy's pairwise-sum order), else in an equivalent
// division-free form —, draws one uniform (caller-supplied in verification mode, else counter-based
// Philox4x32-10 keyed by (seed, global track id, step)), and appends the new point: a coalesced step-major
// int16x2 store (optional) and one `red.global.add.u32` on the presence raster.
//
// This translation unit is compiled with -fmad=false: float64 products and sums must round separately
// to reproduce numpy bit for bit.
#include "stepper.cuh"

#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

#include <mutex>

namespace cg = cooperative_groups;

namespace ssrs {
namespace {

// Fast-lane tables, one row per previous move.  Rows are indexed by the move's "slot" = flat index with the centre
// (4) squeezed out, 0..7, so that eight rows of 16 bytes cover the 32 shared-memory banks exactly once: lanes of a warp
// that hold different previous moves read their rows without bank conflicts.
struct FastLut {
    int4 cand[8];       // element offsets of the three candidates from the centre cell; .w = their slots, 4 bits each
    float4 ninv[8];     // float32(1/sqrt 2) for diagonal candidates, else 1 (movmodel.py:139-141)
    double2 dir01[8];   // directional weights of candidates 0 and 1 (the fallback of movmodel.py:234-236)
    double dir2[8];     // ... and of candidate 2
    double drun[9];     // running sums of the nine directional weights in flat order (the unmasked fallback, :239-240)
    int dlast;          // last flat index with a positive directional weight (4 if none)
};

// One fast-lane step: choose_fast3<true, true> with the selection made branch-free and the new centre's fields taken
// from the chosen candidate's registers instead of a fourth gather.
//   * max.NaN(d_i, 0) replaces `d_i > 0 ? ... : 0`: the product is the same for d_i > 0, a zero for d_i <= 0 and NaN
//     for a NaN difference (q_i == 0 exactly when d_i <= 0, because the clipped updrafts are >= 1e-6);
//   * d_i NaN or no d_i > 0  <=>  choose_fast3's (any_nan || all q == 0): then c2 is NaN or zero, `c2 > target` fails,
//     and the directional weights are handled off the main line (11 % of the steps; a lone lane — the end of every
//     launch — saves their two table loads and six selects on the other 89 %);
//   * when c2 > target the chosen move is the first running sum above the target — the same three comparisons; every
//     other case (weights all zero after the fallback, u * c2 rounding up to c2, a NaN or infinite updraft) goes to
//     the out-of-line copy of the original code.  Same draws, same arithmetic: bit-identical trajectories.
__device__ __forceinline__ void fast_step(const TrackParams& P, const FastLut& lut, int nc, int& lin, unsigned& slot,
                                          float2& fc, double& uc, double u) {
    const int4 cand = lut.cand[slot];
    const float4 nv = lut.ninv[slot];
    // 32-bit cell indices (rows * cols < 2^31): one IMAD.WIDE per address
    const float2 f0 = __ldg(P.fields + (lin + cand.x)), f1 = __ldg(P.fields + (lin + cand.y)), f2 = __ldg(P.fields + (lin + cand.z));
    const float d0 = __fmul_rn(__fsub_rn(fc.y, f0.y), nv.x);            // float32, movmodel.py:301-304
    const float d1 = __fmul_rn(__fsub_rn(fc.y, f1.y), nv.y);
    const float d2 = __fmul_rn(__fsub_rn(fc.y, f2.y), nv.z);
    const double u0 = clip_updraft(f0.x), u1 = clip_updraft(f1.x), u2 = clip_updraft(f2.x);
    const double s0 = uc + u0, s1 = uc + u1, s2 = uc + u2;
    // max.NaN: a NaN potential difference makes its weight, and with it c2, NaN and sends the step to the else branch
    const double q0 = ((double)fmax_nan(d0, 0.0f) * u0) * (s1 * s2);
    const double q1 = ((double)fmax_nan(d1, 0.0f) * u1) * (s0 * s2);
    const double q2 = ((double)fmax_nan(d2, 0.0f) * u2) * (s0 * s1);
    const double c1 = q0 + q1, c2 = c1 + q2;
    const double target = u * c2;
    if (c2 > target) {
        const bool a = q0 > target, b = c1 > target;
        lin += a ? cand.x : (b ? cand.y : cand.z);
        slot = ((unsigned)cand.w >> (a ? 0 : (b ? 4 : 8))) & 15u;
        fc = a ? f0 : (b ? f1 : f2);
        uc = a ? u0 : (b ? u1 : u2);
    } else {
        // c2 is zero (no candidate lies lower: 11 % of all steps), NaN, or u * c2 rounded up to c2
        int idx = -1;
        if (!(fmax_nan(fmax_nan(d0, d1), d2) > 0.0f)) {
            // choose_fast3's (any_nan || all q == 0): the candidates' directional weights (movmodel.py:234-236)
            const double2 dir01 = lut.dir01[slot];
            const double dir2 = lut.dir2[slot];
            const double e1 = dir01.x + dir01.y, e2 = e1 + dir2;
            const double tg = u * e2;
            if (e2 > tg) {
                const bool a = dir01.x > tg, b = e1 > tg;
                lin += a ? cand.x : (b ? cand.y : cand.z);
                slot = ((unsigned)cand.w >> (a ? 0 : (b ? 4 : 8))) & 15u;
                fc = a ? f0 : (b ? f1 : f2);
                uc = a ? u0 : (b ? u1 : u2);
            } else if (e2 == 0.0) {
                // those are zero as well (a track heading away from track_direction in a potential minimum, ~1 % of all
                // steps): the reference drops the mask and draws from all nine directional weights (movmodel.py:239-240).
                // That distribution does not depend on the cell: first flat index whose running sum exceeds u * total,
                // exactly what choose_fast_general<false>(mask 0) evaluates.
                const double tg9 = u * lut.drun[8];
                int cnt = 0;
#pragma unroll
                for (int i = 0; i < 9; ++i) cnt += (lut.drun[i] > tg9) ? 0 : 1;   // running sums are non-decreasing
                idx = cnt < 9 ? cnt : lut.dlast;
            } else idx = -2;
        } else idx = -2;
        if (idx == -2)
            idx = choose_fast3_rare(P, P.fields + lin, nc, (int)flat_of(cand.w & 15), (int)flat_of((cand.w >> 4) & 15),
                                    (int)flat_of((cand.w >> 8) & 15), fc, f0, f1, f2, u);
        if (idx >= 0) {
            const int dr = ((idx * 11) >> 5) - 1, dc = idx - 3 * (dr + 1) - 1;
            lin += dr * nc + dc;
            slot = slot_of((unsigned)idx);
            fc = __ldg(P.fields + lin);
            uc = clip_updraft(fc.x);
        }
    }
#ifndef SSRS_X_NORED                   // (timing experiment: profiles/r02_step_experiments.txt)
    red_add1(P.presence + lin);
#endif
}

// MEM1: track_dirn_restrict == 1 (the default): the mask is exactly the three candidates of the last move, so
// no history register, no mask arithmetic.
// 5 CTAs of 128 threads per SM (96 registers, nothing spilled): measured 2 % faster in the ring than 6 (80 registers,
// 92 bytes spilled) and than 4 (profiles/r02_step_experiments.txt) — the kernel is not occupancy-bound.  The denser
// instantiations (6, 7, 8 CTAs: 80, 72, 64 registers) are kept for the FIRST phase of launches whose tracks fit that many
// resident CTAs per SM but not 5 (100k tracks on 148 SMs need 6, 125k need 7): all tracks of a first phase take the same
// number of steps, so a 5-CTA grid would step the excess in a second, mostly empty round.
#ifndef SSRS_STEP_MINB
#define SSRS_STEP_MINB 5
#endif
#ifndef SSRS_STEP_HYBRID
#define SSRS_STEP_HYBRID 1
#endif
#ifndef SSRS_STEP_HYBRID_MAX
#define SSRS_STEP_HYBRID_MAX 3          // denser instantiations tried for the first phase: 6, 7, 8 CTAs per SM
#endif
template <bool HAS_FIELDS, bool EXACT, bool MEM1, int MINB = SSRS_STEP_MINB>
__global__ void __launch_bounds__(128, MINB) step_tracks_kernel(const TrackParams P) {
    const unsigned n_in = P.in_count != nullptr ? *P.in_count : (unsigned)P.n_tracks;
    if ((unsigned)(blockIdx.x * blockDim.x) >= n_in) return;       // surplus CTA of a late phase: nothing to step
    // per previous move: element offsets of its three candidates, their packed indices, distance factors and
    // directional weights
    __shared__ int4 s_cand[9];                 // general step: indexed by the flat move index, .w = flat indices
    __shared__ FastLut s_lut;                  // fast lane: indexed by slot
    if (threadIdx.x < 9) {
        const unsigned last = threadIdx.x;
        const unsigned c3 = (unsigned)((last < 5 ? (C3_A >> (12 * last)) : (C3_B >> (12 * (last - 5)))) & 0xFFFu);
        const int i0 = c3 & 15, i1 = (c3 >> 4) & 15, i2 = (c3 >> 8) & 15;
        const int4 off = make_int4((i0 / 3 - 1) * P.cols + (i0 % 3 - 1), (i1 / 3 - 1) * P.cols + (i1 % 3 - 1),
                                   (i2 / 3 - 1) * P.cols + (i2 % 3 - 1), (int)c3);
        s_cand[last] = off;
        if (last != 4u) {
            const unsigned sl = slot_of(last);
            s_lut.cand[sl] = make_int4(off.x, off.y, off.z, (int)(slot_of(i0) | (slot_of(i1) << 4) | (slot_of(i2) << 8)));
            // even flat index (0,2,6,8) = diagonal move
            s_lut.ninv[sl] = make_float4((i0 & 1) ? 1.0f : 0.70710677f, (i1 & 1) ? 1.0f : 0.70710677f,
                                         (i2 & 1) ? 1.0f : 0.70710677f, 0.f);
            s_lut.dir01[sl] = make_double2(P.dirp[i0], P.dirp[i1]);
            s_lut.dir2[sl] = P.dirp[i2];
        } else {
            double run = 0.0;
            int lastpos = 4;
            for (int i = 0; i < 9; ++i) {                       // same order of additions as choose_fast_general
                run += P.dirp[i];
                s_lut.drun[i] = run;
                if (P.dirp[i] > 0.0) lastpos = i;
            }
            s_lut.dlast = lastpos;
        }
    }
    __syncthreads();
    long long t = 0;
    const int nr = P.rows, nc = P.cols;

    const int kmax = P.kmax;
    const int kstop = min(kmax, P.kcap);                           // fast-lane budgets end at the phase cap
    unsigned long long steps_local = 0;
    bool alive = false;
    int row = 0, col = 0, k = 0;
    unsigned last = 4;                    // flat index of the previous move (4 = none yet)
    unsigned long long hist = 4;          // 4-bit move codes, most recent in the low nibble (only if !MEM1)
    int hcount = 1;
    unsigned run_mask = 0x1EF;            // AND over the whole history (memory == 0)
    unsigned rng_c = 0, rng_d = 0;        // second half of the last Philox block
    // Fast lane: pairs of ordinary steps (interior cell, previous move known, nu == 1, Philox stream, counts only, no
    // trajectory store) without the general step's bookkeeping; one Philox block feeds both steps of a pair, so there is
    // no parity branch.  The fast lane carries the linear cell index only: a track whose distance to the nearest
    // non-interior cell is m cannot leave the interior in m steps, so it is granted floor((m + 1) / 2) pairs ("budget")
    // without looking at the border and then measures the distance again.  Same arithmetic, same draws.
    const bool fast_lane = HAS_FIELDS && !EXACT && MEM1 && P.nu_is_one && P.uniforms == nullptr && P.traj == nullptr &&
                           P.presence != nullptr;
    // While a lane has budget its state lives in the same registers in fast-lane form: `row` holds the linear cell index,
    // `k` the pair counter k / 2, `last` the previous move's slot (registers: 96 at 5 CTAs/SM).
    bool done = false, in_fast = false;
    int budget = 0;
    float2 fcen = make_float2(0.f, 0.f);
    double ucen = 0.0;

    // One flat loop, two blocks per iteration: the attention block (a lane without budget: refill, budget refresh, or
    // one general step) and the pair block (lanes with budget).  Both reconverge inside the iteration, so a lane whose
    // track has ended waits at most one block for its warp before it starts the next track — with an inner loop
    // around the pairs the convergence barrier would hold it until every lane of the warp had left that loop, i.e.
    // until the warp's longest track ended.
    while (true) {
        if (budget == 0) {
            if (in_fast) {                                                  // leave or refresh: canonical state back
                const int lin = row;
                row = lin / nc; col = lin - row * nc;
                k *= 2;
                last = flat_of(last);
            } else if (!alive) {
                // next entry of the launch's input, first come, first served: track lengths are heavy-tailed (mean 1e4
                // steps, maximum above 1e5), so a fixed list per thread would leave most lanes idle while a few work through
                // long lists; results do not depend on which lane steps which track (the random stream is keyed by the
                // track id, the counts are integer sums).  Lanes that ask together share one atomic.
                const auto g = cg::coalesced_threads();
                unsigned first = 0;
                if (g.thread_rank() == 0) first = atomicAdd(P.in_head, (unsigned)g.size());
                const unsigned idx = g.shfl(first, 0) + g.thread_rank();
                if (idx >= n_in) done = true;
                else if (P.in_list != nullptr) {
                    const uint4 e = P.in_list[idx];
                    t = e.x; row = (int)(e.y & 0xFFFFu); col = (int)(e.y >> 16); k = (int)e.z; last = e.w;
                    alive = true;
                } else {
                    t = idx;
                    int2 s = __ldg(P.start + t);
                    row = s.x; col = s.y; k = 0; last = 4; hist = 4; hcount = 1; run_mask = 0x1EF;
                    if (P.traj != nullptr && P.traj_cap > 0) P.traj[t] = make_short2((short)row, (short)col);
                    if (P.presence != nullptr) atomicAdd(P.presence + (long long)row * nc + col, 1u);
                    alive = true;
                }
            }
            if (alive && k >= P.kcap) {
                // the phase is over for this track (phased launches only, MEM1): its state goes to the next phase's input
                const auto g = cg::coalesced_threads();
                unsigned first = 0;
                if (g.thread_rank() == 0) first = atomicAdd(P.out_count, (unsigned)g.size());
                P.out_list[g.shfl(first, 0) + g.thread_rank()] = make_uint4((unsigned)t, (unsigned)row | ((unsigned)col << 16), (unsigned)k, last);
                alive = false;
                in_fast = false;
            } else if (alive) {
                if (fast_lane && last != 4u && (k & 1) == 0) {
                    const int m = min(min(row - 2, nr - 3 - row), min(col - 1, nc - 3 - col));
                    budget = max(min((m + 1) >> 1, (kstop - k) >> 1), 0);
                }
                if (budget > 0) {
                    row = row * nc + col;
                    last = slot_of(last);
                    k >>= 1;
                    if (!in_fast) {
                        fcen = __ldg(P.fields + row);
                        ucen = clip_updraft(fcen.x);
                        in_fast = true;
                    }
                } else {
                    in_fast = false;
                    // ---- one general step -------------------------------------------------------------------
                    int r = row, c = col;
                    bool finish = k >= kmax;                                    // movmodel.py:285
                    if (!finish) {
                        if (k > P.burnin) {                                     // :287-289
                            finish = !(0 < r && r < nr - 1 && 0 < c && c < nc - 1);
                        } else {                                                // :290-291, :205-217
                            if (r <= 1) r += 2; else if (r >= nr - 2) r -= 2;
                            if (c <= 0) c += 2; else if (c >= nc - 2) c -= 2;
                        }
                    }
                    if (finish) {
                        if (P.traj_len != nullptr) P.traj_len[t] = k + 1;
                        steps_local += (unsigned long long)k;
                        alive = false;
                    } else {
                        const int glin = r * nc + c;                            // rows * cols < 2^31 (checked on the host)
                        const float2* base = HAS_FIELDS ? P.fields + glin : nullptr;
                        const bool three = !EXACT && last != 4u && P.nu_is_one;
                        // issue the gathers first so they overlap the random-number rounds
                        int4 cand = make_int4(0, 0, 0, 0);
                        float2 fc = make_float2(0.f, 0.f), f0 = fc, f1 = fc, f2 = fc;
                        if (three) {
                            cand = s_cand[last];
                            if (HAS_FIELDS) {
                                fc = __ldg(base);
                                f0 = __ldg(base + cand.x);
                                f1 = __ldg(base + cand.y);
                                f2 = __ldg(base + cand.z);
                            }
                        }
                        // direction-memory mask (:307-309)
                        unsigned mask;
                        if (MEM1) mask = (three ? 0u : restrict_mask(last));
                        else if (P.memory == 0) mask = run_mask;
                        else {
                            mask = 0x1EF;
                            int m = P.memory < hcount ? P.memory : hcount;
                            for (int j = 0; j < m; ++j) mask &= restrict_mask((unsigned)((hist >> (4 * j)) & 15));
                        }
                        // one uniform per step (:312)
                        double u;
                        if (P.uniforms != nullptr) {
                            // verification mode: a track that outlives the caller's stream stops here and reports a
                            // negative length (-(points so far)); the host retries with a longer stream
                            if ((long long)k >= P.ustride) {
                                if (P.traj_len != nullptr) P.traj_len[t] = -(k + 1);
                                steps_local += (unsigned long long)k;
                                alive = false;
                                continue;
                            }
                            u = __ldg(P.uniforms + t * P.ustride + k);
                        } else {
                            // Philox4x32-10 yields four words = two uniforms: counter (gid, k >> 1), words {0,1} for even k,
                            // {2,3} for odd k
                            if ((k & 1) == 0) {
                                const unsigned long long gid = (unsigned long long)(P.track_id0 + t);
                                unsigned a, b;
                                philox4x32_10((unsigned)gid, (unsigned)(gid >> 32), (unsigned)(k >> 1), 0u, P, a, b, rng_c, rng_d);
                                u = uniform52(a, b);
                            } else {
                                u = uniform52(rng_c, rng_d);
                            }
                        }
                        int idx;
                        if (EXACT) idx = choose_exact<HAS_FIELDS>(P, base, nc, mask, u);
                        else if (three) {
                            const int i0 = cand.w & 15, i1 = (cand.w >> 4) & 15, i2 = (cand.w >> 8) & 15;
                            idx = choose_fast3<HAS_FIELDS, MEM1>(P, base, nc, mask, i0, i1, i2, fc, f0, f1, f2, u);
                        } else idx = choose_fast_general<HAS_FIELDS>(P, base, nc, mask, u);
                        const int dr = ((idx * 11) >> 5) - 1;                   // idx / 3 - 1 for idx in 0..8
                        const int dc = idx - 3 * (dr + 1) - 1;
                        row = r + dr;                                           // :313-317
                        col = c + dc;
                        ++k;
                        last = (unsigned)idx;
                        if (!MEM1) {
                            hist = (hist << 4) | (unsigned long long)idx;
                            hcount = hcount < 16 ? hcount + 1 : 16;
                            run_mask &= restrict_mask((unsigned)idx);
                        }
                        if (P.traj != nullptr && (long long)k < P.traj_cap)
                            P.traj[(long long)k * P.n_tracks + t] = make_short2((short)row, (short)col);
                        if (P.presence != nullptr) atomicAdd(P.presence + (glin + dr * nc + dc), 1u);
                    }
                }
            }
        }
        if (done && !alive) break;
        if (budget > 0) {
            const unsigned long long gid = (unsigned long long)(P.track_id0 + t);
            unsigned a, b, cc, dd;
            philox4x32_10((unsigned)gid, (unsigned)(gid >> 32), (unsigned)k, 0u, P, a, b, cc, dd);
            fast_step(P, s_lut, nc, row, last, fcen, ucen, uniform52(a, b));
            fast_step(P, s_lut, nc, row, last, fcen, ucen, uniform52(cc, dd));
            ++k; --budget;
        }
    }
    if (P.total_steps != nullptr && steps_local) {
        // lanes leave the loop at different times: one atomic per group of lanes that arrive together
        const auto g = cg::coalesced_threads();
        const unsigned long long sum = cg::reduce(g, steps_local, cg::plus<unsigned long long>());
        if (g.thread_rank() == 0) atomicAdd(P.total_steps, sum);
    }
}

__global__ void interleave_kernel(const float* __restrict__ u, const float* __restrict__ p, float2* __restrict__ out,
                                  long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = make_float2(__ldcs(u + i), __ldcs(p + i));
}

__global__ void presence_from_traj_kernel(const short2* __restrict__ traj, long long traj_cap,
                                          const int* __restrict__ len, long long n_tracks, int rows, int cols,
                                          unsigned* presence) {
    const long long total = traj_cap * n_tracks;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const long long k = i / n_tracks, t = i - k * n_tracks;
        if (k < (long long)__ldg(len + t)) {
            const short2 pt = traj[i];
            if (pt.x >= 0 && pt.x < rows && pt.y >= 0 && pt.y < cols)
                atomicAdd(presence + (long long)pt.x * cols + pt.y, 1u);
        }
    }
}

// step-major trajectories -> packed points: point k of track t goes to points[offsets[t] + k]
__global__ void pack_traj_kernel(const short2* __restrict__ traj, long long traj_cap, const int* __restrict__ len,
                                 const long long* __restrict__ offsets, long long n_tracks, short2* __restrict__ points) {
    const long long total = traj_cap * n_tracks;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const long long k = i / n_tracks, t = i - k * n_tracks;
        if (k < (long long)__ldg(len + t)) points[__ldg(offsets + t) + k] = traj[i];
    }
}

}  // namespace
}  // namespace ssrs

using namespace ssrs;

namespace {
// Track-queue counters: one 8-byte slot per launch, taken round-robin from a small per-device array and zeroed on the
// launch's stream.  A slot is reused 1024 launches later, long after its launch has drained.
constexpr int kQueueSlots = 1024;
constexpr int kMaxDevices = 64;
unsigned long long* g_queue[kMaxDevices] = {nullptr};
unsigned g_queue_next[kMaxDevices] = {0};
std::mutex g_queue_mutex;

int queue_slot(unsigned long long** slot) {
    int dev = 0;
    SSRS_CUDA_TRY(cudaGetDevice(&dev));
    SSRS_REQUIRE(dev >= 0 && dev < kMaxDevices, "ssrs_step_tracks: device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_queue_mutex);
    if (g_queue[dev] == nullptr) SSRS_CUDA_TRY(cudaMalloc(&g_queue[dev], kQueueSlots * sizeof(unsigned long long)));
    *slot = g_queue[dev] + (g_queue_next[dev]++ % kQueueSlots);
    return SSRS_OK;
}
}  // namespace

namespace {
int step_tracks_impl(const float* fields, int rows, int cols, const int32_t* start_rc, int64_t n_tracks,
                     int64_t track_id0, const double* dirprob9_host, int memory, double nu, uint64_t seed,
                     const double* uniforms, int64_t uniforms_stride, int16_t* traj, int64_t traj_cap,
                     int32_t* traj_len, uint32_t* presence, unsigned long long* total_steps, int flags,
                     void* workspace, int64_t workspace_bytes, int first_phase_steps, void* stream) {
    SSRS_REQUIRE(rows >= 5 && cols >= 5, "ssrs_step_tracks: grid %dx%d too small", rows, cols);
    SSRS_REQUIRE(rows <= 32767 && cols <= 32767, "ssrs_step_tracks: int16 trajectories need rows, cols <= 32767");
    SSRS_REQUIRE((long long)rows * cols < 2147483647LL, "ssrs_step_tracks: more than 2^31 cells");
    SSRS_REQUIRE(n_tracks >= 0 && n_tracks < 2147483647LL && track_id0 >= 0, "ssrs_step_tracks: bad track count or id");
    SSRS_REQUIRE(start_rc != nullptr || n_tracks == 0, "ssrs_step_tracks: start_rc is NULL");
    SSRS_REQUIRE(dirprob9_host != nullptr, "ssrs_step_tracks: dirprob9_host is NULL");
    SSRS_REQUIRE(uniforms == nullptr || uniforms_stride > 0, "ssrs_step_tracks: uniforms_stride must be positive");
    SSRS_REQUIRE(traj == nullptr || traj_cap > 0, "ssrs_step_tracks: traj given with traj_cap <= 0");
    SSRS_REQUIRE(workspace == nullptr || workspace_bytes >= phase_workspace_bytes(n_tracks),
                 "ssrs_step_tracks_phased: workspace smaller than ssrs_walk_workspace_bytes(n_tracks)");
    if (memory < 0 || memory > 16) {
        set_error("ssrs_step_tracks: track_dirn_restrict=%d outside the supported range 0..16", memory);
        return SSRS_ERR_UNSUPPORTED;
    }
    if (n_tracks == 0) return SSRS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TrackParams P;
    P.fields = reinterpret_cast<const float2*>(fields);
    P.start = reinterpret_cast<const int2*>(start_rc);
    P.uniforms = uniforms;
    P.traj = reinterpret_cast<short2*>(traj);
    P.traj_len = traj_len;
    P.presence = presence;
    P.total_steps = total_steps;
    P.n_tracks = n_tracks; P.track_id0 = track_id0; P.ustride = uniforms_stride; P.traj_cap = traj ? traj_cap : 0;
    P.seed = seed;
    for (int i = 0; i < 9; ++i) P.dirp[i] = dirprob9_host[i];
    P.nu = nu; P.nu_is_one = (nu == 1.0);
    P.max_moves = (double)rows / 2 * (double)cols / 2;                       // movmodel.py:277
    {   // `k < max_moves` with integer k  <=>  k < ceil(max_moves)
        const double km = ceil(P.max_moves);
        P.kmax = km > 2147483647.0 ? 2147483647 : (int)km;
    }
    P.rows = rows; P.cols = cols;
    P.burnin = (int)((rows < cols ? rows : cols) / 10);                      // movmodel.py:276
    P.memory = memory;
    for (int i = 0; i < 10; ++i) {
        P.rk[2 * i] = (unsigned)seed + (unsigned)i * 0x9E3779B9u;
        P.rk[2 * i + 1] = (unsigned)(seed >> 32) + (unsigned)i * 0xBB67AE85u;
    }
    // verification mode always uses the exact (numpy bit-for-bit) arithmetic
    const bool exact = (uniforms != nullptr) || (flags & SSRS_STEP_EXACT);
    const int threads = 128;
    const bool mem1 = (memory == 1);
    void (*kern)(const TrackParams);
    if (fields != nullptr) {
        if (exact) kern = mem1 ? step_tracks_kernel<true, true, true> : step_tracks_kernel<true, true, false>;
        else kern = mem1 ? step_tracks_kernel<true, false, true> : step_tracks_kernel<true, false, false>;
    } else {
        if (exact) kern = mem1 ? step_tracks_kernel<false, true, true> : step_tracks_kernel<false, true, false>;
        else kern = mem1 ? step_tracks_kernel<false, false, true> : step_tracks_kernel<false, false, false>;
    }
    // as many tracks resident at once as the registers allow: the kernel is latency-bound and its duration is
    // set by the longest track, so every track should start at time zero
    int per_sm = 0;
    SSRS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
    if (per_sm < 1) per_sm = 1;
    long long blocks = cdiv(n_tracks, threads);
    long long cap = (long long)sm_count() * per_sm;
    void (*kern0)(const TrackParams) = kern;              // first phase (or the only launch): every track is in it
    long long blocks0 = blocks < cap ? blocks : cap;
    if (SSRS_STEP_HYBRID && blocks > cap && fields != nullptr && !exact && mem1) {
        // the smallest denser instantiation that holds every track at once, if there is one: one round instead of two
        void (*dense[3])(const TrackParams) = {step_tracks_kernel<true, false, true, SSRS_STEP_MINB + 1>,
                                               step_tracks_kernel<true, false, true, SSRS_STEP_MINB + 2>,
                                               step_tracks_kernel<true, false, true, SSRS_STEP_MINB + 3>};
        for (int j = 0; j < SSRS_STEP_HYBRID_MAX; ++j) {
            int occ = 0;
            SSRS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dense[j], threads, 0));
            if (blocks <= (long long)sm_count() * occ) { kern0 = dense[j]; blocks0 = blocks; break; }
        }
    }
    if (blocks > cap) blocks = cap;
    P.in_count = nullptr; P.in_list = nullptr; P.out_list = nullptr; P.out_count = nullptr; P.kcap = 2147483647;
    if (workspace == nullptr || !mem1) {
        // one launch steps every track to its end (a saved state holds one previous move only: memory 1)
        unsigned long long* slot = nullptr;
        const int rc = queue_slot(&slot);
        if (rc != SSRS_OK) return rc;
        P.in_head = reinterpret_cast<unsigned*>(slot);
        SSRS_CUDA_TRY(cudaMemsetAsync(slot, 0, sizeof(unsigned long long), st));
        kern0<<<(int)blocks0, threads, 0, st>>>(P);
        SSRS_CUDA_TRY(cudaGetLastError());
        return SSRS_OK;
    }
    // phased launch: survivors of each phase are compacted into the next phase's input (see ssrs_step_tracks_phased)
    uint4* lists[2] = {reinterpret_cast<uint4*>(workspace), reinterpret_cast<uint4*>(workspace) + n_tracks};
    unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<uint4*>(workspace) + 2 * n_tracks);   // [phase][head, count]
    int caps[kMaxPhases];
    const int n_phases = phase_caps(rows, cols, P.kmax, first_phase_steps, caps);
    SSRS_CUDA_TRY(cudaMemsetAsync(counters, 0, 2 * (kMaxPhases + 1) * sizeof(unsigned), st));
    for (int p = 0; p < n_phases; ++p) {
        P.in_list = p == 0 ? nullptr : lists[(p - 1) & 1];
        P.out_list = lists[p & 1];
        P.in_head = counters + 2 * p;
        P.in_count = p == 0 ? nullptr : counters + 2 * p + 1;
        P.out_count = counters + 2 * (p + 1) + 1;
        P.kcap = caps[p];
        if (p == 0) kern0<<<(int)blocks0, threads, 0, st>>>(P);
        else kern<<<(int)blocks, threads, 0, st>>>(P);
    }
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}
}  // namespace

extern "C" int ssrs_step_phase_count(int rows, int cols, int first_phase_steps) {
    // kernel launches of one phased stepping call (host arithmetic only; bench.py counts its launches with it)
    if (rows < 5 || cols < 5) return 0;
    const double km = ceil((double)rows / 2 * (double)cols / 2);
    int caps[kMaxPhases];
    return phase_caps(rows, cols, km > 2147483647.0 ? 2147483647 : (int)km, first_phase_steps, caps);
}

extern "C" int ssrs_step_tracks(const float* fields, int rows, int cols, const int32_t* start_rc, int64_t n_tracks,
                                int64_t track_id0, const double* dirprob9_host, int memory, double nu, uint64_t seed,
                                const double* uniforms, int64_t uniforms_stride, int16_t* traj, int64_t traj_cap,
                                int32_t* traj_len, uint32_t* presence, unsigned long long* total_steps,
                                int flags, void* stream) {
    return step_tracks_impl(fields, rows, cols, start_rc, n_tracks, track_id0, dirprob9_host, memory, nu, seed, uniforms,
                            uniforms_stride, traj, traj_cap, traj_len, presence, total_steps, flags, nullptr, 0, 0, stream);
}

extern "C" int ssrs_step_tracks_phased(const float* fields, int rows, int cols, const int32_t* start_rc, int64_t n_tracks,
                                       int64_t track_id0, const double* dirprob9_host, int memory, double nu, uint64_t seed,
                                       const double* uniforms, int64_t uniforms_stride, int16_t* traj, int64_t traj_cap,
                                       int32_t* traj_len, uint32_t* presence, unsigned long long* total_steps,
                                       int flags, void* workspace, int64_t workspace_bytes, int first_phase_steps,
                                       void* stream) {
    SSRS_REQUIRE(workspace != nullptr, "ssrs_step_tracks_phased: workspace is NULL");
    return step_tracks_impl(fields, rows, cols, start_rc, n_tracks, track_id0, dirprob9_host, memory, nu, seed, uniforms,
                            uniforms_stride, traj, traj_cap, traj_len, presence, total_steps, flags, workspace, workspace_bytes,
                            first_phase_steps, stream);
}

extern "C" int ssrs_interleave_fields(const float* updraft, const float* potential, float* fields, int64_t n,
                                      void* stream) {
    SSRS_REQUIRE(updraft && potential && fields, "ssrs_interleave_fields: NULL raster");
    SSRS_REQUIRE(n >= 0, "ssrs_interleave_fields: negative size");
    if (n == 0) return SSRS_OK;
    long long blocks = cdiv(n, 256);
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    interleave_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        updraft, potential, reinterpret_cast<float2*>(fields), n);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int ssrs_presence_counts(const int16_t* traj, int64_t traj_cap, const int32_t* traj_len, int64_t n_tracks,
                                    int rows, int cols, uint32_t* presence, void* stream) {
    SSRS_REQUIRE(traj && traj_len && presence, "ssrs_presence_counts: NULL buffer");
    SSRS_REQUIRE(traj_cap > 0 && n_tracks >= 0 && rows > 0 && cols > 0, "ssrs_presence_counts: bad sizes");
    if (n_tracks == 0) return SSRS_OK;
    long long blocks = cdiv(traj_cap * n_tracks, 256);
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    presence_from_traj_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const short2*>(traj), traj_cap, traj_len, n_tracks, rows, cols, presence);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int ssrs_pack_trajectories(const int16_t* traj, int64_t traj_cap, const int32_t* traj_len, const int64_t* offsets,
                                      int64_t n_tracks, int16_t* points, void* stream) {
    SSRS_REQUIRE(traj && traj_len && offsets && points, "ssrs_pack_trajectories: NULL buffer");
    SSRS_REQUIRE(traj_cap > 0 && n_tracks >= 0, "ssrs_pack_trajectories: bad sizes");
    if (n_tracks == 0) return SSRS_OK;
    long long blocks = cdiv(traj_cap * n_tracks, 256);
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    pack_traj_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const short2*>(traj), traj_cap, traj_len, reinterpret_cast<const long long*>(offsets), n_tracks,
        reinterpret_cast<short2*>(points));
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}
