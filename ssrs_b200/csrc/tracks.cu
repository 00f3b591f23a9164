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
#include "common.cuh"

#include <math.h>

#include <mutex>

namespace ssrs {
namespace {

// direction-memory masks, get_track_restrictions (movmodel.py:185-202) as a table; bit i = flat move
// index 3*(dr+1)+(dc+1).  Previous move SW,S,SE,W,(0,0),E,NW,N,NE:
//   0x00B 0x007 0x026 0x049 0x1EF 0x124 0x0C8 0x1C0 0x1A0     packed 9 bits each into two words.
constexpr unsigned long long LUT_A = (0x00BULL) | (0x007ULL << 9) | (0x026ULL << 18) | (0x049ULL << 27) |
                                     (0x1EFULL << 36) | (0x124ULL << 45) | (0x0C8ULL << 54);
constexpr unsigned long long LUT_B = (0x1C0ULL) | (0x1A0ULL << 9);

__device__ __forceinline__ unsigned restrict_mask(unsigned move) {
    unsigned long long w = move < 7 ? (LUT_A >> (9 * move)) : (LUT_B >> (9 * (move - 7)));
    return (unsigned)w & 0x1FFu;
}

struct TrackParams {
    const float2* fields;
    const int2* start;
    const double* uniforms;
    short2* traj;
    int* traj_len;
    unsigned* presence;
    unsigned long long* total_steps;
    unsigned long long* next_track;       // device counter, zero at launch: tracks beyond the first per thread are drawn from it
    long long n_tracks, track_id0, ustride, traj_cap;
    unsigned long long seed;
    double dirp[9];
    double nu;
    double max_moves;
    int rows, cols, burnin, memory, nu_is_one, kmax;
    unsigned rk[20];                      // Philox round keys (seed + i * Weyl constants), formed once on the host
};

// Philox4x32-10 (Salmon et al. 2011), counter = (track_lo, track_hi, step_lo, step_hi), key = seed.
// The ten round keys (k0 + i * 0x9E3779B9, k1 + i * 0xBB67AE85) are the same for every block of a launch: they sit in
// the kernel parameters, where the xor reads them as constant-bank operands (no key-schedule instructions).
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, const TrackParams& P,
                                              unsigned& o0, unsigned& o1, unsigned& o2, unsigned& o3) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned n0 = hi1 ^ c1 ^ P.rk[2 * i], n1 = lo1, n2 = hi0 ^ c3 ^ P.rk[2 * i + 1], n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

__device__ __forceinline__ double uniform52(unsigned a, unsigned b) {
    // 52 random mantissa bits under exponent 0 give [1,2); subtract 1 -> [0,1) on a 2^-52 lattice
    const unsigned long long bits = 0x3FF0000000000000ULL | ((unsigned long long)a << 20) | (unsigned long long)(b >> 12);
    return __longlong_as_double((long long)bits) - 1.0;
}

__device__ __forceinline__ double pairwise9(const double* p) {
    // numpy add.reduce over 9 contiguous float64: 8 accumulators folded pairwise, then the tail
    return (((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]))) + p[8];
}

// ---- move selection, exact arithmetic ---------------------------------------------------------------
// Reproduces numpy bit for bit (movmodel.py:294-312): used in verification mode and on request.
template <bool HAS_FIELDS>
__device__ __forceinline__ int choose_exact(const TrackParams& P, const float2* base, int nc, unsigned mask, double u) {
    const float NINV_D = 0.70710677f;     // float32(1/sqrt(2)), movmodel.py:139-141
    double p[9];
    bool any_nz = false, any_nan = false;
    if (HAS_FIELDS) {
        const float2 fc = __ldg(base);
        const double uc = fmax((double)fc.x, 1e-06);                    // :295
        const double iuc = 1.0 / uc;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            // all nine entries, masked or not, like the reference: its NaN test (:228) sees the whole 3x3 patch
            const int dr = i / 3 - 1, dc = i % 3 - 1;
            const float2 f = __ldg(base + dr * nc + dc);
            const double ui = fmax((double)f.x, 1e-06);
            const double w = 2.0 / (iuc + 1.0 / ui);                    // :296, :260-261
            const float ninv = (i == 4) ? 0.0f : ((dr != 0 && dc != 0) ? NINV_D : 1.0f);
            const float d = __fmul_rn(__fsub_rn(fc.y, f.y), ninv);      // float32, :301-304
            double v = w * (double)d;                                   // :305
            any_nan |= (v != v);
            v = v > 0.0 ? v : 0.0;                                      // clip(min=0), :231
            if (i == 4 || !((mask >> i) & 1u)) v = 0.0;                 // :232-233
            p[i] = v;
            any_nz |= (v != 0.0);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) {                                    // 'drw': p = directional, :298-299
            p[i] = (i != 4 && ((mask >> i) & 1u)) ? P.dirp[i] : 0.0;
            any_nz |= (p[i] != 0.0);
        }
    }
    if (any_nan || !any_nz) {                                           // :228-230, :234-236
        any_nz = false;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            p[i] = (i != 4 && ((mask >> i) & 1u)) ? P.dirp[i] : 0.0;
            any_nz |= (p[i] != 0.0);
        }
        if (!any_nz) {                                                  // :239-240 (mask ignored)
#pragma unroll
            for (int i = 0; i < 9; ++i) p[i] = P.dirp[i];
        }
    }
    double s = pairwise9(p);                                            // :241
#pragma unroll
    for (int i = 0; i < 9; ++i) p[i] = (p[i] != 0.0) ? p[i] / s : 0.0;
    if (!P.nu_is_one) {                                                 // :242
#pragma unroll
        for (int i = 0; i < 9; ++i) p[i] = pow(p[i], P.nu);
    }
    s = pairwise9(p);                                                   // :243
#pragma unroll
    for (int i = 0; i < 9; ++i) p[i] = (p[i] != 0.0) ? p[i] / s : 0.0;
    // np.random.choice (:312): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, side='right')
    double cdf[9];
    cdf[0] = p[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) cdf[i] = cdf[i - 1] + p[i];
    const double tot = cdf[8];
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) idx += ((cdf[i] / tot) <= u) ? 1 : 0;
    return idx > 8 ? 8 : idx;
}

// ---- move selection, production arithmetic --------------------------------------------------------------
// Same distribution with the normalisations cancelled: only ratios of the weights matter, so
//   q_i = max(d_i, 0) * u_i / (u_c + u_i)      (= p_i / (2 u_c), movmodel.py:296-305)
// and the move is the first i (ascending flat index) whose running sum exceeds u * sum(q).  The two
// normalising divisions, the cdf division and 2/(1/a+1/b) of the exact form are gone: at most one division
// per allowed neighbour.  A draw differs from the exact form only if u lies within rounding (~1e-16) of a
// cdf boundary.  oracle/ssrs_oracle.c implements the same arithmetic (mode "fast"), so production runs are
// still reproduced bit for bit on the CPU.
template <bool HAS_FIELDS>
__device__ __forceinline__ double weight_fast(const TrackParams& P, const float2* base, const float2 fc, double uc,
                                              int nc, int i, bool& any_nan) {
    if (!HAS_FIELDS) return P.dirp[i];
    const int dr = i / 3 - 1, dc = i % 3 - 1;
    const float2 f = __ldg(base + dr * nc + dc);
    const float ninv = (dr != 0 && dc != 0) ? 0.70710677f : 1.0f;
    const float d = __fmul_rn(__fsub_rn(fc.y, f.y), ninv);               // float32, :301-304
    any_nan |= (d != d);
    if (!(d > 0.0f)) return 0.0;
    const double ui = fmax((double)f.x, 1e-06);
    return ((double)d * ui) / (uc + ui);
}

// all nine entries (first step of a track, nu != 1, or the unmasked directional fallback)
template <bool HAS_FIELDS>
__device__ __noinline__ int choose_fast_general(const TrackParams& P, const float2* base, int nc, unsigned mask, double u) {
    double q[9];
    bool any_nz = false, any_nan = false;
    float2 fc = make_float2(0.f, 0.f);
    double uc = 0.0;
    if (HAS_FIELDS) { fc = __ldg(base); uc = fmax((double)fc.x, 1e-06); }
    for (int i = 0; i < 9; ++i) {
        q[i] = (i != 4 && ((mask >> i) & 1u)) ? weight_fast<HAS_FIELDS>(P, base, fc, uc, nc, i, any_nan) : 0.0;
        any_nz |= (q[i] != 0.0);
    }
    if (any_nan || !any_nz) {
        any_nz = false;
        for (int i = 0; i < 9; ++i) { q[i] = (i != 4 && ((mask >> i) & 1u)) ? P.dirp[i] : 0.0; any_nz |= (q[i] != 0.0); }
        if (!any_nz)
            for (int i = 0; i < 9; ++i) q[i] = P.dirp[i];
    }
    if (!P.nu_is_one)
        for (int i = 0; i < 9; ++i) q[i] = pow(q[i], P.nu);
    double run = 0.0, tot = 0.0;
    for (int i = 0; i < 9; ++i) tot += q[i];
    const double target = u * tot;
    int idx = -1, last_pos = 4;
    for (int i = 0; i < 9; ++i) {
        run += q[i];
        if (q[i] > 0.0) last_pos = i;
        if (idx < 0 && run > target) idx = i;
    }
    return idx >= 0 ? idx : last_pos;
}

// fmax((double)x, 1e-6) (movmodel.py:294-295) for a float32 x, decided in float32: (double)x < 1e-6 exactly when
// x <= float32(1e-6) = 9.99999997e-07, the largest float32 below 1e-6; NaN -> 1e-6 like fmax.
__device__ __forceinline__ double clip_updraft(float x) {
    return (x > 9.99999997475242707e-07f) ? (double)x : 1e-06;
}

// the three neighbours within 45 degrees of the previous move, ascending flat index, 4 bits each
constexpr unsigned long long C3_A = (0x310ULL) | (0x210ULL << 12) | (0x521ULL << 24) | (0x630ULL << 36) | (0x000ULL << 48);
constexpr unsigned long long C3_B = (0x852ULL) | (0x763ULL << 12) | (0x876ULL << 24) | (0x875ULL << 36);

// Three-candidate step (every step after a track's first, nu == 1).  Division-free: with s_j = u_c + u_j,
//   q_i = max(d_i, 0) u_i / s_i   is proportional to   (d_i u_i) * (s_a s_b),  {a, b} = the other two candidates,
// so the weights need 3 adds and 9 multiplies.  f0..f2 are the candidates' {updraft, potential} pairs in
// ascending flat-index order, already loaded by the caller (so the loads overlap the Philox rounds).
template <bool HAS_FIELDS, bool MEM1>
__device__ __forceinline__ int choose_fast3(const TrackParams& P, const float2* base, int nc, unsigned mask,
                                            int i0, int i1, int i2, float2 fc, float2 f0, float2 f1, float2 f2,
                                            double u) {
    const bool e0 = MEM1 || ((mask >> i0) & 1u), e1 = MEM1 || ((mask >> i1) & 1u), e2 = MEM1 || ((mask >> i2) & 1u);
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
    bool any_nan = false;
    if (HAS_FIELDS) {
        const float n0 = ((i0 & 1) == 0) ? 0.70710677f : 1.0f;      // even flat index (0,2,6,8) = diagonal move
        const float n1 = ((i1 & 1) == 0) ? 0.70710677f : 1.0f;
        const float n2 = ((i2 & 1) == 0) ? 0.70710677f : 1.0f;
        const float d0 = __fmul_rn(__fsub_rn(fc.y, f0.y), n0);      // float32, movmodel.py:301-304
        const float d1 = __fmul_rn(__fsub_rn(fc.y, f1.y), n1);
        const float d2 = __fmul_rn(__fsub_rn(fc.y, f2.y), n2);
        any_nan = (e0 && d0 != d0) || (e1 && d1 != d1) || (e2 && d2 != d2);
        const double uc = clip_updraft(fc.x);
        const double u0 = clip_updraft(f0.x), u1 = clip_updraft(f1.x), u2 = clip_updraft(f2.x);
        const double s0 = uc + u0, s1 = uc + u1, s2 = uc + u2;
        if (e0 && d0 > 0.0f) q0 = ((double)d0 * u0) * (s1 * s2);
        if (e1 && d1 > 0.0f) q1 = ((double)d1 * u1) * (s0 * s2);
        if (e2 && d2 > 0.0f) q2 = ((double)d2 * u2) * (s0 * s1);
    } else {
        q0 = e0 ? P.dirp[i0] : 0.0;
        q1 = e1 ? P.dirp[i1] : 0.0;
        q2 = e2 ? P.dirp[i2] : 0.0;
    }
    if (any_nan || (q0 == 0.0 && q1 == 0.0 && q2 == 0.0)) {
        q0 = e0 ? P.dirp[i0] : 0.0;
        q1 = e1 ? P.dirp[i1] : 0.0;
        q2 = e2 ? P.dirp[i2] : 0.0;
        if (q0 == 0.0 && q1 == 0.0 && q2 == 0.0) return choose_fast_general<false>(P, base, nc, 0u, u);  // unmasked directional
    }
    const double c0 = q0, c1 = c0 + q1, c2 = c1 + q2;
    const double target = u * c2;
    if (c0 > target) return i0;
    if (c1 > target) return i1;
    if (c2 > target) return i2;
    return q2 > 0.0 ? i2 : (q1 > 0.0 ? i1 : i0);
}

// Out-of-line copy of the three-candidate step for the fast lane's rare cases (see fast_step).
__device__ __noinline__ int choose_fast3_rare(const TrackParams& P, const float2* base, int nc, int i0, int i1, int i2,
                                              float2 fc, float2 f0, float2 f1, float2 f2, double u) {
    return choose_fast3<true, true>(P, base, nc, 0u, i0, i1, i2, fc, f0, f1, f2, u);
}

__device__ __forceinline__ float fmax_nan(float a, float b) {      // NaN if either operand is NaN
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

__device__ __forceinline__ void red_add1(unsigned* p) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

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
__device__ __forceinline__ unsigned slot_of(unsigned flat) { return flat - (flat > 4u ? 1u : 0u); }
__device__ __forceinline__ unsigned flat_of(unsigned slot) { return slot + (slot >= 4u ? 1u : 0u); }

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
    red_add1(P.presence + lin);
}

// MEM1: track_dirn_restrict == 1 (the default): the mask is exactly the three candidates of the last move, so
// no history register, no mask arithmetic.
template <bool HAS_FIELDS, bool EXACT, bool MEM1>
__global__ void __launch_bounds__(128, 6) step_tracks_kernel(const TrackParams P) {
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
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nr = P.rows, nc = P.cols;
    const int kmax = P.kmax;
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
    // `k` the pair counter k / 2, `last` the previous move's slot (the kernel sits at the register limit of 6 CTAs/SM).
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
                if (t >= P.n_tracks) done = true;
                else {
                    int2 s = __ldg(P.start + t);
                    row = s.x; col = s.y; k = 0; last = 4; hist = 4; hcount = 1; run_mask = 0x1EF;
                    if (P.traj != nullptr && P.traj_cap > 0) P.traj[t] = make_short2((short)row, (short)col);
                    if (P.presence != nullptr) atomicAdd(P.presence + (long long)row * nc + col, 1u);
                    alive = true;
                }
            }
            if (!done) {
                if (fast_lane && last != 4u && (k & 1) == 0) {
                    const int m = min(min(row - 2, nr - 3 - row), min(col - 1, nc - 3 - col));
                    budget = max(min((m + 1) >> 1, (kmax - k) >> 1), 0);
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
                        // next track: first come, first served.  Track lengths are heavy-tailed (mean 1e4 steps, maximum
                        // above 1e5), so a fixed list per thread would leave most lanes idle while a few work through
                        // long lists; results do not depend on which lane steps which track (the random stream is keyed
                        // by the track id, the counts are integer sums).
                        t = stride + (long long)atomicAdd(P.next_track, 1ULL);
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
                                t = stride + (long long)atomicAdd(P.next_track, 1ULL);
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
        if (done) break;
        if (budget > 0) {
            const unsigned long long gid = (unsigned long long)(P.track_id0 + t);
            unsigned a, b, cc, dd;
            philox4x32_10((unsigned)gid, (unsigned)(gid >> 32), (unsigned)k, 0u, P, a, b, cc, dd);
            fast_step(P, s_lut, nc, row, last, fcen, ucen, uniform52(a, b));
            fast_step(P, s_lut, nc, row, last, fcen, ucen, uniform52(cc, dd));
            ++k; --budget;
        }
    }
    if (P.total_steps != nullptr) {
        // one atomic per warp
        for (int o = 16; o > 0; o >>= 1) steps_local += __shfl_down_sync(0xffffffffu, steps_local, o);
        if ((threadIdx.x & 31) == 0 && steps_local) atomicAdd(P.total_steps, steps_local);
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

extern "C" int ssrs_step_tracks(const float* fields, int rows, int cols, const int32_t* start_rc, int64_t n_tracks,
                                int64_t track_id0, const double* dirprob9_host, int memory, double nu, uint64_t seed,
                                const double* uniforms, int64_t uniforms_stride, int16_t* traj, int64_t traj_cap,
                                int32_t* traj_len, uint32_t* presence, unsigned long long* total_steps,
                                int flags, void* stream) {
    SSRS_REQUIRE(rows >= 5 && cols >= 5, "ssrs_step_tracks: grid %dx%d too small", rows, cols);
    SSRS_REQUIRE(rows <= 32767 && cols <= 32767, "ssrs_step_tracks: int16 trajectories need rows, cols <= 32767");
    SSRS_REQUIRE((long long)rows * cols < 2147483647LL, "ssrs_step_tracks: more than 2^31 cells");
    SSRS_REQUIRE(n_tracks >= 0 && track_id0 >= 0, "ssrs_step_tracks: negative track count or id");
    SSRS_REQUIRE(start_rc != nullptr || n_tracks == 0, "ssrs_step_tracks: start_rc is NULL");
    SSRS_REQUIRE(dirprob9_host != nullptr, "ssrs_step_tracks: dirprob9_host is NULL");
    SSRS_REQUIRE(uniforms == nullptr || uniforms_stride > 0, "ssrs_step_tracks: uniforms_stride must be positive");
    SSRS_REQUIRE(traj == nullptr || traj_cap > 0, "ssrs_step_tracks: traj given with traj_cap <= 0");
    if (memory < 0 || memory > 16) {
        set_error("ssrs_step_tracks: track_dirn_restrict=%d outside the supported range 0..16", memory);
        return SSRS_ERR_UNSUPPORTED;
    }
    if (n_tracks == 0) return SSRS_OK;
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
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    {
        const int rc = queue_slot(&P.next_track);
        if (rc != SSRS_OK) return rc;
    }
    SSRS_CUDA_TRY(cudaMemsetAsync(P.next_track, 0, sizeof(unsigned long long), static_cast<cudaStream_t>(stream)));
    kern<<<(int)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(P);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
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
