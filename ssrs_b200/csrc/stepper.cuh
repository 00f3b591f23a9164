// Device code shared by the stepping kernels (tracks.cu: gather-and-evaluate stepper; walk.cu: transition-table walk):
// direction-memory masks, Philox4x32-10, the move selection in the reference's exact arithmetic and in the production
// arithmetic.  Translation units that include this header are compiled with -fmad=false (bit-exact float64).
#pragma once
#include "common.cuh"

#include <math.h>

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
    // Input of a launch: entries are taken first come, first served through the device counter `in_head` (zero at launch).
    // Fresh tracks (in_list == NULL): entry i is track i of `start`.  A later phase of a phased launch: entry i is the
    // saved state {track, row | col << 16, step, previous move} of a track that outlived the previous phase; their number
    // is the device counter `in_count`.  Tracks that reach step `kcap` are appended to out_list / out_count.
    unsigned* in_head;
    const unsigned* in_count;
    const uint4* in_list;
    uint4* out_list;
    unsigned* out_count;
    int kcap;
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

// Phase schedule of a phased launch (host): step counts at which the survivors are compacted, multiples of 4, geometric
// (x 1.25) from about one crossing of the grid's short side (or `first`, for tests) up to kmax; the last phase runs to
// the end.  Workspace of a phased launch: two state lists of n_tracks uint4 each, then [phase][head, count] counters.
constexpr int kMaxPhases = 64;
inline int phase_caps(int rows, int cols, int kmax, int first, int* caps) {
    long long c = (rows < cols ? rows : cols);
    if (c < 1024) c = 1024;
    if (first > 0) c = first;
    int n = 0;
    while (n < kMaxPhases - 1) {
        c = (c + 3) & ~3LL;
        if (c >= kmax) break;
        caps[n++] = (int)c;
        c += c / 4 > 4 ? c / 4 : 4;
    }
    caps[n++] = 2147483644;
    return n;
}
inline long long phase_workspace_bytes(long long n_tracks) { return 2 * n_tracks * (long long)sizeof(uint4) + 4096; }

// "slot" of a move = its flat index with the centre (4) squeezed out, 0..7
__device__ __forceinline__ unsigned slot_of(unsigned flat) { return flat - (flat > 4u ? 1u : 0u); }
__device__ __forceinline__ unsigned flat_of(unsigned slot) { return slot + (slot >= 4u ? 1u : 0u); }

}  // namespace
}  // namespace ssrs
