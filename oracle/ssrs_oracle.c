/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the SSRS track stepper and presence count.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the library built
 * from this file; the product (ssrs_b200/) never does.  It restates, line for line,
 *   generate_simulated_tracks      /root/reference/ssrs/movmodel.py:264-318
 *   move_away_from_boundary        /root/reference/ssrs/movmodel.py:205-217
 *   get_track_restrictions         /root/reference/ssrs/movmodel.py:185-202  (as a table)
 *   generate_move_probabilities    /root/reference/ssrs/movmodel.py:220-244
 *   np.random.choice(range(9), p)  numpy legacy: cdf = cumsum(p); cdf /= cdf[-1];
 *                                  searchsorted(u, 'right') with ONE random_sample() per call
 *   compute_presence_counts        /root/reference/ssrs/movmodel.py:410-419  (int32, no int16 wrap)
 * in the reference's arithmetic: float32 potential differences (:301-304), float64 elsewhere, numpy's
 * pairwise summation order for the 9-element sums.  Pinned by tests/test_oracle_golden.py against
 * trajectories produced by the unmodified reference (tests/golden/, made by oracle/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (oracle/build_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const unsigned RESTRICT_LUT[9] = {0x00B, 0x007, 0x026, 0x049, 0x1EF, 0x124, 0x0C8, 0x1C0, 0x1A0};

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t* o) {
    for (int i = 0; i < 10; ++i) {
        uint64_t m0 = (uint64_t)0xD2511F53u * c0, m1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(m1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)m1;
        uint32_t n2 = (uint32_t)(m0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)m0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

/* Uniform for step `step` of track `track`: Philox4x32-10 with counter (track, step >> 1) and key = seed yields
 * four words = two uniforms (words {0,1} for even steps, {2,3} for odd steps); 52 mantissa bits -> [0,1).
 * Same construction as ssrs_b200/csrc/tracks.cu. */
double oracle_philox_uniform(uint64_t seed, uint64_t track, uint32_t step) {
    uint32_t w[4];
    philox4x32_10((uint32_t)track, (uint32_t)(track >> 32), step >> 1, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    const uint32_t a = w[2 * (step & 1u)], b = w[2 * (step & 1u) + 1];
    const uint64_t bits = 0x3FF0000000000000ULL | ((uint64_t)a << 20) | (uint64_t)(b >> 12);
    double d;
    memcpy(&d, &bits, 8);
    return d - 1.0;
}

void oracle_philox_words(uint64_t seed, uint64_t track, uint32_t block, uint32_t* out4) {
    philox4x32_10((uint32_t)track, (uint32_t)(track >> 32), block, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), out4);
}

static double pairwise9(const double* p) {
    return (((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]))) + p[8];
}

/* Move selection in the reference's exact operation order (movmodel.py:294-312). */
static int choose_exact(const float* U, const float* P, int nc, int r, int c, const double* dirp, unsigned mask,
                        double nu, double u) {
    const float ninv_d = 0.70710677f;
    double p[9];
    int any_nan = 0;
    if (U != NULL) {
        const int64_t o = (int64_t)r * nc + c;
        double uc = (double)U[o]; if (uc < 1e-06) uc = 1e-06;      /* :295 */
        const double iuc = 1.0 / uc;
        for (int i = 0; i < 9; ++i) {
            const int dr = i / 3 - 1, dc = i % 3 - 1;
            const int64_t q = o + (int64_t)dr * nc + dc;
            double ui = (double)U[q]; if (ui < 1e-06) ui = 1e-06;
            const double w = 2.0 / (iuc + 1.0 / ui);               /* :296, :260-261 */
            const float ninv = (i == 4) ? 0.0f : ((dr != 0 && dc != 0) ? ninv_d : 1.0f);
            const float d = (float)(P[o] - P[q]) * ninv;           /* float32, :301-304 */
            p[i] = w * (double)d;                                  /* :305 */
            if (p[i] != p[i]) any_nan = 1;
        }
    } else {
        for (int i = 0; i < 9; ++i) p[i] = dirp[i];                /* :298-299 */
    }
    if (any_nan) for (int i = 0; i < 9; ++i) p[i] = dirp[i];       /* :228-230 */
    int nz = 0;
    for (int i = 0; i < 9; ++i) {                                  /* :231-233 */
        if (p[i] < 0.0) p[i] = 0.0;
        if (i == 4 || !((mask >> i) & 1u)) p[i] = 0.0;
        if (p[i] != 0.0) nz++;
    }
    if (nz == 0) {                                                 /* :234-238 */
        for (int i = 0; i < 9; ++i) {
            p[i] = (i == 4 || !((mask >> i) & 1u)) ? 0.0 : dirp[i];
            if (p[i] != 0.0) nz++;
        }
    }
    if (nz == 0) for (int i = 0; i < 9; ++i) p[i] = dirp[i];       /* :239-240 */
    double s = pairwise9(p);                                       /* :241 */
    for (int i = 0; i < 9; ++i) p[i] = p[i] / s;
    if (nu != 1.0) for (int i = 0; i < 9; ++i) p[i] = pow(p[i], nu);   /* :242 */
    s = pairwise9(p);                                              /* :243 */
    for (int i = 0; i < 9; ++i) p[i] = p[i] / s;
    double cdf[9];
    cdf[0] = p[0];
    for (int i = 1; i < 9; ++i) cdf[i] = cdf[i - 1] + p[i];
    const double tot = cdf[8];
    int idx = 0;
    for (int i = 0; i < 9; ++i) if (cdf[i] / tot <= u) idx++;      /* searchsorted side='right' */
    if (idx > 8) idx = 8;
    return idx;
}

/* "Production arithmetic" of the CUDA stepper (ssrs_b200/csrc/tracks.cu, choose_fast): the normalisations of
 * movmodel.py:241-243 and of np.random.choice cancel, so q_i = max(d_i,0) * u_i / (u_c + u_i) (= p_i / (2 u_c))
 * and the move is the first index whose running sum exceeds u * sum(q).  Same distribution as choose_exact up to
 * rounding of the cdf boundaries; restated here operation for operation so GPU production runs can be checked
 * bit for bit on the CPU.  Only mask-allowed neighbours are evaluated (also for the NaN test). */
static const int CAND3[9][3] = {{0, 1, 3}, {0, 1, 2}, {1, 2, 5}, {0, 3, 6}, {-1, -1, -1}, {2, 5, 8}, {3, 6, 7}, {6, 7, 8}, {5, 7, 8}};

static int choose_fast(const float* U, const float* P, int nc, int r, int c, const double* dirp, unsigned mask,
                       double nu, double u, int last) {
    const float ninv_d = 0.70710677f;
    if (last != 4 && nu == 1.0) {
        /* three-candidate step, division-free: q_i ~ (d_i u_i) (s_a s_b), s_j = u_c + u_j, {a,b} the other two */
        const int* ci = CAND3[last];
        const int64_t o = (int64_t)r * nc + c;
        double q3[3] = {0.0, 0.0, 0.0};
        int en[3], nan3 = 0;
        for (int j = 0; j < 3; ++j) en[j] = (mask >> ci[j]) & 1u;
        if (U != NULL) {
            double uc = (double)U[o]; if (uc < 1e-06) uc = 1e-06;
            double uu[3], ss[3]; float dd[3];
            for (int j = 0; j < 3; ++j) {
                const int dr = ci[j] / 3 - 1, dc = ci[j] % 3 - 1;
                const int64_t qn = o + (int64_t)dr * nc + dc;
                const float ninv = (dr != 0 && dc != 0) ? ninv_d : 1.0f;
                dd[j] = (float)(P[o] - P[qn]) * ninv;
                if (en[j] && dd[j] != dd[j]) nan3 = 1;
                uu[j] = (double)U[qn]; if (uu[j] < 1e-06) uu[j] = 1e-06;
                ss[j] = uc + uu[j];
            }
            if (en[0] && dd[0] > 0.0f) q3[0] = ((double)dd[0] * uu[0]) * (ss[1] * ss[2]);
            if (en[1] && dd[1] > 0.0f) q3[1] = ((double)dd[1] * uu[1]) * (ss[0] * ss[2]);
            if (en[2] && dd[2] > 0.0f) q3[2] = ((double)dd[2] * uu[2]) * (ss[0] * ss[1]);
        } else {
            for (int j = 0; j < 3; ++j) q3[j] = en[j] ? dirp[ci[j]] : 0.0;
        }
        int fall = 0;
        if (nan3 || (q3[0] == 0.0 && q3[1] == 0.0 && q3[2] == 0.0)) {
            for (int j = 0; j < 3; ++j) q3[j] = en[j] ? dirp[ci[j]] : 0.0;
            if (q3[0] == 0.0 && q3[1] == 0.0 && q3[2] == 0.0) fall = 1;
        }
        if (!fall) {
            const double c0 = q3[0], c1 = c0 + q3[1], c2 = c1 + q3[2];
            const double target = u * c2;
            if (c0 > target) return ci[0];
            if (c1 > target) return ci[1];
            if (c2 > target) return ci[2];
            return q3[2] > 0.0 ? ci[2] : (q3[1] > 0.0 ? ci[1] : ci[0]);
        }
        U = NULL; P = NULL; mask = 0u;          /* unmasked directional weights through the general path */
    }
    double q[9];
    int any_nan = 0, nz = 0;
    const int64_t o = (int64_t)r * nc + c;
    double uc = 0.0;
    if (U != NULL) { uc = (double)U[o]; if (uc < 1e-06) uc = 1e-06; }
    for (int i = 0; i < 9; ++i) {
        q[i] = 0.0;
        if (i == 4 || !((mask >> i) & 1u)) continue;
        if (U == NULL) { q[i] = dirp[i]; }
        else {
            const int dr = i / 3 - 1, dc = i % 3 - 1;
            const int64_t qn = o + (int64_t)dr * nc + dc;
            const float ninv = (dr != 0 && dc != 0) ? ninv_d : 1.0f;
            const float d = (float)(P[o] - P[qn]) * ninv;
            if (d != d) any_nan = 1;
            if (d > 0.0f) {
                double ui = (double)U[qn]; if (ui < 1e-06) ui = 1e-06;
                q[i] = ((double)d * ui) / (uc + ui);
            }
        }
        if (q[i] != 0.0) nz++;
    }
    if (any_nan || nz == 0) {
        nz = 0;
        for (int i = 0; i < 9; ++i) {
            q[i] = (i == 4 || !((mask >> i) & 1u)) ? 0.0 : dirp[i];
            if (q[i] != 0.0) nz++;
        }
        if (nz == 0) for (int i = 0; i < 9; ++i) q[i] = dirp[i];
    }
    if (nu != 1.0) for (int i = 0; i < 9; ++i) q[i] = pow(q[i], nu);
    double tot = 0.0;
    for (int i = 0; i < 9; ++i) tot += q[i];
    const double target = u * tot;
    double run = 0.0;
    int idx = -1, last_pos = 4;
    for (int i = 0; i < 9; ++i) {
        run += q[i];
        if (q[i] > 0.0) last_pos = i;
        if (idx < 0 && run > target) idx = i;
    }
    return idx >= 0 ? idx : last_pos;
}

/* One track.  U (float32, widened to double as the reference's float64 updraft) and P (float32) are
 * [rows][cols]; either both given (fluid-flow) or both NULL ('drw').  uniforms: per-step numbers for this
 * track or NULL for Philox(seed, gid, k).  traj: optional int16 [cap][2].  presence: optional int32
 * [rows][cols] (incremented; atomically when built with OpenMP).  Returns the number of points. */
static int64_t one_track(const float* U, const float* P, int nr, int nc, int row, int col, const double* dirp,
                         int memory, double nu, uint64_t seed, uint64_t gid, const double* uniforms,
                         int16_t* traj, int64_t cap, int32_t* presence, int fast) {
    const int burnin = (int)((nr < nc ? nr : nc) / 10);                /* :276 */
    const double max_moves = (double)nr / 2 * (double)nc / 2;          /* :277 */
    int hist_len = 1, hist_cap = 64;
    unsigned char* hist = (unsigned char*)malloc((size_t)hist_cap);    /* directions list, :280-281 */
    hist[0] = 4;
    int64_t k = 0;
    if (traj && cap > 0) { traj[0] = (int16_t)row; traj[1] = (int16_t)col; }
    if (presence) {
#pragma omp atomic
        presence[(int64_t)row * nc + col] += 1;
    }
    while ((double)k < max_moves) {                                    /* :285 */
        int r = row, c = col;
        if (k > burnin) {                                              /* :287-289 */
            if (!(0 < r && r < nr - 1 && 0 < c && c < nc - 1)) break;
        } else {                                                       /* :290-291 */
            if (r <= 1) r += 2; else if (r >= nr - 2) r -= 2;
            if (c <= 0) c += 2; else if (c >= nc - 2) c -= 2;
        }
        unsigned mask = 0x1EF;                                         /* :307-309 */
        {
            int m = (memory == 0 || memory > hist_len) ? hist_len : memory;   /* directions[-0:] is the whole list */
            for (int j = 0; j < m; ++j) mask &= RESTRICT_LUT[hist[hist_len - 1 - j]];
        }
        const double u = uniforms ? uniforms[k] : oracle_philox_uniform(seed, gid, (uint32_t)k);
        int idx;
        if (fast) idx = choose_fast(U, P, nc, r, c, dirp, mask, nu, u, (int)hist[hist_len - 1]);
        else idx = choose_exact(U, P, nc, r, c, dirp, mask, nu, u);
        row = r + (idx / 3 - 1);                                       /* :313-317 */
        col = c + (idx % 3 - 1);
        ++k;
        if (hist_len == hist_cap) { hist_cap *= 2; hist = (unsigned char*)realloc(hist, (size_t)hist_cap); }
        hist[hist_len++] = (unsigned char)idx;
        if (traj && k < cap) { traj[2 * k] = (int16_t)row; traj[2 * k + 1] = (int16_t)col; }
        if (presence) {
#pragma omp atomic
            presence[(int64_t)row * nc + col] += 1;
        }
    }
    free(hist);
    return k + 1;
}

/* ---- "table" mode: the transition-table walk of ssrs_b200/csrc/walk.cu, restated step for step ----------------------
 * With direction memory 1 and nu = 1 the move distribution is a function of (cell, previous move).  walk.cu tabulates
 * it once per (case, realisation) as two cumulative thresholds on the 2^-31 lattice and steps with one 31-bit Philox
 * word per move; here the same thresholds are evaluated on demand.  Step types and random-number mapping as documented
 * at the top of walk.cu. */
static const uint32_t W_ONE = 0x80000000u, W_UNMASKED = 0xFFFFFFFFu;

static uint32_t prob31(double p) {
    const double x = p * 2147483648.0 + 0.5;
    return x >= 2147483648.0 ? W_ONE : (uint32_t)x;
}

static double clip_updraft(float x) { return (x > 9.99999997475242707e-07f) ? (double)x : 1e-06; }

static void table_entry(const float* U, const float* P, int nc, int r, int c, int last, const double* dirp,
                        uint32_t* t1, uint32_t* t2) {
    const float ninv_d = 0.70710677f;
    const int* ci = CAND3[last];
    const int64_t o = (int64_t)r * nc + c;
    const double uc = clip_updraft(U[o]);
    double dq[3], ss[3];
    for (int j = 0; j < 3; ++j) {
        const int dr = ci[j] / 3 - 1, dc = ci[j] % 3 - 1;
        const int64_t qn = o + (int64_t)dr * nc + dc;
        const float ninv = (dr != 0 && dc != 0) ? ninv_d : 1.0f;
        const float d = (float)(P[o] - P[qn]) * ninv;
        const float dm = (d != d) ? d : (d > 0.0f ? d : 0.0f);                 /* max.NaN(d, 0) */
        const double u = clip_updraft(U[qn]);
        dq[j] = (double)dm * u;
        ss[j] = uc + u;
    }
    double q0 = dq[0] * (ss[1] * ss[2]), q1 = dq[1] * (ss[0] * ss[2]), q2 = dq[2] * (ss[0] * ss[1]);
    double c1 = q0 + q1, c2 = c1 + q2;
    if (!(c2 > 0.0 && c2 <= 1.7976931348623157e308)) {
        q0 = dirp[ci[0]]; q1 = dirp[ci[1]]; q2 = dirp[ci[2]];
        c1 = q0 + q1; c2 = c1 + q2;
        if (!(c2 > 0.0)) { *t1 = W_UNMASKED; *t2 = 0; return; }
    }
    const double inv = 1.0 / c2;
    uint32_t a = (q0 > 0.0) ? prob31(q0 * inv) : 0u;
    uint32_t b = (q2 > 0.0) ? prob31(c1 * inv) : W_ONE;
    if (!(q1 > 0.0) && !(q2 > 0.0)) a = W_ONE;
    if (b < a) b = a;
    *t1 = a; *t2 = b;
}

static int64_t one_track_table(const float* U, const float* P, int nr, int nc, int row, int col, const double* dirp,
                               uint64_t seed, uint64_t gid, int32_t* presence) {
    const int burnin = (int)((nr < nc ? nr : nc) / 10);
    const double max_moves = (double)nr / 2 * (double)nc / 2;
    const double km = ceil(max_moves);
    const int64_t kmax = km > 2147483647.0 ? 2147483647 : (int64_t)km;
    uint32_t dthr[9];
    {
        double run[9], acc = 0.0;
        int lastpos = -1;
        for (int i = 0; i < 9; ++i) { acc += dirp[i]; run[i] = acc; if (dirp[i] > 0.0) lastpos = i; }
        for (int i = 0; i < 9; ++i) {
            const double x = run[i] / run[8] * 2147483648.0 + 0.5;
            dthr[i] = (i >= lastpos || x >= 2147483648.0) ? W_ONE : (uint32_t)x;
        }
    }
    int64_t k = 0;
    int last = 4, tmode = 0;
    if (presence) {
#pragma omp atomic
        presence[(int64_t)row * nc + col] += 1;
    }
    for (;;) {
        const int elig = last != 4 && row >= 2 && row <= nr - 3 && col >= 1 && col <= nc - 3;
        if ((k & 3) == 0) tmode = elig && k + 4 <= kmax;
        else tmode = tmode && elig;
        int idx;
        if (tmode) {
            uint32_t w[4], t1, t2;
            philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(k >> 2), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
            const uint32_t r31 = w[k & 3] >> 1;
            table_entry(U, P, nc, row, col, last, dirp, &t1, &t2);
            if (t1 == W_UNMASKED) {
                int cnt = 0;
                for (int i = 0; i < 9; ++i) cnt += (r31 >= dthr[i]) ? 1 : 0;
                idx = cnt < 8 ? cnt : 8;
            } else idx = CAND3[last][(r31 < t1) ? 0 : ((r31 < t2) ? 1 : 2)];
            row += idx / 3 - 1;
            col += idx % 3 - 1;
        } else {
            if (k >= kmax) break;                                              /* movmodel.py:285 */
            int r = row, c = col;
            if (k > burnin) {
                if (!(0 < r && r < nr - 1 && 0 < c && c < nc - 1)) break;
            } else {
                if (r <= 1) r += 2; else if (r >= nr - 2) r -= 2;
                if (c <= 0) c += 2; else if (c >= nc - 2) c -= 2;
            }
            uint32_t w[4];
            philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)k, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
            const uint64_t bits = 0x3FF0000000000000ULL | ((uint64_t)w[0] << 20) | (uint64_t)(w[1] >> 12);
            double u;
            memcpy(&u, &bits, 8);
            u -= 1.0;
            idx = choose_fast(U, P, nc, r, c, dirp, RESTRICT_LUT[last], 1.0, u, last);
            row = r + (idx / 3 - 1);
            col = c + (idx % 3 - 1);
        }
        ++k;
        last = idx;
        if (presence) {
#pragma omp atomic
            presence[(int64_t)row * nc + col] += 1;
        }
    }
    return k + 1;
}

/* Batch driver.  start_rc int32 [n][2]; uniforms [n][ustride] or NULL; traj int16 [n][cap][2] track-major
 * or NULL; traj_len int32 [n] or NULL; presence int32 [rows][cols] or NULL.  Returns total track-steps. */
int64_t oracle_step_tracks(const float* U, const float* P, int rows, int cols, const int32_t* start_rc, int64_t n,
                           int64_t track_id0, const double* dirp, int memory, double nu, uint64_t seed,
                           const double* uniforms, int64_t ustride, int16_t* traj, int64_t cap, int32_t* traj_len,
                           int32_t* presence, int nthreads, int fast) {
    int64_t total = 0;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total) num_threads(nthreads)
#endif
    for (int64_t t = 0; t < n; ++t) {
        int64_t len;
        if (fast == 2)      /* transition-table walk (memory 1, nu 1, Philox, fields given, no trajectories) */
            len = one_track_table(U, P, rows, cols, start_rc[2 * t], start_rc[2 * t + 1], dirp, seed,
                                  (uint64_t)(track_id0 + t), presence);
        else
            len = one_track(U, P, rows, cols, start_rc[2 * t], start_rc[2 * t + 1], dirp, memory, nu, seed,
                            (uint64_t)(track_id0 + t), uniforms ? uniforms + t * ustride : NULL,
                            traj ? traj + t * cap * 2 : NULL, cap, presence, fast);
        if (traj_len) traj_len[t] = (int32_t)len;
        total += len - 1;
    }
    return total;
}

/* compute_presence_counts, movmodel.py:410-419, for track-major trajectories. */
void oracle_presence_counts(const int16_t* traj, int64_t cap, const int32_t* traj_len, int64_t n, int rows, int cols,
                            int32_t* presence) {
    for (int64_t t = 0; t < n; ++t) {
        int64_t len = traj_len[t] < cap ? traj_len[t] : cap;
        for (int64_t k = 0; k < len; ++k) {
            int r = traj[(t * cap + k) * 2], c = traj[(t * cap + k) * 2 + 1];
            if (r >= 0 && r < rows && c >= 0 && c < cols) presence[(int64_t)r * cols + c] += 1;
        }
    }
}
