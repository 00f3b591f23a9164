// Stage 2 — Brandes–Ombalski fluid-flow potential on the conductivity raster (sm_100a).
//
// Replaces MovModel.assemble_sparse_linear_system + solve_sparse_linear_system
// (ssrs/movmodel.py:59-84, :86-128): the reference builds the 8-neighbour conductance graph entry by
// entry in Python, row-normalises it and hands it to SuperLU.  Here the same linear system
//     sum_j g_ij (phi_i - phi_j) = 0  at free nodes,   phi = given at Dirichlet nodes,
//     g_ij = hm(K_i, K_j) / f_ij,  hm = 1e-8 if K_i == 0 or K_j == 0 else 2/(1/K_i + 1/K_j),
//     f_ij = 1 (axial) | float32(sqrt 2) (diagonal), with the reference's last-column S/SW factor swap
//  is solved without ever forming the fine matrix:
//   * fine level: matrix-free, from four forward link weights per cell (float64 for the exact operator);
//   * preconditioner: aggregation AMG.  The conductances span ten orders of magnitude (half the cells sit
//     on the 1e-8 floor), so coarsening must follow the strong couplings: per level one pass of pairwise
//     matching by mutual strongest connection (handshake rounds), then unmatched nodes join the aggregate
//     of their strongest neighbour.  Prolongation is piecewise constant, coarse operators are Galerkin
//     sums (CSR for setup, sliced ELL with 4-byte bfloat16 entries for the cycle), smoothing is weighted Jacobi, V(1,1) cycle with an
//     over-corrected coarse-grid correction, dense float64 inverse on the coarsest level;
//   * outer iteration: right-preconditioned BiCGStab in float64 (the last-column quirk makes the operator
//     non-symmetric), true-residual restarts; result rounded to float32 like the reference (:128).
// Geometric multigrid (also operator-dependent/BoxMG interpolation) stalls on this problem: conducting
// islands that contain no coarse point lose their constant mode.  See DESIGN.md.
//
// All kernels are lambdas over pfor()/pfor2d()/preduce() (pfor.cuh).
#include "pfor.cuh"

#include <math.h>
#include <stdio.h>

#include <chrono>

#include "../../include/ssrs_b200.h"

namespace ssrs {
void set_error(const char* fmt, ...);
#ifdef SSRS_HOST_EMU
#include <stdarg.h>
static thread_local char g_emu_err[512];
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_emu_err, sizeof(g_emu_err), fmt, ap);
    va_end(ap);
}
#else
int sm_count();
namespace par {
int grid_cap() { return sm_count() * 8; }
Arena& arena(bool temp) {
    static thread_local Arena a[64][2];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    return a[dev][temp ? 1 : 0];
}
double* reduce_scratch() {   // small persistent buffer (plain cudaMalloc, lives for the process)
    static thread_local double* buf = nullptr;
    static thread_local int dev_of = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (buf == nullptr || dev_of != dev) {
        if (cudaMalloc(&buf, sizeof(double) * (2 * RED_BLOCKS + 2)) != cudaSuccess) return nullptr;
        dev_of = dev;
    }
    return buf;
}
}  // namespace par
#endif

namespace amg {
using namespace par;

typedef int64_t i64;
#ifdef SSRS_CYCLE_F32
typedef float real;      // experiment only: float32 cycle vectors lose the island modes (see the cycle's header comment)
#else
typedef double real;     // precision of the V-cycle's vectors (its operators are float32)
#endif

constexpr double HM_FLOOR = 1e-08;                                    // movmodel.py:104-105
constexpr double INV_SQRT2_F32 = 1.0 / 1.41421353816986083984375;     // facs are float32, movmodel.py:82
enum { E_SKIP = 0, E_OFF = 1, E_DIR = 2, E_DIAG = 3 };

SSRS_HD inline bool sign_set(float x) {
    unsigned u;
    memcpy(&u, &x, 4);
    return (u >> 31) != 0;
}
SSRS_HD inline unsigned pair_hash(unsigned a, unsigned b) {
    unsigned lo = a < b ? a : b, hi = a < b ? b : a;
    unsigned h = lo * 0x9E3779B1u ^ (hi + 0x7F4A7C15u) * 0x85EBCA6Bu;
    h ^= h >> 15; h *= 0xC2B2AE35u; h ^= h >> 13;
    return h;
}

// ---- ownership of a level's nodes (row-sharded solve) ------------------------------------------------------
// part p owns the contiguous index range [lo[p], lo[p+1]); n == 1 is the single-GPU case
struct Parts {
    int n = 1;
    i64 lo[SSRS_MAX_RANKS + 1] = {0};
};
SSRS_HD inline bool parts_cross(const Parts& P, i64 i, i64 j) {
    for (int p = 1; p < P.n; ++p)
        if ((i < P.lo[p]) != (j < P.lo[p])) return true;
    return false;
}
SSRS_HD inline int parts_owner(const Parts& P, i64 i) {
    int p = 0;
    while (p + 1 < P.n && i >= P.lo[p + 1]) ++p;
    return p;
}

// ---- graph providers ----------------------------------------------------------------------------
// entry(i, k, j, a): k-th stored entry of row i -> kind, column j, value a (off-diagonals are negative).
struct FineGraph {
    const float* kd;     // conductivity with the sign bit marking Dirichlet nodes
    int rows, cols;
    int interior_dirichlet;   // 1 if some Dirichlet node is more than one cell away from the border
    Parts parts;
    SSRS_HD bool crosses(i64 i, i64 j) const { return parts_cross(parts, i, j); }
    SSRS_HD i64 size() const { return (i64)rows * cols; }
    SSRS_HD bool excluded(i64 i) const { return sign_set(kd[i]); }
    SSRS_HD void range(i64, i64& k0, i64& k1) const { k0 = 0; k1 = 9; }
    SSRS_HD int entry(i64 i, i64 k, i64& j, double& a) const {
        if (k == 4) return E_SKIP;
        const int r = (int)(i / cols), c = (int)(i - (i64)r * cols);
        const int dr = (int)k / 3 - 1, dc = (int)k % 3 - 1;
        const int rr = r + dr, cc = c + dc;
        if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) return E_SKIP;          // movmodel.py:75
        j = (i64)rr * cols + cc;
        const float kj = kd[j];
        const double ka = fabs((double)kd[i]), kb = fabs((double)kj);
        const double hm = (ka == 0.0 || kb == 0.0) ? HM_FLOOR : 2.0 * ka * kb / (ka + kb);   // :442-447
        bool diagonal = (dr != 0) && (dc != 0);
        if (c == cols - 1 && r >= 1 && r <= rows - 2 && dr == -1) {               // :73-79 last-column quirk
            if (dc == 0) diagonal = true;
            else if (dc == -1) diagonal = false;
        }
        a = -(diagonal ? hm * INV_SQRT2_F32 : hm);
        return sign_set(kj) ? E_DIR : E_OFF;
    }
};

struct CsrGraph {
    const i64* rowptr;
    const int* col;
    const double* val;
    i64 n;
    Parts parts;
    SSRS_HD bool crosses(i64 i, i64 j) const { return parts_cross(parts, i, j); }
    SSRS_HD i64 size() const { return n; }
    SSRS_HD bool excluded(i64) const { return false; }
    SSRS_HD void range(i64 i, i64& k0, i64& k1) const { k0 = rowptr[i]; k1 = rowptr[i + 1]; }
    SSRS_HD int entry(i64 i, i64 k, i64& j, double& a) const {
        j = col[k];
        a = val[k];
        return j == i ? E_DIAG : E_OFF;
    }
};

// ---- fine level ---------------------------------------------------------------------------------------
// The generic providers above are used for setup only.  The cycle and Krylov kernels evaluate a cell's eight
// links from precomputed forward link weights.  link_weight<true> is the cheap float32 form (not used by the
// shipped kernels, kept for experiments); weights are computed in float64 and rounded once for the cycle.
template <bool FAST>
SSRS_HD inline double link_weight(float ka_raw, float kb_raw, bool diagonal) {
    if (FAST) {
        const float a = fabsf(ka_raw), b = fabsf(kb_raw);
        float hm;
        if (a == 0.0f || b == 0.0f) hm = 1e-08f;
        else {
#ifdef __CUDA_ARCH__
            hm = __fdividef(2.0f * a * b, a + b);
#else
            hm = 2.0f * a * b / (a + b);
#endif
        }
        return (double)(diagonal ? hm * 0.70710677f : hm);
    } else {
        const double a = fabs((double)ka_raw), b = fabs((double)kb_raw);
        const double hm = (a == 0.0 || b == 0.0) ? HM_FLOOR : 2.0 * a * b / (a + b);
        return diagonal ? hm * INV_SQRT2_F32 : hm;
    }
}

// Precomputed forward link weights (E, N, NE, NW of every cell; 0 where the neighbour is outside the grid) in
// float32 (cycle) and float64 (exact operator).  A cell's other four links are its neighbours' forward links —
// except on the reference's quirk column (last column, interior rows) whose S and SW links carry swapped
// distance factors and are evaluated from K directly.  All eight neighbours are included: for error-equation
// vectors the caller keeps x = 0 at Dirichlet nodes, which is exactly "only load the diagonal".
struct FineWeights {
    const float* wf;    // [4][n]
    const double* wd;   // [4][n]
};

// (A x)_i of the exact float64 operator for cell (r, c)
SSRS_HD inline double fine_apply64(const FineGraph& g, const FineWeights& W, int r, int c, const double* x) {
    const int cols = g.cols;
    const i64 n = (i64)g.rows * cols;
    const i64 i = (i64)r * cols + c;
    const double* w = W.wd;
    if (r > 0 && r < g.rows - 1 && c > 0 && c < cols - 1) {
        // interior cell (the quirk column is on the border): fixed offsets, same summation order as below
        const double* wN = w + n; const double* wNE = w + 2 * n; const double* wNW = w + 3 * n;
        const double xi = x[i];
        double s = w[i] * (xi - x[i + 1]) + w[i - 1] * (xi - x[i - 1]);
        s += wN[i] * (xi - x[i + cols]) + wN[i - cols] * (xi - x[i - cols]);
        s += wNE[i] * (xi - x[i + cols + 1]) + wNE[i - cols - 1] * (xi - x[i - cols - 1]);
        s += wNW[i] * (xi - x[i + cols - 1]) + wNW[i - cols + 1] * (xi - x[i - cols + 1]);
        return s;
    }
    const bool hW = c > 0, hE = c < cols - 1, hS = r > 0, hN = r < g.rows - 1;
    // a missing neighbour aliases the cell itself: its difference is exactly zero whatever weight is read
    const i64 jE = hE ? i + 1 : i, jW = hW ? i - 1 : i, jN = hN ? i + cols : i, jS = hS ? i - cols : i;
    const i64 jNE = (hN && hE) ? i + cols + 1 : i, jNW = (hN && hW) ? i + cols - 1 : i;
    const i64 jSW = (hS && hW) ? i - cols - 1 : i, jSE = (hS && hE) ? i - cols + 1 : i;
    double wS = w[n + jS], wSW = w[2 * n + jSW];
    if (c == cols - 1 && hS && hN) {                                        // movmodel.py:73-79
        wS = link_weight<false>(g.kd[i], g.kd[jS], true);
        wSW = link_weight<false>(g.kd[i], g.kd[jSW], false);
    }
    const double xi = x[i];
    double s = w[i] * (xi - x[jE]) + w[jW] * (xi - x[jW]);
    s += w[n + i] * (xi - x[jN]) + wS * (xi - x[jS]);
    s += w[2 * n + i] * (xi - x[jNE]) + wSW * (xi - x[jSW]);
    s += w[3 * n + i] * (xi - x[jNW]) + w[3 * n + jSE] * (xi - x[jSE]);
    return s;
}

// ---- storage ------------------------------------------------------------------------------------
// Blocks of one scope, bump-allocated from a workspace arena (pfor.cuh) and handed back together when the scope
// ends: `temp` pools are scoped temporaries (stack discipline), the other pool lives as long as the solve.
struct Pool {
    Arena& ar;
#ifdef SSRS_HOST_EMU
    size_t mk;
#else
    Arena::Mark mk;
#endif
    size_t bytes = 0;
    explicit Pool(stream_t, bool temp = false) : ar(arena(temp)), mk(ar.mark()) {}
    template <class T> T* get(i64 count) {
        void* p = ar.alloc(sizeof(T) * (size_t)(count > 0 ? count : 1));
        if (p) bytes += sizeof(T) * (size_t)count;
        return static_cast<T*>(p);
    }
    void release(void*) {}          // reclaimed with the pool
    ~Pool() { ar.release_to(mk); }
};

struct Level {
    i64 n = 0, nnz = 0;
    i64* rowptr = nullptr; int* col = nullptr; double* val = nullptr;       // CSR (levels >= 1)
    int* agg = nullptr; i64 nc = 0; i64* memptr = nullptr; int* mem = nullptr;  // map to the next level
    i64* sptr = nullptr; int* ecol = nullptr; float* eval = nullptr; i64 ell_entries = 0;   // sliced ELL, float32 (levels >= 1)
    unsigned* epack = nullptr;     // packed entries (bf16 value | 16-bit column offset) when every offset fits; then ecol/eval are unused
    float *excess = nullptr, *dinv = nullptr;                                 // per-row operator data, float32 like the entries
    real *x32 = nullptr, *b32 = nullptr, *t32 = nullptr, *r32 = nullptr;      // cycle vectors (levels >= 1)
    // row-sharded solve: ownership of this level's nodes and, per part, the index range its rows reference
    Parts parts;
    i64 ref_lo[SSRS_MAX_RANKS] = {0}, ref_hi[SSRS_MAX_RANKS] = {0};             // [ref_lo, ref_hi) incl. the own range
};

#define AMG_TRY(expr) do { if ((expr) != 0) { set_error("ssrs_potential_solve: device operation failed: %s", #expr); return SSRS_ERR_CUDA; } } while (0)
#define AMG_RC(expr) do { const int rc_ = (expr); if (rc_) return rc_; } while (0)
#define AMG_ALLOC(var, T, count) do { var = pool.get<T>(count); if (!var) { set_error("ssrs_potential_solve: out of device memory (%lld x %zu B)", (long long)(count), sizeof(T)); return SSRS_ERR_CUDA; } } while (0)

// ---- distributed setup (row-sharded solve) --------------------------------------------------------------
// While a level is distributed every rank builds ITS rows of the hierarchy only: aggregates never straddle a slab
// boundary, so matching, numbering, member lists, Galerkin rows and the ELL slices of a part depend on that part's rows
// alone — plus the aggregate ids of the ghost rows (one halo exchange per level) and three tiny exchanges of per-part
// integers (aggregate counts, referenced index ranges, entry counts).  Per-row results are the same operations on the
// same operands as in the redundant setup, so the hierarchy — and with it every iterate — is bit-identical to it.
struct SetupDist {
    const ssrs_comm* comm = nullptr;   // nullptr: this level is built in full by every caller (one GPU, or a redundant level)
    int rank = 0, size = 1;
    stream_t st = nullptr;
    bool on() const { return comm != nullptr; }
    // v[q] for q != rank must be 0 on entry: on return every rank holds all parts' values (exact below 2^53)
    int gather(double* v, int count) const {
        if (comm == nullptr) return 0;
        if (comm->allreduce_sum(comm->ctx, v, count, (void*)st) != 0) { set_error("ssrs_potential_solve: all-reduce failed during setup"); return SSRS_ERR_CUDA; }
        return 0;
    }
};

// ---- coarsening: pairwise matching + joins ---------------------------------------------------------
template <class G>
int coarsen(const G g, Level& L, Pool& pool, double theta, int rounds, int join_rounds, Parts* next, stream_t st,
            const SetupDist D = SetupDist()) {
    const i64 n = g.size();
    // rows this caller works on: its own slab of a distributed level, else the whole level
    const i64 i0 = D.on() ? g.parts.lo[D.rank] : 0, i1 = D.on() ? g.parts.lo[D.rank + 1] : n;
    float* rowmax; int *mate, *best, *root, *root2;
    Pool tmp(st, true);
    rowmax = tmp.get<float>(n); mate = tmp.get<int>(n); best = tmp.get<int>(n); root = tmp.get<int>(n); root2 = tmp.get<int>(n);
    if (!rowmax || !mate || !best || !root || !root2) { set_error("ssrs_potential_solve: out of device memory in coarsen"); return SSRS_ERR_CUDA; }
    const float th = (float)theta;
    AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
        float m = 0.0f;
        if (!g.excluded(i)) {
            i64 k0, k1; g.range(i, k0, k1);
            for (i64 k = k0; k < k1; ++k) { i64 j; double a; if (g.entry(i, k, j, a) == E_OFF) { float w = (float)(-a); if (w > m) m = w; } }
        }
        rowmax[i] = m;
        mate[i] = g.excluded(i) ? -2 : -1;
    }));
    for (int rnd = 0; rnd < rounds; ++rnd) {
        AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
            int bj = -1;
            if (mate[i] == -1) {
                float bw = 0.0f; unsigned bh = 0;
                const float rmi = rowmax[i];
                i64 k0, k1; g.range(i, k0, k1);
                for (i64 k = k0; k < k1; ++k) {
                    i64 j; double a;
                    if (g.entry(i, k, j, a) != E_OFF) continue;
                    const float w = (float)(-a);
                    if (!(w > 0.0f) || mate[j] != -1 || g.crosses(i, j)) continue;   // aggregates never straddle a slab boundary
                    const float rmj = rowmax[j];
                    if (w < th * (rmi > rmj ? rmi : rmj)) continue;           // strong from both sides
                    const unsigned h = pair_hash((unsigned)i, (unsigned)j);
                    if (bj < 0 || w > bw || (w == bw && (h > bh || (h == bh && (int)j > bj)))) { bw = w; bh = h; bj = (int)j; }
                }
            }
            best[i] = bj;
        }));
        AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
            const int b = best[i];
            if (b >= 0 && best[b] == (int)i) mate[i] = b;
        }));
    }
    AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
        const int m = mate[i];
        root[i] = m >= 0 ? ((int)i < m ? (int)i : m) : (m == -2 ? -2 : -1);
    }));
    for (int jr = 0; jr < join_rounds; ++jr) {
        AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
            int out = root[i];
            if (out == -1) {
                int bj = -1, btwo = 0; float bw = 0.0f; unsigned bh = 0;
                const float rmi = rowmax[i];
                i64 k0, k1; g.range(i, k0, k1);
                for (i64 k = k0; k < k1; ++k) {
                    i64 j; double a;
                    if (g.entry(i, k, j, a) != E_OFF) continue;
                    const float w = (float)(-a);
                    if (!(w > 0.0f) || root[j] < 0 || w < th * rmi || g.crosses(i, j)) continue;  // strong for i, target already aggregated
                    const int two = (w >= th * rowmax[j]) ? 1 : 0;
                    const unsigned h = pair_hash((unsigned)i, (unsigned)j);
                    if (bj < 0 || two > btwo || (two == btwo && (w > bw || (w == bw && (h > bh || (h == bh && (int)j > bj)))))) {
                        btwo = two; bw = w; bh = h; bj = (int)j;
                    }
                }
                if (bj >= 0) out = root[bj];
            }
            root2[i] = out;
        }));
        int* sw = root; root = root2; root2 = sw;
    }
    // coarse numbering in root order
    i64* flag = tmp.get<i64>(n + 1);
    if (!flag) { set_error("ssrs_potential_solve: out of device memory in coarsen"); return SSRS_ERR_CUDA; }
    {
        int* rt = root;
        AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
            if (rt[i] == -1) rt[i] = (int)i;                  // singleton
            flag[i] = (rt[i] == (int)i) ? 1 : 0;
        }));
    }
    i64 nc = 0, clo = 0, chi = 0;      // coarse nodes in total; this caller's range of them
    next->n = g.parts.n;
    if (D.on()) {
        // number the own slab's roots, then shift by the aggregates of the parts before it
        i64 nc_local = 0;
        AMG_TRY(exclusive_scan_i64(flag + i0, i1 - i0, &nc_local, st));
        double cnts[SSRS_MAX_RANKS] = {0.0};
        cnts[D.rank] = (double)nc_local;
        AMG_RC(D.gather(cnts, D.size));
        next->lo[0] = 0;
        for (int p = 0; p < D.size; ++p) next->lo[p + 1] = next->lo[p] + (i64)cnts[p];
        nc = next->lo[D.size]; clo = next->lo[D.rank]; chi = next->lo[D.rank + 1];
        if (clo > 0) { const i64 base = clo; AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) { flag[i] += base; })); }
    } else {
        AMG_TRY(exclusive_scan_i64(flag, n, &nc, st));
        // ownership of the coarse nodes: numbering follows the roots, so every part's aggregates are contiguous
        next->lo[0] = 0;
        next->lo[next->n] = nc;
        for (int p = 1; p < g.parts.n; ++p) {
            if (g.parts.lo[p] >= n) next->lo[p] = nc;
            else AMG_TRY(copy_d2h(&next->lo[p], flag + g.parts.lo[p], sizeof(i64), st));
        }
        chi = nc;
    }
    AMG_ALLOC(L.agg, int, n);
    AMG_ALLOC(L.memptr, i64, nc + 1);
    AMG_ALLOC(L.mem, int, n);
    L.nc = nc;
    int* agg = L.agg; i64* memptr = L.memptr; int* mem = L.mem;
    int* cnt = tmp.get<int>(nc + 1);
    if (!cnt) { set_error("ssrs_potential_solve: out of device memory in coarsen"); return SSRS_ERR_CUDA; }
    AMG_TRY(dev_zero(cnt, sizeof(int) * (size_t)(nc + 1), st));
    {
        const int* rt = root;
        AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
            const int r = rt[i];
            const int c = r >= 0 ? (int)flag[r] : -1;
            agg[i] = c;
            if (c >= 0) atomic_add_int(cnt + c, 1);
        }));
    }
    // member lists of the own aggregates (offsets local to this caller's `mem`)
    AMG_TRY(pfor_range(clo, chi + 1, st, [=] SSRS_HD(i64 I) { memptr[I] = I < chi ? (i64)cnt[I] : 0; }));
    i64 total = 0;
    AMG_TRY(exclusive_scan_i64(memptr + clo, chi - clo + 1, &total, st));
    AMG_TRY(dev_zero(cnt, sizeof(int) * (size_t)(nc + 1), st));
    AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
        const int c = agg[i];
        if (c >= 0) { const int slot = atomic_add_int(cnt + c, 1); mem[memptr[c] + slot] = (int)i; }
    }));
    // members in ascending order -> deterministic Galerkin sums
    AMG_TRY(pfor_range(clo, chi, st, [=] SSRS_HD(i64 I) {
        const i64 a = memptr[I], b = memptr[I + 1];
        for (i64 p = a + 1; p < b; ++p) {
            const int v = mem[p];
            i64 q = p - 1;
            while (q >= a && mem[q] > v) { mem[q + 1] = mem[q]; --q; }
            mem[q + 1] = v;
        }
    }));
    AMG_TRY(sync(st));
    return SSRS_OK;
}

// ---- Galerkin coarse operator for piecewise-constant prolongation ---------------------------------------
// [clo, chi): the coarse rows this caller builds (its own aggregates of a distributed level — `agg` must then hold the
// ghost rows' ids too —, else all of them); C.rowptr / col / val are local to the caller in the distributed case.
template <class G>
int galerkin(const G g, const Level& L, Level& C, Pool& pool, stream_t st, i64 clo = 0, i64 chi = -1) {
    const i64 nc = L.nc;
    if (chi < 0) chi = nc;
    const int* agg = L.agg; const i64* memptr = L.memptr; const int* mem = L.mem;
    Pool tmp(st, true);
    i64* off = tmp.get<i64>(nc + 1);
    int* len = tmp.get<int>(nc);
    if (!off || !len) { set_error("ssrs_potential_solve: out of device memory in galerkin"); return SSRS_ERR_CUDA; }
    AMG_TRY(pfor_range(clo, chi + 1, st, [=] SSRS_HD(i64 I) {
        i64 ub = 0;
        if (I < chi)
            for (i64 p = memptr[I]; p < memptr[I + 1]; ++p) { i64 k0, k1; g.range(mem[p], k0, k1); ub += (k1 - k0) + 1; }
        off[I] = ub;
    }));
    i64 scratch_n = 0;
    AMG_TRY(exclusive_scan_i64(off + clo, chi - clo + 1, &scratch_n, st));
    int* scol = tmp.get<int>(scratch_n);
    double* sval = tmp.get<double>(scratch_n);
    if (!scol || !sval) { set_error("ssrs_potential_solve: out of device memory in galerkin (%lld entries)", (long long)scratch_n); return SSRS_ERR_CUDA; }
    AMG_TRY(pfor_range(clo, chi, st, [=] SSRS_HD(i64 I) {
        int* cj = scol + off[I];
        double* cv = sval + off[I];
        int cnt = 0;
        for (i64 p = memptr[I]; p < memptr[I + 1]; ++p) {
            const i64 i = mem[p];
            i64 k0, k1; g.range(i, k0, k1);
            double diag = 0.0, linksum = 0.0; bool have_diag = false;
            for (i64 k = k0; k <= k1; ++k) {
                int J; double a;
                if (k < k1) {
                    i64 j;
                    const int kind = g.entry(i, k, j, a);
                    if (kind == E_DIAG) { diag = a; have_diag = true; continue; }
                    if (kind == E_DIR) { linksum += a; continue; }
                    if (kind != E_OFF) continue;
                    linksum += a;
                    J = agg[j];
                    if (J < 0) continue;
                } else {                                  // the diagonal entry goes last
                    J = (int)I;
                    a = have_diag ? diag : -linksum;
                }
                int q = 0;
                while (q < cnt && cj[q] != J) ++q;
                if (q == cnt) { cj[cnt] = J; cv[cnt] = a; ++cnt; }
                else cv[q] += a;
            }
        }
        len[I] = cnt;
    }));
    AMG_ALLOC(C.rowptr, i64, nc + 1);
    i64* rowptr = C.rowptr;
    AMG_TRY(pfor_range(clo, chi + 1, st, [=] SSRS_HD(i64 I) { rowptr[I] = I < chi ? (i64)len[I] : 0; }));
    i64 nnz = 0;
    AMG_TRY(exclusive_scan_i64(rowptr + clo, chi - clo + 1, &nnz, st));
    AMG_ALLOC(C.col, int, nnz);
    AMG_ALLOC(C.val, double, nnz);
    int* col = C.col; double* val = C.val;
    AMG_TRY(pfor_range(clo, chi, st, [=] SSRS_HD(i64 I) {
        const i64 s = off[I], d = rowptr[I];
        for (int q = 0; q < len[I]; ++q) { col[d + q] = scol[s + q]; val[d + q] = sval[s + q]; }
    }));
    C.n = nc; C.nnz = nnz;
    AMG_TRY(sync(st));
    return SSRS_OK;
}

// ---- the cycle: float32 operators, float64 vectors ------------------------------------------------------
// The V-cycle is only a preconditioner, so its operators are stored in reduced precision (fine level: 4 forward link
// weights per cell as bfloat16 pairs; coarse levels: sliced ELL with bfloat16 values; diagonals and row anchoring in
// float32), computed in float64 and rounded once.  Its VECTORS stay float64:
// with conductances spanning 1e-8..1 a conducting island is anchored to the rest of the grid by links 1e-8 of
// its internal ones, so float32 rounding noise of an island's (large, nearly constant) correction, multiplied by
// the strong internal links in the outer A*y, swamps the signal carried by the weak links — a float32-vector
// cycle (-DSSRS_CYCLE_F32, experiment) diverges on the 500 x 600 test grid.  Operators are applied in
// difference form, (A x)_i = excess_i x_i + sum_j a_ij (x_j - x_i), with the row excess (the anchoring to the
// Dirichlet set) formed in float64.  The outer BiCGStab iteration evaluates the exact float64 operator, so
// rounding inside the cycle cannot change the solution it converges to.
// bfloat16 pairs in one word: .lo() = low half, .hi() = high half, each the upper 16 bits of a float32
struct Bf16Pairs {
    const unsigned* p;
    SSRS_HD float lo(i64 i) const { const unsigned u = p[i] << 16; float v; memcpy(&v, &u, 4); return v; }
    SSRS_HD float hi(i64 i) const { const unsigned u = p[i] & 0xFFFF0000u; float v; memcpy(&v, &u, 4); return v; }
};
SSRS_HD inline unsigned bf16_bits(float v) {               // round to nearest even
    unsigned u;
    memcpy(&u, &v, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return u >> 16;
}
// Forward link weights of the cycle's fine-level operator, 0 where the neighbour is outside the grid, as bfloat16
// pairs {E, N} and {NE, NW}: 8 bytes per cell instead of 16 in passes that are bound by their bytes, and six weight
// loads per cell instead of eight.  Like the coarse levels' packed entries this only perturbs a preconditioner (the
// difference form keeps the constant mode exact; the diagonal `dinv` is formed from the float32 weights).
struct Fine32W {
    Bf16Pairs en, dd;
    SSRS_HD float E(i64 i) const { return en.lo(i); }
    SSRS_HD float N(i64 i) const { return en.hi(i); }
    SSRS_HD float NE(i64 i) const { return dd.lo(i); }
    SSRS_HD float NW(i64 i) const { return dd.hi(i); }
};
struct Fine32 {
    Fine32W w;
    const float* dinv;                 // 1 / (sum of the eight link weights); 0 at Dirichlet nodes
    const float* kd;                    // conductivity, sign bit = Dirichlet
    int rows, cols;
};

// (A e)_i of the error equation (e = 0 at Dirichlet nodes) for cell (r, c); x(j) returns e_j
template <class X>
SSRS_HD inline real fine_apply32(const Fine32& F, int r, int c, const X& x) {
    const int cols = F.cols;
    const i64 i = (i64)r * cols + c;
    if (r > 0 && r < F.rows - 1 && c > 0 && c < cols - 1) {
        // interior cell (all but the raster's border; the quirk column is on the border): fixed offsets, same order
        const real xi = x(i);
        const Fine32W& w = F.w;
        real s = w.E(i) * (xi - x(i + 1)) + w.E(i - 1) * (xi - x(i - 1));
        s += w.N(i) * (xi - x(i + cols)) + w.N(i - cols) * (xi - x(i - cols));
        s += w.NE(i) * (xi - x(i + cols + 1)) + w.NE(i - cols - 1) * (xi - x(i - cols - 1));
        s += w.NW(i) * (xi - x(i + cols - 1)) + w.NW(i - cols + 1) * (xi - x(i - cols + 1));
        return s;
    }
    const bool hW = c > 0, hE = c < cols - 1, hS = r > 0, hN = r < F.rows - 1;
    // a missing neighbour aliases the cell itself: its difference is exactly zero whatever weight is read
    const i64 jE = hE ? i + 1 : i, jW = hW ? i - 1 : i, jN = hN ? i + cols : i, jS = hS ? i - cols : i;
    const i64 jNE = (hN && hE) ? i + cols + 1 : i, jNW = (hN && hW) ? i + cols - 1 : i;
    const i64 jSW = (hS && hW) ? i - cols - 1 : i, jSE = (hS && hE) ? i - cols + 1 : i;
    const Fine32W& w = F.w;
    real wS = w.N(jS), wSW = w.NE(jSW);
    if (c == cols - 1 && hS && hN) {                                        // movmodel.py:73-79
        wS = (real)link_weight<false>(F.kd[i], F.kd[jS], true);
        wSW = (real)link_weight<false>(F.kd[i], F.kd[jSW], false);
    }
    const real xi = x(i);
    real s = w.E(i) * (xi - x(jE)) + w.E(jW) * (xi - x(jW));
    s += w.N(i) * (xi - x(jN)) + wS * (xi - x(jS));
    s += w.NE(i) * (xi - x(jNE)) + wSW * (xi - x(jSW));
    s += w.NW(i) * (xi - x(jNW)) + w.NW(jSE) * (xi - x(jSE));
    return s;
}

struct VecF { const real* p; SSRS_HD real operator()(i64 j) const { return p[j]; } };
// x1 = omega D^-1 b, the first Jacobi sweep from a zero guess, evaluated on the fly
struct FirstSweepFine { const double* b; const float* dinv; real omega; SSRS_HD real operator()(i64 j) const { return omega * dinv[j] * (real)b[j]; } };
struct FirstSweepF { const real* b; const float* dinv; real omega; SSRS_HD real operator()(i64 j) const { return omega * dinv[j] * b[j]; } };

// All cycle kernels take the range they compute ([r0, r1) rows of the fine grid, [i0, i1) nodes of a coarse
// level): the whole level on one GPU, the owned slab in the row-sharded solve.
inline int fine_first32(const Fine32 F, int r0, int r1, const double* b, real* x, real omega, stream_t st) {
    return pfor2d_rows(r0, r1, F.cols, st, [=] SSRS_HD(int r, int c) {
        const i64 i = (i64)r * F.cols + c;
        x[i] = omega * F.dinv[i] * (real)b[i];
    });
}
// (No early exit on Dirichlet cells in these kernels: a branch on a loaded value in front of the stencil would put
// one more memory latency on every cell's critical path.  The stencil is safe to evaluate there; the result is
// discarded by a select.)
// x = first sweep, res = b - A x in one pass over b
inline int fine_first_residual32(const Fine32 F, int r0, int r1, const double* b, real* x, real* res, real omega, stream_t st) {
    return pfor2d_rows(r0, r1, F.cols, st, [=] SSRS_HD(int r, int c) {
        const i64 i = (i64)r * F.cols + c;
        const FirstSweepFine x1 = {b, F.dinv, omega};
        const bool free_node = F.dinv[i] != 0.0f;
        const real bi = (real)b[i];
        const real ax = fine_apply32(F, r, c, x1);
        x[i] = x1(i);                                   // 0 at Dirichlet cells (dinv = 0)
        res[i] = free_node ? bi - ax : (real)0.0;
    });
}
inline int fine_residual32(const Fine32 F, int r0, int r1, const double* b, const real* x, real* res, stream_t st) {
    return pfor2d_rows(r0, r1, F.cols, st, [=] SSRS_HD(int r, int c) {
        const i64 i = (i64)r * F.cols + c;
        const bool free_node = F.dinv[i] != 0.0f;
        const real bi = (real)b[i];
        const real ax = fine_apply32(F, r, c, VecF{x});
        res[i] = free_node ? bi - ax : (real)0.0;
    });
}
template <class OutT>
inline int fine_jacobi32(const Fine32 F, int r0, int r1, const double* b, const real* x, OutT* xn, real omega, stream_t st) {
    return pfor2d_rows(r0, r1, F.cols, st, [=] SSRS_HD(int r, int c) {
        const i64 i = (i64)r * F.cols + c;
        const real di = F.dinv[i];
        const real bi = (real)b[i];
        const real ax = fine_apply32(F, r, c, VecF{x});
        xn[i] = (di == (real)0.0) ? (OutT)0 : (OutT)(x[i] + omega * di * (bi - ax));
    });
}

// Coarse operators: sliced ELL (slices of 32 rows, column-major inside a slice so that a warp reads
// consecutive entries), off-diagonals only; padding entries point at the row itself with value 0.
struct Ell {
    const i64* sptr;        // [slices + 1]
    const int* col;
    const float* val;
    const unsigned* pack;   // packed entries, or nullptr: then col/val hold the entries
    const float* excess;   // a_ii + sum_off a_ij (formed in float64, rounded once)
    const float* dinv;     // 1 / a_ii
    i64 n;
};
// Packed entry: high half = the value rounded to bfloat16 (the upper 16 bits of its float32), low half = column - row
// as a signed 16-bit offset.  Half the bytes of {int32 column, float32 value}; the sweeps over the coarse levels are
// bound by exactly these bytes.  The 8 significant bits only perturb the off-diagonal weights of a preconditioner:
// the row's anchoring (`excess`, formed from the exact sums) and the diagonal stay float32, and the difference form
// keeps the constant mode exact whatever the weights are.
SSRS_HD inline unsigned pack_entry(float v, i64 delta) {
    unsigned u;
    memcpy(&u, &v, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);                       // round to nearest even at bit 16
    return (u & 0xFFFF0000u) | ((unsigned)(int)delta & 0xFFFFu);
}
SSRS_HD inline float packed_value(unsigned w) {
    const unsigned u = w & 0xFFFF0000u;
    float v;
    memcpy(&v, &u, 4);
    return v;
}
SSRS_HD inline i64 packed_col(unsigned w, i64 i) { return i + (i64)(short)(w & 0xFFFFu); }

template <bool PACKED, class X>
SSRS_HD inline real ell_apply_fmt(const Ell& e, i64 i, const X& x) {
    const i64 s = i >> 5;
    const i64 p1 = e.sptr[s + 1];
    const real xi = x(i);
    real acc = (real)0.0;
    i64 p = e.sptr[s] + (i & 31);
    // eight, then four entries per trip: the entries, then the gathers they address, are independent loads — a row's
    // entries are otherwise a chain of dependent L2 round trips (the small levels are pure latency).  Rows have ~10
    // entries, so most rows cost two entry/gather round trips.
    if (PACKED) {
        for (; p + 224 < p1; p += 256) {
            unsigned w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) w[q] = e.pack[p + 32 * q];
            real xs[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) xs[q] = x(packed_col(w[q], i));
            acc += ((packed_value(w[0]) * (xs[0] - xi) + packed_value(w[1]) * (xs[1] - xi)) +
                    (packed_value(w[2]) * (xs[2] - xi) + packed_value(w[3]) * (xs[3] - xi))) +
                   ((packed_value(w[4]) * (xs[4] - xi) + packed_value(w[5]) * (xs[5] - xi)) +
                    (packed_value(w[6]) * (xs[6] - xi) + packed_value(w[7]) * (xs[7] - xi)));
        }
    }
    for (; p + 96 < p1; p += 128) {
        i64 c0, c1, c2, c3;
        float v0, v1, v2, v3;
        if (PACKED) {
            const unsigned w0 = e.pack[p], w1 = e.pack[p + 32], w2 = e.pack[p + 64], w3 = e.pack[p + 96];
            c0 = packed_col(w0, i); c1 = packed_col(w1, i); c2 = packed_col(w2, i); c3 = packed_col(w3, i);
            v0 = packed_value(w0); v1 = packed_value(w1); v2 = packed_value(w2); v3 = packed_value(w3);
        } else {
            c0 = e.col[p]; c1 = e.col[p + 32]; c2 = e.col[p + 64]; c3 = e.col[p + 96];
            v0 = e.val[p]; v1 = e.val[p + 32]; v2 = e.val[p + 64]; v3 = e.val[p + 96];
        }
        const real x0 = x(c0), x1 = x(c1), x2 = x(c2), x3 = x(c3);
        acc += (v0 * (x0 - xi) + v1 * (x1 - xi)) + (v2 * (x2 - xi) + v3 * (x3 - xi));
    }
    for (; p < p1; p += 32) {
        if (PACKED) { const unsigned w = e.pack[p]; acc += packed_value(w) * (x(packed_col(w, i)) - xi); }
        else acc += e.val[p] * (x(e.col[p]) - xi);
    }
    return e.excess[i] * xi + acc;
}
template <class X>
SSRS_HD inline real ell_apply(const Ell& e, i64 i, const X& x) {
    return e.pack != nullptr ? ell_apply_fmt<true>(e, i, x) : ell_apply_fmt<false>(e, i, x);
}
inline int ell_first32(const Ell e, i64 i0, i64 i1, const real* b, real* x, real omega, stream_t st) {
    return pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) { x[i] = omega * e.dinv[i] * b[i]; });
}
inline int ell_first_residual32(const Ell e, i64 i0, i64 i1, const real* b, real* x, real* res, real omega, stream_t st) {
    return pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
        const FirstSweepF x1 = {b, e.dinv, omega};
        x[i] = x1(i);
        res[i] = b[i] - ell_apply(e, i, x1);
    });
}
inline int ell_residual32(const Ell e, i64 i0, i64 i1, const real* b, const real* x, real* res, stream_t st) {
    return pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) { res[i] = b[i] - ell_apply(e, i, VecF{x}); });
}
inline int ell_jacobi32(const Ell e, i64 i0, i64 i1, const real* b, const real* x, real* xn, real omega, stream_t st) {
    return pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) { xn[i] = x[i] + omega * e.dinv[i] * (b[i] - ell_apply(e, i, VecF{x})); });
}
// Small levels (<= SMALL_ROWS rows): rows are long (tens to hundreds of entries) and few, so a warp walks each row
// of the CSR arrays instead of a thread walking an ELL row: same operator (float32-rounded entries, difference form),
// a handful of coalesced trips instead of a chain of dependent gathers.
constexpr i64 SMALL_ROWS = 32768;
static bool g_ell_packed = true;     // SSRS_X_ELLPACK=0 keeps {int32, float32} entries (A/B measurements)
SSRS_HD inline real csr_row_partial(const CsrGraph& g, i64 i, int lane, const real* x) {
    const real xi = x[i];
    real acc = (real)0.0;
    for (i64 k = g.rowptr[i] + lane; k < g.rowptr[i + 1]; k += 32) {
        const i64 j = g.col[k];
        if (j != i) acc += (real)(float)g.val[k] * (x[j] - xi);
    }
    return acc;
}
inline int csr_residual32(const CsrGraph g, const float* excess, i64 i0, i64 i1, const real* b, const real* x, real* res, stream_t st) {
    return pfor_warp_rows(i0, i1, st, [=] SSRS_HD(i64 i, int lane) { return (double)csr_row_partial(g, i, lane, x); },
                          [=] SSRS_HD(i64 i, double total) { res[i] = b[i] - (excess[i] * x[i] + (real)total); });
}
inline int csr_jacobi32(const CsrGraph g, const float* excess, const float* dinv, i64 i0, i64 i1, const real* b, const real* x,
                        real* xn, real omega, stream_t st) {
    return pfor_warp_rows(i0, i1, st, [=] SSRS_HD(i64 i, int lane) { return (double)csr_row_partial(g, i, lane, x); },
                          [=] SSRS_HD(i64 i, double total) { xn[i] = x[i] + omega * dinv[i] * (b[i] - (excess[i] * x[i] + (real)total)); });
}

// bc_I = sum of the residual over the members of aggregate I (members ascending: deterministic), I in [I0, I1);
// with x1 != nullptr also the next level's first sweep from a zero guess, x1_I = omega dinv_I bc_I
inline int restrict32(const Level& L, i64 I0, i64 I1, const real* res, real* bc, real* x1, const float* dinv, real omega, stream_t st) {
    const i64* memptr = L.memptr; const int* mem = L.mem;
    return pfor_range(I0, I1, st, [=] SSRS_HD(i64 I) {
        real s = (real)0.0;
        for (i64 p = memptr[I]; p < memptr[I + 1]; ++p) s += res[mem[p]];
        bc[I] = s;
        if (x1 != nullptr) x1[I] = omega * dinv[I] * s;
    });
}
inline int prolong_add32(const Level& L, i64 i0, i64 i1, real* x, const real* xc, real scale, stream_t st) {
    const int* agg = L.agg;
    return pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) { const int c = agg[i]; if (c >= 0) x[i] += scale * xc[c]; });
}

struct Hierarchy {
    FineGraph fine;
    std::vector<Level> lv;     // lv[0] = fine level (agg/mem only), lv[l>=1] CSR + ELL
    double* cinv = nullptr;    // dense inverse of the coarsest operator (float64, row-major)
    i64 cn = 0;
    int coarse_sweeps = 0;     // > 0: coarsest level too large for a dense inverse, Jacobi sweeps instead
    real omega = (real)0.8;        // Jacobi weight
    int nu = 1;                // pre- and post-smoothing sweeps
    int nu_coarse = 1, nu_from = 1 << 30;   // experiment: nu_coarse sweeps on levels >= nu_from
    real overcorrect = (real)1.2;  // scale of the coarse-grid correction (plain aggregation under-corrects smooth error)
    FineWeights fw = {nullptr, nullptr};
    Fine32 f32;
    real *xf = nullptr, *tf = nullptr, *resf = nullptr;   // fine-level cycle vectors
    // row-sharded solve (comm == nullptr: one GPU, every range is the whole level)
    const ssrs_comm* comm = nullptr;
    int rank = 0, nparts = 1;
    int lrep = 1 << 30;        // levels >= lrep are computed redundantly by every rank
    bool fuse_coarse_first = false;   // coarse levels: first sweep fused into the residual pass (two gathers per entry) or separate (one)
    stream_t st = nullptr;
    // Levels >= lg (small, computed in full by this rank, no communication) run as ONE CUDA graph per cycle: their ~45
    // launches of a few microseconds each are pure launch latency — 0.45 ms of a cycle, a fifth of an iteration of the
    // 8-GPU solve, where the large levels' work is divided by eight and these are not.
    int lg = 1 << 30;
    void* gexec = nullptr;            // cudaGraphExec_t of the coarse part (device build)
};

inline CsrGraph csr_of(const Level& L) { CsrGraph g; g.rowptr = L.rowptr; g.col = L.col; g.val = L.val; g.n = L.n; g.parts = L.parts; return g; }
inline Ell ell_of(const Level& L) { Ell e; e.sptr = L.sptr; e.col = L.ecol; e.val = L.eval; e.pack = L.epack; e.excess = L.excess; e.dinv = L.dinv; e.n = L.n; return e; }

inline int level_residual(const Level& L, i64 i0, i64 i1, const real* b, const real* x, real* res, stream_t st) {
    if (L.n <= SMALL_ROWS) return csr_residual32(csr_of(L), L.excess, i0, i1, b, x, res, st);
    return ell_residual32(ell_of(L), i0, i1, b, x, res, st);
}
inline int level_jacobi(const Level& L, i64 i0, i64 i1, const real* b, const real* x, real* xn, real omega, stream_t st) {
    if (L.n <= SMALL_ROWS) return csr_jacobi32(csr_of(L), L.excess, L.dinv, i0, i1, b, x, xn, omega, st);
    return ell_jacobi32(ell_of(L), i0, i1, b, x, xn, omega, st);
}

// range of level l this rank computes
inline void own_range(const Hierarchy& H, int l, i64& i0, i64& i1) {
    const Level& L = H.lv[(size_t)l];
    if (H.comm == nullptr || l >= H.lrep) { i0 = 0; i1 = L.n; }
    else { i0 = L.parts.lo[H.rank]; i1 = L.parts.lo[H.rank + 1]; }
}
// Refreshes the ghost entries of a level-l vector (elements of `esize` bytes) from the two neighbouring ranks:
// rank r receives [ref_lo[r], lo[r]) from r-1 and [lo[r+1], ref_hi[r]) from r+1, and sends what they reference.
int exchange_ghosts(const Hierarchy& H, int l, void* vec, size_t esize) {
    if (H.comm == nullptr || l >= H.lrep) return SSRS_OK;
    const Level& L = H.lv[(size_t)l];
    const int r = H.rank, R = H.nparts;
    const i64 lo = L.parts.lo[r], hi = L.parts.lo[r + 1];
    i64 su_off = 0, su_n = 0, ru_off = 0, ru_n = 0, sd_off = 0, sd_n = 0, rd_off = 0, rd_n = 0;
    if (r > 0) {
        ru_off = L.ref_lo[r]; ru_n = lo - L.ref_lo[r];                  // my upper ghost zone
        su_off = lo; su_n = L.ref_hi[r - 1] - lo;                       // what rank r-1 references beyond its range
    }
    if (r + 1 < R) {
        rd_off = hi; rd_n = L.ref_hi[r] - hi;
        sd_off = L.ref_lo[r + 1]; sd_n = hi - L.ref_lo[r + 1];
    }
    const i64 e = (i64)esize;
    if (H.comm->exchange(H.comm->ctx, vec, su_off * e, su_n * e, ru_off * e, ru_n * e, sd_off * e, sd_n * e, rd_off * e, rd_n * e, (void*)H.st) != 0) {
        set_error("ssrs_potential_solve: halo exchange failed on level %d", l);
        return SSRS_ERR_CUDA;
    }
    return SSRS_OK;
}

// CSR (float64, diagonal stored) -> sliced ELL (float32).  With D.on() (a distributed level during setup) only the
// caller's own rows are laid out — slice offsets are local to the caller, a slice that straddles a part boundary is
// sized by the caller's rows in it — and the per-part quantities every rank needs are exchanged: whether the packed
// entry format applies (decided for the level as a whole, as in the redundant setup) and the index range each part's
// rows reference.
int build_ell(Level& L, Pool& pool, stream_t st, const SetupDist D = SetupDist()) {
    const CsrGraph g = csr_of(L);
    const i64 n = L.n, slices = (n + 31) / 32;
    const i64 i0 = D.on() ? L.parts.lo[D.rank] : 0, i1 = D.on() ? L.parts.lo[D.rank + 1] : n;
    const i64 s0 = i0 / 32, s1 = i1 > i0 ? (i1 + 31) / 32 : s0;          // slices that hold own rows
    AMG_ALLOC(L.sptr, i64, slices + 1);
    AMG_ALLOC(L.excess, float, n);
    AMG_ALLOC(L.dinv, float, n);
    AMG_ALLOC(L.x32, real, n);
    AMG_ALLOC(L.b32, real, n);
    AMG_ALLOC(L.t32, real, n);
    AMG_ALLOC(L.r32, real, n);
    i64* sptr = L.sptr; float* excess = L.excess; float* dinv = L.dinv;
    AMG_TRY(pfor_range(s0, s1 + 1, st, [=] SSRS_HD(i64 s) {
        i64 width = 0;
        if (s < s1)
            for (i64 i = s * 32; i < s * 32 + 32 && i < n; ++i) {
                if (i < i0 || i >= i1) continue;
                i64 len = 0;
                for (i64 k = g.rowptr[i]; k < g.rowptr[i + 1]; ++k) len += (g.col[k] != (int)i);
                if (len > width) width = len;
            }
        sptr[s] = width * 32;
    }));
    if (D.on() && i1 - i0 >= 64) {
        // a slice shared with a neighbour gets the width the whole slice has (the order in which a row's entries are
        // summed depends on it, and that order is to be the redundant setup's)
        const bool sh_lo = (i0 & 31) != 0, sh_hi = (i1 & 31) != 0 && i1 < n;
        i64 wl = 0, wh = 0;
        if (sh_lo) AMG_TRY(copy_d2h(&wl, sptr + s0, sizeof(i64), st));
        if (sh_hi) AMG_TRY(copy_d2h(&wh, sptr + s1 - 1, sizeof(i64), st));
        AMG_TRY(sync(st));
        double v[2 * SSRS_MAX_RANKS] = {0.0};
        v[2 * D.rank] = (double)wl; v[2 * D.rank + 1] = (double)wh;
        AMG_RC(D.gather(v, 2 * D.size));
        if (sh_lo && D.rank > 0 && (i64)v[2 * (D.rank - 1) + 1] > wl) { wl = (i64)v[2 * (D.rank - 1) + 1]; AMG_TRY(copy_h2d(sptr + s0, &wl, sizeof(i64), st)); }
        if (sh_hi && D.rank + 1 < D.size && (i64)v[2 * (D.rank + 1)] > wh) { wh = (i64)v[2 * (D.rank + 1)]; AMG_TRY(copy_h2d(sptr + s1 - 1, &wh, sizeof(i64), st)); }
        AMG_TRY(sync(st));
    } else if (D.on()) {
        double v[2 * SSRS_MAX_RANKS] = {0.0};        // keep the collective call sequence identical on every rank
        AMG_RC(D.gather(v, 2 * D.size));
    }
    i64 total = 0;
    AMG_TRY(exclusive_scan_i64(sptr + s0, s1 - s0 + 1, &total, st));
    // packed 4-byte entries when every column lies within +-32767 of its row (node numbering follows the raster, so
    // this holds unless a raster row has more than ~1e5 cells); else {int32 column, float32 value}
    double far_entries = 0.0;
    AMG_TRY(preduce_sum_range(i0, i1, st, &far_entries, [=] SSRS_HD(i64 i) {
        double cnt = 0.0;
        for (i64 k = g.rowptr[i]; k < g.rowptr[i + 1]; ++k) { const i64 dlt = (i64)g.col[k] - i; cnt += (dlt > 32767 || dlt < -32767) ? 1.0 : 0.0; }
        return cnt;
    }));
    AMG_RC(D.gather(&far_entries, 1));
    const bool packed = g_ell_packed && far_entries == 0.0;
    L.ell_entries = total;
    if (packed) { AMG_ALLOC(L.epack, unsigned, total); }
    else { AMG_ALLOC(L.ecol, int, total); AMG_ALLOC(L.eval, float, total); }
    int* ecol = L.ecol; float* eval = L.eval; unsigned* epack = L.epack;
    AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
        const i64 s = i >> 5;
        i64 p = sptr[s] + (i & 31);
        const i64 p1 = sptr[s + 1];
        double d = 0.0, off = 0.0;
        for (i64 k = g.rowptr[i]; k < g.rowptr[i + 1]; ++k) {
            if (g.col[k] == (int)i) { d = g.val[k]; continue; }
            off += g.val[k];
            if (packed) epack[p] = pack_entry((float)g.val[k], (i64)g.col[k] - i);
            else { ecol[p] = g.col[k]; eval[p] = (float)g.val[k]; }
            p += 32;
        }
        for (; p < p1; p += 32) {
            if (packed) epack[p] = 0u;
            else { ecol[p] = (int)i; eval[p] = 0.0f; }
        }
        excess[i] = (float)(d + off);
        dinv[i] = (float)(1.0 / d);
    }));
    if (D.on() && s1 > s0) {
        // rows of a straddling slice that belong to the neighbours: their lanes must read as empty entries
        const i64 a0 = s0 * 32, a1 = (s1 * 32 < n ? s1 * 32 : n);
        AMG_TRY(pfor_range(a0, a1, st, [=] SSRS_HD(i64 i) {
            if (i >= i0 && i < i1) return;
            const i64 s = i >> 5;
            for (i64 p = sptr[s] + (i & 31); p < sptr[s + 1]; p += 32) {
                if (packed) epack[p] = 0u;
                else { ecol[p] = (int)i; eval[p] = 0.0f; }
            }
        }));
    }
    // row-sharded solve: the index range each part's rows reference (its own range plus the ghost zones)
    const Parts P = L.parts;
    for (int q = 0; q < P.n; ++q) { L.ref_lo[q] = P.lo[q]; L.ref_hi[q] = P.lo[q + 1]; }
    if (P.n > 1) {
        Pool tmp(st, true);
        i64* ref = tmp.get<i64>(2 * P.n);
        if (!ref) { set_error("ssrs_potential_solve: out of device memory in build_ell"); return SSRS_ERR_CUDA; }
        i64 init[2 * SSRS_MAX_RANKS];
        for (int q = 0; q < P.n; ++q) { init[q] = L.ref_lo[q]; init[P.n + q] = L.ref_hi[q]; }
        AMG_TRY(copy_h2d(ref, init, sizeof(i64) * 2 * (size_t)P.n, st));
        AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) {
            i64 lo = i, hi = i;
            for (i64 k = g.rowptr[i]; k < g.rowptr[i + 1]; ++k) { const i64 j = g.col[k]; lo = j < lo ? j : lo; hi = j > hi ? j : hi; }
            const int q = parts_owner(P, i);
            if (lo < P.lo[q]) atomic_min_i64(ref + q, lo);
            if (hi >= P.lo[q + 1]) atomic_max_i64(ref + P.n + q, hi + 1);
        }));
        AMG_TRY(copy_d2h(init, ref, sizeof(i64) * 2 * (size_t)P.n, st));
        AMG_TRY(sync(st));
        if (D.on()) {            // every rank knows its own part's range: share them
            double v[2 * SSRS_MAX_RANKS] = {0.0};
            v[D.rank] = (double)init[D.rank]; v[P.n + D.rank] = (double)init[P.n + D.rank];
            AMG_RC(D.gather(v, 2 * P.n));
            for (int q = 0; q < 2 * P.n; ++q) init[q] = (i64)v[q];
        }
        for (int q = 0; q < P.n; ++q) { L.ref_lo[q] = init[q]; L.ref_hi[q] = init[P.n + q]; }
    }
    return SSRS_OK;
}

int coarse_solve(Hierarchy& H, Level& C, stream_t st) {
    if (H.coarse_sweeps > 0) {
        const Ell e = ell_of(C);
        AMG_TRY(ell_first32(e, 0, C.n, C.b32, C.x32, H.omega, st));
        for (int s = 1; s < H.coarse_sweeps; ++s) {
            AMG_TRY(ell_jacobi32(e, 0, C.n, C.b32, C.x32, C.t32, H.omega, st));
            real* sw = C.x32; C.x32 = C.t32; C.t32 = sw;
        }
        return SSRS_OK;
    }
    const double* inv = H.cinv; const real* b = C.b32; real* x = C.x32; const i64 n = C.n;
    return pfor_warp_rows(0, n, st,
                          [=] SSRS_HD(i64 i, int lane) { double s = 0.0; for (i64 j = lane; j < n; j += 32) s += inv[i * n + j] * (double)b[j]; return s; },
                          [=] SSRS_HD(i64 i, double total) { x[i] = (real)total; });
}

// out = M^-1 rhs: one V(nu, nu) cycle from a zero guess; rhs must be zero at Dirichlet nodes.  In the row-sharded
// solve rhs/out are valid on the owned rows (rhs ghost rows are refreshed here; out's are not), levels below
// H.lrep are distributed and every operator application is preceded by a ghost exchange of its input vector.
int vcycle(Hierarchy& H, const double* rhs, double* out, stream_t st) {
    const int nl = (int)H.lv.size(), nu = H.nu;
    const Fine32 F = H.f32;
    const real om = H.omega;
    const int cols = F.cols;
    i64 f0, f1;
    own_range(H, 0, f0, f1);
    const int r0 = (int)(f0 / cols), r1 = (int)(f1 / cols);
    AMG_RC(exchange_ghosts(H, 0, (void*)rhs, sizeof(double)));
    if (nl == 1) {          // no coarse level: plain Jacobi sweeps
        AMG_TRY(fine_first32(F, r0, r1, rhs, H.xf, om, st));
        for (int s = 1; s < 2 * nu - 1; ++s) {
            AMG_RC(exchange_ghosts(H, 0, H.xf, sizeof(real)));
            AMG_TRY(fine_jacobi32(F, r0, r1, rhs, H.xf, H.tf, om, st));
            real* sw = H.xf; H.xf = H.tf; H.tf = sw;
        }
        AMG_RC(exchange_ghosts(H, 0, H.xf, sizeof(real)));
        return fine_jacobi32(F, r0, r1, rhs, H.xf, out, om, st);
    }
    if (nu == 1) { AMG_TRY(fine_first_residual32(F, r0, r1, rhs, H.xf, H.resf, om, st)); }
    else {
        AMG_TRY(fine_first32(F, r0, r1, rhs, H.xf, om, st));
        for (int s = 1; s < nu; ++s) {
            AMG_RC(exchange_ghosts(H, 0, H.xf, sizeof(real)));
            AMG_TRY(fine_jacobi32(F, r0, r1, rhs, H.xf, H.tf, om, st));
            real* sw = H.xf; H.xf = H.tf; H.tf = sw;
        }
        AMG_RC(exchange_ghosts(H, 0, H.xf, sizeof(real)));
        AMG_TRY(fine_residual32(F, r0, r1, rhs, H.xf, H.resf, st));
    }
    // restriction to level l+1 covers the aggregates this rank owns there (all of their members are local);
    // entering the redundantly computed levels the pieces are gathered on every rank
    // also writes the next level's first sweep when that level goes on to smooth (not the coarsest, not gathered)
    bool first_done = false;
    auto restrict_to = [&](int l, const real* res) -> int {
        Level& C = H.lv[(size_t)l + 1];
        i64 c0 = 0, c1 = C.n;
        const bool gather = H.comm != nullptr && l < H.lrep && l + 1 >= H.lrep;
        if (H.comm != nullptr && l < H.lrep) { c0 = C.parts.lo[H.rank]; c1 = C.parts.lo[H.rank + 1]; }
        first_done = !gather && l + 1 < nl - 1 && !(nu == 1 && H.fuse_coarse_first);
        AMG_TRY(restrict32(H.lv[(size_t)l], c0, c1, res, C.b32, first_done ? C.x32 : nullptr, C.dinv, om, st));
        if (gather) {
            i64 offs[SSRS_MAX_RANKS + 1];
            for (int q = 0; q <= H.nparts; ++q) offs[q] = C.parts.lo[q] * (i64)sizeof(real);
            if (H.comm->allgather(H.comm->ctx, C.b32, offs, (void*)st) != 0) { set_error("ssrs_potential_solve: all-gather failed on level %d", l + 1); return SSRS_ERR_CUDA; }
        }
        return SSRS_OK;
    };
    AMG_RC(restrict_to(0, H.resf));
    // one level on the way down / up; `keep`: the smoothed iterate is copied back instead of swapping the level's two
    // buffers, so that a captured graph stays valid from cycle to cycle
    auto down = [&](int l) -> int {
        Level& L = H.lv[(size_t)l];
        const Ell e = ell_of(L);
        i64 i0, i1;
        own_range(H, l, i0, i1);
        AMG_RC(exchange_ghosts(H, l, L.b32, sizeof(real)));
        const int nul = l >= H.nu_from ? H.nu_coarse : nu;
        if (nul == 1 && H.fuse_coarse_first) { AMG_TRY(ell_first_residual32(e, i0, i1, L.b32, L.x32, L.r32, om, st)); }
        else {
            if (!first_done) AMG_TRY(ell_first32(e, i0, i1, L.b32, L.x32, om, st));
            for (int s = 1; s < nul; ++s) {
                AMG_RC(exchange_ghosts(H, l, L.x32, sizeof(real)));
                AMG_TRY(level_jacobi(L, i0, i1, L.b32, L.x32, L.t32, om, st));
                real* sw = L.x32; L.x32 = L.t32; L.t32 = sw;
            }
            // (the first sweep is elementwise, so the ghosts of x can be formed locally from the exchanged b)
            if (nul == 1 && H.comm != nullptr && l < H.lrep) {
                const i64 g0 = L.ref_lo[H.rank], g1 = L.ref_hi[H.rank];
                AMG_TRY(ell_first32(e, g0, i0, L.b32, L.x32, om, st));
                AMG_TRY(ell_first32(e, i1, g1, L.b32, L.x32, om, st));
            } else AMG_RC(exchange_ghosts(H, l, L.x32, sizeof(real)));
            AMG_TRY(level_residual(L, i0, i1, L.b32, L.x32, L.r32, st));
        }
        return restrict_to(l, L.r32);
    };
    auto up = [&](int l, bool keep) -> int {
        Level& L = H.lv[(size_t)l];
        i64 i0, i1;
        own_range(H, l, i0, i1);
        AMG_TRY(prolong_add32(L, i0, i1, L.x32, H.lv[(size_t)l + 1].x32, H.overcorrect, st));
        const int nul = l >= H.nu_from ? H.nu_coarse : nu;
        for (int s = 0; s < nul; ++s) {
            AMG_RC(exchange_ghosts(H, l, L.x32, sizeof(real)));
            AMG_TRY(level_jacobi(L, i0, i1, L.b32, L.x32, L.t32, om, st));
            if (keep) { real* xx = L.x32; const real* tt = L.t32; AMG_TRY(pfor_range(i0, i1, st, [=] SSRS_HD(i64 i) { xx[i] = tt[i]; })); }
            else { real* sw = L.x32; L.x32 = L.t32; L.t32 = sw; }
        }
        return SSRS_OK;
    };
    const int lg = H.lg < nl ? (H.lg > 1 ? H.lg : 1) : nl;       // levels >= lg: the graph's part
    for (int l = 1; l < nl - 1 && l < lg; ++l) AMG_RC(down(l));
    auto coarse_part = [&](bool keep) -> int {
        for (int l = lg; l < nl - 1; ++l) AMG_RC(down(l));
        AMG_RC(coarse_solve(H, H.lv[(size_t)nl - 1], st));
        for (int l = nl - 2; l >= lg; --l) AMG_RC(up(l, keep));
        return SSRS_OK;
    };
    bool replayed = false;
#ifndef SSRS_HOST_EMU
    if (lg < nl) {
        if (H.gexec == nullptr) {
            // captured once per solve (same buffers, same branches every cycle), then replayed
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                H.lg = 1 << 30;              // this stream cannot be captured: plain launches from now on
            } else {
                const int rcc = coarse_part(true);
                const cudaError_t ec = cudaStreamEndCapture(st, &graph);
                if (rcc != 0 || ec != cudaSuccess || graph == nullptr) { cudaGetLastError(); set_error("ssrs_potential_solve: graph capture of the coarse levels failed"); return SSRS_ERR_CUDA; }
                cudaGraphExec_t ge = nullptr;
                const cudaError_t ei = cudaGraphInstantiate(&ge, graph, 0);
                cudaGraphDestroy(graph);
                if (ei != cudaSuccess) { cudaGetLastError(); set_error("ssrs_potential_solve: graph instantiation failed"); return SSRS_ERR_CUDA; }
                H.gexec = ge;
            }
        }
        if (H.gexec != nullptr) {
            AMG_TRY(cudaGraphLaunch((cudaGraphExec_t)H.gexec, st) == cudaSuccess ? 0 : -1);
            replayed = true;
        }
    }
#endif
    if (!replayed) AMG_RC(coarse_part(false));
    for (int l = (lg < nl - 1 ? lg : nl - 1) - 1; l >= 1; --l) AMG_RC(up(l, false));
    AMG_TRY(prolong_add32(H.lv[0], f0, f1, H.xf, H.lv[1].x32, H.overcorrect, st));
    for (int s = 0; s < nu - 1; ++s) {
        AMG_RC(exchange_ghosts(H, 0, H.xf, sizeof(real)));
        AMG_TRY(fine_jacobi32(F, r0, r1, rhs, H.xf, H.tf, om, st));
        real* sw = H.xf; H.xf = H.tf; H.tf = sw;
    }
    AMG_RC(exchange_ghosts(H, 0, H.xf, sizeof(real)));
    return fine_jacobi32(F, r0, r1, rhs, H.xf, out, om, st);
}

#ifndef SSRS_HOST_EMU
// Gauss-Jordan on the coarsest level in ONE launch: a single CTA walks the pivots, the n x n matrices stay in L2
// (n <= 512: 2 MB each).  Per element the same operations in the same order as the pivot-by-pivot loop below, which
// it replaces on the device (that loop costs two launches per pivot: 318 launches for the 159 coarsest rows of a
// 5000 x 6000 grid).
__global__ void __launch_bounds__(1024) dense_inverse_kernel(double* __restrict__ D, double* __restrict__ I, double* __restrict__ colk,
                                                             double* __restrict__ rowD, double* __restrict__ rowI, int n) {
    __shared__ double s_p;
    for (int k = 0; k < n; ++k) {
        if (threadIdx.x == 0) s_p = D[(size_t)k * n + k];
        __syncthreads();
        const double p = s_p;
        // the row factors colk[i] / p once per pivot instead of once per element (same quotient, n instead of n^2 divisions)
        for (int i = threadIdx.x; i < n; i += blockDim.x) { colk[i] = D[(size_t)i * n + k] / p; rowD[i] = D[(size_t)k * n + i]; rowI[i] = I[(size_t)k * n + i]; }
        __syncthreads();
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e / n, j = e - i * n;
            if (i == k) { D[e] = rowD[j] / p; I[e] = rowI[j] / p; }
            else { const double f = colk[i]; if (f != 0.0) { D[e] -= f * rowD[j]; I[e] -= f * rowI[j]; } }
        }
        __syncthreads();
    }
}
#endif

int dense_inverse(Hierarchy& H, const Level& C, Pool& pool, stream_t st) {
    const i64 n = C.n;
    double *D, *I, *colk, *rowD, *rowI;
    Pool tmp(st, true);
    D = tmp.get<double>(n * n); colk = tmp.get<double>(n); rowD = tmp.get<double>(n); rowI = tmp.get<double>(n);
    AMG_ALLOC(I, double, n * n);
    if (!D || !colk || !rowD || !rowI) { set_error("ssrs_potential_solve: out of device memory in dense_inverse"); return SSRS_ERR_CUDA; }
    AMG_TRY(dev_zero(D, sizeof(double) * (size_t)(n * n), st));
    AMG_TRY(dev_zero(I, sizeof(double) * (size_t)(n * n), st));
    const CsrGraph g = csr_of(C);
    AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
        I[i * n + i] = 1.0;
        for (i64 k = g.rowptr[i]; k < g.rowptr[i + 1]; ++k) D[i * n + g.col[k]] = g.val[k];
    }));
#ifndef SSRS_HOST_EMU
    if (n <= 512) {
        dense_inverse_kernel<<<1, 1024, 0, st>>>(D, I, colk, rowD, rowI, (int)n);
        AMG_TRY(cudaGetLastError() == cudaSuccess ? 0 : -1);
        AMG_TRY(sync(st));
        H.cinv = I; H.cn = n;
        return SSRS_OK;
    }
#endif
    for (i64 k = 0; k < n; ++k) {       // Gauss-Jordan; diagonally dominant M-matrix: no pivoting needed
        AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) { colk[i] = D[i * n + k]; rowD[i] = D[k * n + i]; rowI[i] = I[k * n + i]; }));
        AMG_TRY(pfor(n * n, st, [=] SSRS_HD(i64 e) {
            const i64 i = e / n, j = e - i * n;
            const double p = colk[k];
            if (i == k) { D[e] = rowD[j] / p; I[e] = rowI[j] / p; }
            else { const double f = colk[i] / p; if (f != 0.0) { D[e] -= f * rowD[j]; I[e] -= f * rowI[j]; } }
        }));
    }
    AMG_TRY(sync(st));
    H.cinv = I; H.cn = n;
    return SSRS_OK;
}

// Distributed setup, first redundant level: its CSR rows were built by their owners (local row pointers); every rank
// now needs all of them.  Row lengths are all-gathered and scanned, the owners' entries land at their global offsets
// and are all-gathered in place; the level's ELL form is then rebuilt over all rows.
int gather_level(Hierarchy& H, Level& L, Pool& pool, stream_t st) {
    const ssrs_comm* comm = H.comm;
    const Parts P = L.parts;
    const int R = comm->rank;
    const i64 n = L.n, lo = P.lo[R], hi = P.lo[R + 1];
    i64* rp;
    AMG_ALLOC(rp, i64, n + 1);
    { const i64* lrp = L.rowptr; AMG_TRY(pfor_range(lo, hi, st, [=] SSRS_HD(i64 i) { rp[i] = lrp[i + 1] - lrp[i]; })); }
    i64 offs[SSRS_MAX_RANKS + 1];
    for (int q = 0; q <= P.n; ++q) offs[q] = P.lo[q] * (i64)sizeof(i64);
    if (comm->allgather(comm->ctx, rp, offs, (void*)st) != 0) { set_error("ssrs_potential_solve: all-gather of row lengths failed"); return SSRS_ERR_CUDA; }
    AMG_TRY(pfor_range(n, n + 1, st, [=] SSRS_HD(i64 i) { rp[i] = 0; }));
    i64 nnz = 0;
    AMG_TRY(exclusive_scan_i64(rp, n + 1, &nnz, st));
    i64 starts[SSRS_MAX_RANKS + 1];
    for (int q = 0; q <= P.n; ++q) {
        if (P.lo[q] >= n) starts[q] = nnz;
        else AMG_TRY(copy_d2h(&starts[q], rp + P.lo[q], sizeof(i64), st));
    }
    AMG_TRY(sync(st));
    int* col; double* val;
    AMG_ALLOC(col, int, nnz);
    AMG_ALLOC(val, double, nnz);
    if (starts[R + 1] - starts[R] != L.nnz) { set_error("ssrs_potential_solve: gathered row lengths do not match the local rows"); return SSRS_ERR_CUDA; }
    AMG_TRY(copy_d2d(col + starts[R], L.col, sizeof(int) * (size_t)L.nnz, st));
    AMG_TRY(copy_d2d(val + starts[R], L.val, sizeof(double) * (size_t)L.nnz, st));
    for (int q = 0; q <= P.n; ++q) offs[q] = starts[q] * (i64)sizeof(int);
    if (comm->allgather(comm->ctx, col, offs, (void*)st) != 0) { set_error("ssrs_potential_solve: all-gather of the coarse columns failed"); return SSRS_ERR_CUDA; }
    for (int q = 0; q <= P.n; ++q) offs[q] = starts[q] * (i64)sizeof(double);
    if (comm->allgather(comm->ctx, val, offs, (void*)st) != 0) { set_error("ssrs_potential_solve: all-gather of the coarse values failed"); return SSRS_ERR_CUDA; }
    L.rowptr = rp; L.col = col; L.val = val; L.nnz = nnz;
    return build_ell(L, pool, st);            // over all rows; the owner-only arrays of the first build stay in the pool
}

// out = b - A x at free nodes (b = 0 there: the Dirichlet values live in x), 0 at Dirichlet nodes; *nrm2 = |out|^2
int fine_residual(const FineGraph fg, const FineWeights W, int r0, int r1, const double* x, double* out, double* nrm2, stream_t st) {
    double dummy;
    AMG_TRY(preduce2d_sum2_rows(r0, r1, fg.cols, st, nrm2, &dummy, [=] SSRS_HD(int r, int c, double& u0, double& u1) {
        const i64 i = (i64)r * fg.cols + c;
        const bool free_node = !fg.excluded(i);
        const double ax = fine_apply64(fg, W, r, c, x);
        const double v = free_node ? -ax : 0.0;
        out[i] = v;
        u0 = v * v; u1 = 0.0;
    }));
    return SSRS_OK;
}
// out = A in for a correction vector `in` (zero at Dirichlet nodes); *d0 = <w0, out>, *d1 = <out, out>
int fine_apply_dots(const FineGraph fg, const FineWeights W, int r0, int r1, const double* in, double* out, const double* w0,
                    double* d0, double* d1, stream_t st) {
    AMG_TRY(preduce2d_sum2_rows(r0, r1, fg.cols, st, d0, d1, [=] SSRS_HD(int r, int c, double& u0, double& u1) {
        const i64 i = (i64)r * fg.cols + c;
        const bool free_node = !fg.excluded(i);
        const double wi = w0[i];                       // (issued with the stencil's loads, not after them)
        const double ax = fine_apply64(fg, W, r, c, in);
        const double v = free_node ? ax : 0.0;
        out[i] = v;
        u0 = wi * v; u1 = v * v;
    }));
    return SSRS_OK;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace amg
}  // namespace ssrs

using namespace ssrs;
using namespace ssrs::amg;

#ifdef SSRS_HOST_EMU
#define SOLVE_NAME ssrs_emu_potential_solve
#define SOLVE_SHARDED_NAME ssrs_emu_potential_solve_sharded
extern "C" __attribute__((visibility("default"))) const char* ssrs_emu_last_error(void) { return ssrs::g_emu_err; }
#else
#define SOLVE_NAME ssrs_potential_solve
#define SOLVE_SHARDED_NAME ssrs_potential_solve_sharded
#endif

namespace ssrs { namespace amg {
int solve_impl(const float* K, int rows, int cols, const int64_t* bnodes_host, const double* bvalues_host,
               int64_t n_bnodes, double rtol, int max_iter, float* phi, ssrs_solve_stats* stats, const ssrs_comm* comm,
               void* stream, double* phi64 = nullptr) {
    if (comm != nullptr && comm->size <= 1) comm = nullptr;
    if (comm != nullptr) {
        if (comm->size > SSRS_MAX_RANKS || comm->rank < 0 || comm->rank >= comm->size || !comm->exchange || !comm->allreduce_sum || !comm->allgather) {
            set_error("ssrs_potential_solve_sharded: bad communicator (size %d, rank %d)", comm->size, comm->rank);
            return SSRS_ERR_INVALID;
        }
        if (rows < 4 * comm->size) { set_error("ssrs_potential_solve_sharded: %d rows cannot be split over %d ranks", rows, comm->size); return SSRS_ERR_INVALID; }
    }
    if (K == nullptr || phi == nullptr) { set_error("ssrs_potential_solve: NULL raster"); return SSRS_ERR_INVALID; }
    if (rows < 3 || cols < 3) { set_error("ssrs_potential_solve: grid %dx%d too small", rows, cols); return SSRS_ERR_INVALID; }
    if ((int64_t)rows * cols > 2000000000LL) { set_error("ssrs_potential_solve: more than 2e9 cells"); return SSRS_ERR_UNSUPPORTED; }
    if (n_bnodes <= 0 || bnodes_host == nullptr || bvalues_host == nullptr) { set_error("ssrs_potential_solve: no Dirichlet nodes"); return SSRS_ERR_INVALID; }
    if (!(rtol > 0.0)) rtol = 0.0;          // default: iterate to the attainable accuracy (see below)
    if (max_iter <= 0) max_iter = 300;
    stream_t st = (stream_t)stream;
#ifndef SSRS_HOST_EMU
    // The coarse levels of the cycle are replayed as a CUDA graph, which the legacy default stream cannot capture: a
    // solve called on it runs on an internal stream ordered after it (the solve is synchronous, so ordering at the
    // end is the final synchronisation).
    if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) {
        static thread_local cudaStream_t own[64] = {nullptr};
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("ssrs_potential_solve: bad device"); return SSRS_ERR_CUDA; }
        if (own[dev] == nullptr && cudaStreamCreateWithFlags(&own[dev], cudaStreamNonBlocking) != cudaSuccess) { set_error("ssrs_potential_solve: cannot create a stream"); return SSRS_ERR_CUDA; }
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { set_error("ssrs_potential_solve: cannot create an event"); return SSRS_ERR_CUDA; }
        cudaEventRecord(ev, st);
        cudaStreamWaitEvent(own[dev], ev, 0);
        cudaEventDestroy(ev);
        st = own[dev];
    }
#endif
    const i64 n = (i64)rows * cols;
    const double t_begin = now_ms();
    const bool trace = getenv("SSRS_SOLVE_TRACE") != nullptr;      // per-iteration residuals on stderr
#ifndef SSRS_HOST_EMU
    // measured: 304 B/cell live until the solve returns, ~150 B/cell of scoped temporaries (Galerkin scratch)
    arena(false).reserve((size_t)n * 330 + ((size_t)64 << 20));
    arena(true).reserve((size_t)n * 170 + ((size_t)64 << 20));
#endif
    Pool pool(st);
    Hierarchy H;
    g_ell_packed = !(getenv("SSRS_X_ELLPACK") && atoi(getenv("SSRS_X_ELLPACK")) == 0);
    if (getenv("SSRS_X_OC")) H.overcorrect = (float)atof(getenv("SSRS_X_OC"));
    if (getenv("SSRS_X_OMEGA")) H.omega = (float)atof(getenv("SSRS_X_OMEGA"));
    if (getenv("SSRS_X_NU")) H.nu = atoi(getenv("SSRS_X_NU"));
    if (getenv("SSRS_X_NUC")) H.nu_coarse = atoi(getenv("SSRS_X_NUC"));
    if (getenv("SSRS_X_NUL")) H.nu_from = atoi(getenv("SSRS_X_NUL"));
    if (getenv("SSRS_X_FUSEC")) H.fuse_coarse_first = atoi(getenv("SSRS_X_FUSEC")) != 0;

    // Dirichlet nodes arrive as the reference's column-major ids (movmodel.py:25-29): i = col*nrow + row
    std::vector<int> bidx((size_t)n_bnodes);
    std::vector<double> bval((size_t)n_bnodes);
    double bsum = 0.0;
    for (int64_t q = 0; q < n_bnodes; ++q) {
        const int64_t id = bnodes_host[q];
        if (id < 0 || id >= n) { set_error("ssrs_potential_solve: boundary node id %lld outside the grid", (long long)id); return SSRS_ERR_INVALID; }
        bidx[(size_t)q] = (int)((id % rows) * cols + id / rows);
        bval[(size_t)q] = bvalues_host[q];
        bsum += bvalues_host[q];
    }
    const double guess = bsum / (double)n_bnodes;
    int* d_bidx; double* d_bval; float* kd;
    AMG_ALLOC(d_bidx, int, n_bnodes);
    AMG_ALLOC(d_bval, double, n_bnodes);
    AMG_ALLOC(kd, float, n);
    AMG_TRY(copy_h2d(d_bidx, bidx.data(), sizeof(int) * (size_t)n_bnodes, st));
    AMG_TRY(copy_h2d(d_bval, bval.data(), sizeof(double) * (size_t)n_bnodes, st));
    double *x, *r, *rh, *p, *v, *s, *t, *y, *z;
    AMG_ALLOC(x, double, n); AMG_ALLOC(r, double, n); AMG_ALLOC(rh, double, n); AMG_ALLOC(p, double, n);
    AMG_ALLOC(v, double, n); AMG_ALLOC(s, double, n); AMG_ALLOC(t, double, n);
    AMG_ALLOC(y, double, n); AMG_ALLOC(z, double, n);
    AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) { float k = K[i]; kd[i] = k < 0.0f ? 0.0f : fabsf(k); x[i] = guess; }));
    AMG_TRY(pfor(n_bnodes, st, [=] SSRS_HD(i64 q) {
        const int i = d_bidx[q];
        unsigned u; float k = kd[i];
        memcpy(&u, &k, 4); u |= 0x80000000u; memcpy(&k, &u, 4);
        kd[i] = k;
        x[i] = d_bval[q];
    }));
    AMG_TRY(sync(st));
    int interior = 0;
    for (int64_t q = 0; q < n_bnodes; ++q) {
        const int rr = bidx[(size_t)q] / cols, cc = bidx[(size_t)q] % cols;
        if (rr > 0 && rr < rows - 1 && cc > 0 && cc < cols - 1) { interior = 1; break; }
    }
    H.fine.kd = kd; H.fine.rows = rows; H.fine.cols = cols; H.fine.interior_dirichlet = interior;
    H.comm = comm; H.st = st;
    H.lv.emplace_back();
    H.lv[0].n = n;
    const int fake_parts = (comm == nullptr && getenv("SSRS_X_FAKEPARTS")) ? atoi(getenv("SSRS_X_FAKEPARTS")) : 0;   // experiment: slab-constrained coarsening on one rank
    if (comm != nullptr || fake_parts > 1) {          // contiguous row slabs; the fine level's ghost zone is one row on each side
        if (comm != nullptr) { H.rank = comm->rank; H.nparts = comm->size; }
        Parts P;
        P.n = comm != nullptr ? comm->size : fake_parts;
        for (int q = 0; q <= P.n; ++q) P.lo[q] = (i64)((int64_t)rows * q / P.n) * cols;
        H.fine.parts = P;
        H.lv[0].parts = P;
        for (int q = 0; q < P.n; ++q) {
            H.lv[0].ref_lo[q] = q > 0 ? P.lo[q] - cols : 0;
            H.lv[0].ref_hi[q] = q + 1 < P.n ? P.lo[q + 1] + cols : n;
        }
    }
    // Distributed setup (default for the row-sharded solve): every rank builds its own rows of the distributed levels;
    // SSRS_X_REDUNDANT_SETUP=1 keeps round 1's redundant setup (every rank builds everything) for A/B comparison.
    const bool dist_setup = comm != nullptr && !(getenv("SSRS_X_REDUNDANT_SETUP") && atoi(getenv("SSRS_X_REDUNDANT_SETUP")) != 0);
    {
        float *wf, *dinv; double* wd; unsigned *wen, *wdd;
        Pool wtmp(st, true);                                   // float32 weights: only needed to form the diagonal
        wf = wtmp.get<float>(4 * n);
        if (!wf) { set_error("ssrs_potential_solve: out of device memory (fine weights)"); return SSRS_ERR_CUDA; }
        AMG_ALLOC(wen, unsigned, n);
        AMG_ALLOC(wdd, unsigned, n);
        AMG_ALLOC(wd, double, 4 * n);
        AMG_ALLOC(dinv, float, n);
        const FineGraph fgw = H.fine;
        // row-sharded: this rank reads the links of its own rows and of the halo row on each side, and — the cycle's
        // fused first sweep forms x = omega dinv b on the halo rows locally — the diagonals of the halo rows, which
        // need the links of one more row
        int wr0 = 0, wr1 = rows, dr0 = 0, dr1 = rows;
        if (dist_setup) {
            const int R0s = (int)(H.lv[0].parts.lo[H.rank] / cols), R1s = (int)(H.lv[0].parts.lo[H.rank + 1] / cols);
            wr0 = R0s > 2 ? R0s - 2 : 0; wr1 = R1s + 2 < rows ? R1s + 2 : rows;
            dr0 = R0s > 1 ? R0s - 1 : 0; dr1 = R1s + 1 < rows ? R1s + 1 : rows;
        }
        AMG_TRY(pfor2d_rows(wr0, wr1, cols, st, [=] SSRS_HD(int r, int c) {
            const i64 i = (i64)r * cols + c;
            const float kc = fgw.kd[i];
            const bool hasE = c < cols - 1, hasN = r < rows - 1, hasW = c > 0;
            const double e = hasE ? link_weight<false>(kc, fgw.kd[i + 1], false) : 0.0;
            const double nn = hasN ? link_weight<false>(kc, fgw.kd[i + cols], false) : 0.0;
            const double ne = (hasN && hasE) ? link_weight<false>(kc, fgw.kd[i + cols + 1], true) : 0.0;
            const double nw = (hasN && hasW) ? link_weight<false>(kc, fgw.kd[i + cols - 1], true) : 0.0;
            wd[i] = e; wd[n + i] = nn; wd[2 * n + i] = ne; wd[3 * n + i] = nw;
            wf[i] = (float)e; wf[n + i] = (float)nn; wf[2 * n + i] = (float)ne; wf[3 * n + i] = (float)nw;
            wen[i] = bf16_bits((float)e) | (bf16_bits((float)nn) << 16);
            wdd[i] = bf16_bits((float)ne) | (bf16_bits((float)nw) << 16);
        }));
        H.fw.wf = nullptr; H.fw.wd = wd;
        Fine32 F;
        F.w.en.p = wen; F.w.dd.p = wdd; F.dinv = dinv; F.kd = kd; F.rows = rows; F.cols = cols;
        const float *fE = wf, *fN = wf + n, *fNE = wf + 2 * n, *fNW = wf + 3 * n;
        AMG_TRY(pfor2d_rows(dr0, dr1, cols, st, [=] SSRS_HD(int r, int c) {       // Jacobi diagonal of the float32 operator
            const i64 i = (i64)r * cols + c;
            if (fgw.excluded(i)) { dinv[i] = 0.0f; return; }
            const bool hW = c > 0, hE = c < cols - 1, hS = r > 0, hN = r < rows - 1;
            real wS = hS ? fN[i - cols] : (real)0.0, wSW = (hS && hW) ? fNE[i - cols - 1] : (real)0.0;
            if (c == cols - 1 && hS && hN) {
                wS = (real)link_weight<false>(F.kd[i], F.kd[i - cols], true);
                wSW = (real)link_weight<false>(F.kd[i], F.kd[i - cols - 1], false);
            }
            const real d = ((fE[i] + (hW ? fE[i - 1] : (real)0.0)) + (fN[i] + wS)) +
                            ((fNE[i] + wSW) + (fNW[i] + ((hS && hE) ? fNW[i - cols + 1] : (real)0.0)));
            dinv[i] = (float)((real)1.0 / d);
        }));
        H.f32 = F;
        AMG_ALLOC(H.xf, real, n);
        AMG_ALLOC(H.tf, real, n);
        AMG_ALLOC(H.resf, real, n);
    }

    // ---- setup: hierarchy ----
    const double theta = 0.5;
    const i64 coarse_target = 400, dense_cap = 2048;
    i64 total_nnz = 9 * n, ell_total = 0;
    // Row-sharded solve: a level stays distributed while it is large and every part's rows reference the
    // adjacent parts only; its aggregates must then not straddle a slab boundary.  From the first level that is
    // not (H.lrep) all ranks compute redundantly and coarsening is unconstrained again — constraining every
    // level leaves a seam that the coarse levels never close and costs ~40 % more iterations.  Measured on 4 GPUs
    // (5000 x 6000 / 10000 x 12000): threshold 65 536 rows -> 56 / 47 iterations, 166 / 382 ms; 1.2e6 -> 43 / 44,
    // 129 / 362 ms; 5e6 -> 42 / 44, 155 / 382 ms (single GPU: 41 / 43 iterations, 283 / 1128 ms).
    const i64 rep_rows = getenv("SSRS_X_REPROWS") ? atoll(getenv("SSRS_X_REPROWS")) : 1000000;
    bool distributed = H.lv[0].parts.n > 1;
    for (int l = 0; l < 40; ++l) {
        Level& L = H.lv[(size_t)l];
        if (l > 0 && L.n <= coarse_target) break;
        if (distributed && l > 0) {
            bool ok = L.n >= rep_rows;
            for (int q = 0; q < L.parts.n && ok; ++q) {
                if (q > 0 && L.ref_lo[q] < L.parts.lo[q - 1]) ok = false;
                if (q + 1 < L.parts.n && L.ref_hi[q] > L.parts.lo[q + 2]) ok = false;
            }
            if (!ok) {
                distributed = false; H.lrep = l;
                if (dist_setup) { const int rcg = gather_level(H, L, pool, st); if (rcg) return rcg; }    // every rank continues with the whole level
            }
        }
        SetupDist D;                                   // this level's rows are built by their owners only
        if (dist_setup && distributed) { D.comm = comm; D.rank = comm->rank; D.size = comm->size; D.st = st; }
        Parts next;
        int rc;
        if (l == 0) { FineGraph g = H.fine; if (!distributed) g.parts = Parts(); rc = coarsen(g, L, pool, theta, 8, 3, &next, st, D); }
        else { CsrGraph g = csr_of(L); if (!distributed) g.parts = Parts(); rc = coarsen(g, L, pool, theta, 8, 3, &next, st, D); }
        if (rc) return rc;
        if (L.nc < 1 || (double)L.nc > 0.9 * (double)L.n) {        // stalled: stop here
            pool.release(L.agg); pool.release(L.memptr); pool.release(L.mem);
            L.agg = nullptr; L.memptr = nullptr; L.mem = nullptr; L.nc = 0;
            break;
        }
        Level C;
        if (D.on()) {
            // Galerkin rows need the aggregate ids of the columns in the neighbours' slabs
            AMG_RC(exchange_ghosts(H, l, L.agg, sizeof(int)));
            const i64 clo = next.lo[D.rank], chi = next.lo[D.rank + 1];
            rc = (l == 0) ? galerkin(H.fine, L, C, pool, st, clo, chi) : galerkin(csr_of(L), L, C, pool, st, clo, chi);
        } else rc = (l == 0) ? galerkin(H.fine, L, C, pool, st) : galerkin(csr_of(L), L, C, pool, st);
        if (rc) return rc;
        { double nz = (double)C.nnz; AMG_RC(D.gather(&nz, 1)); total_nnz += (i64)nz; }
        C.parts = next;
        rc = build_ell(C, pool, st, D);
        if (rc) return rc;
        ell_total += C.ell_entries;
        H.lv.push_back(C);
        // the cycle forms the ghost entries of a level's first sweep locally (x = omega dinv b): the ghost rows' inverse
        // diagonals come from their owners
        if (D.on()) AMG_RC(exchange_ghosts(H, (int)H.lv.size() - 1, H.lv.back().dinv, sizeof(float)));
    }
    if (dist_setup && distributed && H.lv.size() > 1) {
        // the coarsest level is always redundant: if the hierarchy ended while still distributed, gather it now
        distributed = false; H.lrep = (int)H.lv.size() - 1;
        const int rcg = gather_level(H, H.lv.back(), pool, st);
        if (rcg) return rcg;
    }
    {
        Level& C = H.lv.back();
        if (H.lv.size() > 1) {
            if (C.n <= dense_cap) { int rc = dense_inverse(H, C, pool, st); if (rc) return rc; }
            else H.coarse_sweeps = 60;
        }
    }
    if (H.lv[0].parts.n > 1) {
        const int nl = (int)H.lv.size();
        if (H.lrep > nl - 1) H.lrep = nl - 1;       // the coarsest level is always redundant
        if (H.lrep < 1) H.lrep = 1;
        if (trace) fprintf(stderr, "ssrs_potential_solve: rank %d of %d, levels %d, redundant from level %d\n", H.rank, H.lv[0].parts.n, nl, H.lrep);
    }
#ifndef SSRS_HOST_EMU
    if (!(getenv("SSRS_X_NOGRAPH") && atoi(getenv("SSRS_X_NOGRAPH")) != 0)) {
        // the graph's part of the cycle: from the first level that is small and computed in full by this rank
        const int nl = (int)H.lv.size();
        for (int l = 1; l < nl; ++l)
            if ((comm == nullptr || l >= H.lrep) && H.lv[(size_t)l].n <= (i64)1 << 20) { H.lg = l; break; }
    }
    struct GraphGuard { Hierarchy& h; ~GraphGuard() { if (h.gexec) { cudaGraphExecDestroy((cudaGraphExec_t)h.gexec); h.gexec = nullptr; } } } graph_guard{H};
#endif
    const double t_setup = now_ms();

    // ---- BiCGStab, right preconditioned ----
    // Row-sharded: every vector operation covers the owned rows [R0, R1) (cells [I0, I1)); inner products are
    // summed over the ranks (bit-identical on every rank, so all ranks take the same branches); an operator's
    // input gets its ghost rows refreshed first.
    const FineGraph fg = H.fine;
    i64 I0, I1;
    own_range(H, 0, I0, I1);
    const int R0 = (int)(I0 / cols), R1 = (int)(I1 / cols);
    auto allsum = [&](double* a, double* b) -> int {
        if (comm == nullptr) return 0;
        double h[2] = {*a, b ? *b : 0.0};
        if (comm->allreduce_sum(comm->ctx, h, 2, (void*)st) != 0) { set_error("ssrs_potential_solve: all-reduce failed"); return SSRS_ERR_CUDA; }
        *a = h[0]; if (b) *b = h[1];
        return 0;
    };
    auto true_residual = [&](double* out, double* nrm2) -> int {
        AMG_RC(exchange_ghosts(H, 0, x, sizeof(double)));
        AMG_RC(fine_residual(fg, H.fw, R0, R1, x, out, nrm2, st));
        return allsum(nrm2, nullptr);
    };
    double r0n2 = 0.0;
    AMG_RC(true_residual(r, &r0n2));
    const double r0 = sqrt(r0n2);
    // Attainable accuracy: the nearest float64-representable potential leaves a residual of about
    // d_i * ulp(phi_i) / 2 per cell (d_i = row diagonal); below that the recurrence residual keeps falling but
    // the true one does not (the estimate is 1.5-4x above the floor measured on 300 k .. 30 M cell grids).  The
    // recurrence residual is iterated to the larger of rtol and a fraction of that floor, then the true residual decides:
    // accepted below the same fraction (or when it has stagnated), else the iteration restarts from the current iterate.
    double floor2 = 0.0, bmax = 0.0;
    for (int64_t q = 0; q < n_bnodes; ++q) bmax = fabs(bvalues_host[q]) > bmax ? fabs(bvalues_host[q]) : bmax;
    { const float* dinv = H.f32.dinv; const double scale = 0.5 * 2.220446049250313e-16 * bmax;
      AMG_TRY(preduce_sum_range(I0, I1, st, &floor2, [=] SSRS_HD(i64 i) { const double di = (double)dinv[i]; const double e = di > 0.0 ? scale / di : 0.0; return e * e; }));
      AMG_RC(allsum(&floor2, nullptr)); }
    const double floor_rel = (r0 > 0.0) ? sqrt(floor2) / r0 : 0.0;
    if (trace) fprintf(stderr, "ssrs_potential_solve: r0 %.3e attainable relative residual ~ %.3e\n", r0, floor_rel);
    // The fraction is set by accuracy, measured against the refined truth of the 1000 x 1200 / 10 m system
    // (tests/golden/potential_truth10m.npz; tools/solver_truth_sweep.py, profiles/r02_solver_truth_sweep.txt): floor/2 leaves
    // the float32 potential up to 5 ulp from the truth (27 iterations; 36 and 256 ms at 5000 x 6000), floor/8 2 ulp
    // (30; 38, 268 ms), floor/16 1 ulp with < 1 % of the cells differing at all (32; 40, 280 ms), floor/32 1 ulp and
    // 0.01 % (37; 44, 303 ms).  The reference's own unrefined SuperLU answer is 14 ulp off.  floor/16: tracks are steered
    // by the potential's last float32 bits (SURVEY §0 findings 4 and 6), so the extra 9 % buys the 1-ulp answer.
    // The error behind a given residual grows with the conditioning, i.e. with the cell count: against a float64
    // self-truth (the same solver driven to rtol 1e-12; tools/solver_truth_large.py, profiles/r02_solver_truth_large.txt)
    // floor/16 is 0.5 ulp at 5000 x 6000 but 6.5 ulp at 10000 x 12000, where floor/64 is 0.54 ulp (49 iterations instead
    // of 47, +3 %).  So the fraction shrinks in proportion to the cell count above 3e7 cells.
    const double size_scale = (double)n > 3.0e7 ? 3.0e7 / (double)n : 1.0;
    const double floor_frac = getenv("SSRS_X_FLOORFRAC") ? atof(getenv("SSRS_X_FLOORFRAC")) : 0.0625 * size_scale;
    const double accept_frac = getenv("SSRS_X_ACCEPT") ? atof(getenv("SSRS_X_ACCEPT")) : 0.0625 * size_scale;
    const double tol_eff = rtol > floor_frac * floor_rel ? rtol : floor_frac * floor_rel;
    int iters = 0, restarts = 0, converged = (r0 == 0.0);
    double best_true = 1.0;
    double rel = (r0 == 0.0) ? 0.0 : 1.0;
    const size_t own_bytes = sizeof(double) * (size_t)(I1 - I0);
    while (!converged && iters < max_iter && restarts <= 6) {
        AMG_TRY(copy_d2d(rh + I0, r + I0, own_bytes, st));
        AMG_TRY(dev_zero(p + I0, own_bytes, st));
        AMG_TRY(dev_zero(v + I0, own_bytes, st));
        double rho = 1.0, alpha = 1.0, om = 1.0, rho_new = 0.0;
        bool breakdown = false;
        { const double *a_ = rh, *b_ = r; AMG_TRY(preduce_sum_range(I0, I1, st, &rho_new, [=] SSRS_HD(i64 i) { return a_[i] * b_[i]; })); }
        AMG_RC(allsum(&rho_new, nullptr));
        while (iters < max_iter) {
            if (rho_new == 0.0 || !(fabs(rho_new) < 1e300)) { breakdown = true; break; }
            const double beta = (rho_new / rho) * (alpha / om);
            rho = rho_new;
            { double *pp = p; const double *rr = r, *vv = v; const double om_ = om;
              AMG_TRY(pfor_range(I0, I1, st, [=] SSRS_HD(i64 i) { pp[i] = rr[i] + beta * (pp[i] - om_ * vv[i]); })); }
            AMG_RC(vcycle(H, p, y, st));
            double rhv = 0.0, vv2 = 0.0;
            AMG_RC(exchange_ghosts(H, 0, y, sizeof(double)));
            AMG_RC(fine_apply_dots(fg, H.fw, R0, R1, y, v, rh, &rhv, &vv2, st));
            AMG_RC(allsum(&rhv, nullptr));
            if (rhv == 0.0) { breakdown = true; break; }
            alpha = rho / rhv;
            { double* ss = s; const double *rr = r, *vv = v; const double al = alpha;
              AMG_TRY(pfor_range(I0, I1, st, [=] SSRS_HD(i64 i) { ss[i] = rr[i] - al * vv[i]; })); }
            AMG_RC(vcycle(H, s, z, st));
            double ts = 0.0, tt = 0.0;
            AMG_RC(exchange_ghosts(H, 0, z, sizeof(double)));
            AMG_RC(fine_apply_dots(fg, H.fw, R0, R1, z, t, s, &ts, &tt, st));
            AMG_RC(allsum(&ts, &tt));
            om = (tt > 0.0) ? ts / tt : 0.0;
            double rn2 = 0.0;
            { double *xx = x, *rr = r; const double *yy = y, *zz = z, *ss = s, *tv = t, *rh_ = rh; const double al = alpha, om_ = om;
              AMG_TRY(preduce_sum2_range(I0, I1, st, &rn2, &rho_new, [=] SSRS_HD(i64 i, double& u0, double& u1) {
                  xx[i] += al * yy[i] + om_ * zz[i];
                  const double rv = ss[i] - om_ * tv[i];
                  rr[i] = rv;
                  u0 = rv * rv; u1 = rh_[i] * rv;
              })); }
            AMG_RC(allsum(&rn2, &rho_new));
            ++iters;
            rel = sqrt(rn2) / r0;
            if (trace) {
                double tr2 = 0.0;
                AMG_RC(true_residual(t, &tr2));
                fprintf(stderr, "ssrs_potential_solve: iter %d rel %.3e true %.3e alpha %.3e omega %.3e\n", iters, rel, sqrt(tr2) / r0, alpha, om);
            }
            if (!(rel == rel)) { breakdown = true; break; }
            if (rel <= tol_eff || om == 0.0) break;
        }
        // true residual: accept, or restart from the current iterate
        double tn2 = 0.0;
        AMG_RC(true_residual(r, &tn2));
        rel = sqrt(tn2) / r0;
        if (trace) fprintf(stderr, "ssrs_potential_solve: true residual %.3e after %d iterations\n", rel, iters);
        if (!(rel == rel)) { set_error("ssrs_potential_solve: NaN residual (NaN in the conductivity raster?)"); return SSRS_ERR_NOT_CONVERGED; }
        if (rtol > 0.0 && rel <= 4.0 * rtol) converged = 1;
        else if (rel <= accept_frac * floor_rel) converged = 2;                       // float64 attainable accuracy reached
        else if (!breakdown && rel > 0.5 * best_true && rel <= 1e-6) converged = 2;   // stagnated there
        else ++restarts;
        if (rel < best_true) best_true = rel;
    }
    { const double* xx = x; AMG_TRY(pfor_range(I0, I1, st, [=] SSRS_HD(i64 i) { phi[i] = (float)xx[i]; })); }     // movmodel.py:128
    if (phi64 != nullptr) {         // diagnostics: the float64 iterate before rounding (single-rank solves)
        const double* xx = x; AMG_TRY(pfor_range(I0, I1, st, [=] SSRS_HD(i64 i) { phi64[i] = xx[i]; }));
    }
    if (comm != nullptr) {          // every rank returns the full raster (the stepping stage replicates the fields)
        i64 offs[SSRS_MAX_RANKS + 1];
        for (int q = 0; q <= H.nparts; ++q) offs[q] = H.lv[0].parts.lo[q] * (i64)sizeof(float);
        if (comm->allgather(comm->ctx, phi, offs, (void*)st) != 0) { set_error("ssrs_potential_solve: all-gather of the potential failed"); return SSRS_ERR_CUDA; }
    }
    AMG_TRY(sync(st));
    const double t_end = now_ms();
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->iterations = iters; stats->restarts = restarts; stats->levels = (int)H.lv.size(); stats->converged = converged;
        stats->rel_residual = rel; stats->setup_ms = t_setup - t_begin; stats->solve_ms = t_end - t_setup;
        stats->operator_complexity = (double)total_nnz / (double)(9 * n);
        for (size_t l = 0; l < H.lv.size() && l < 24; ++l) stats->level_rows[l] = H.lv[l].n;
        stats->coarsest_rows = H.lv.back().n;
        stats->workspace_bytes = (int64_t)pool.bytes;
    }
    if (!converged) {
        set_error("ssrs_potential_solve: not converged after %d iterations (relative residual %.3e, target %.1e)", iters, rel, rtol);
        return SSRS_ERR_NOT_CONVERGED;
    }
    return SSRS_OK;
}
}}  // namespace ssrs::amg

extern "C" __attribute__((visibility("default")))
int SOLVE_NAME(const float* K, int rows, int cols, const int64_t* bnodes_host, const double* bvalues_host,
               int64_t n_bnodes, double rtol, int max_iter, float* phi, ssrs_solve_stats* stats, void* stream) {
    return ssrs::amg::solve_impl(K, rows, cols, bnodes_host, bvalues_host, n_bnodes, rtol, max_iter, phi, stats, nullptr, stream);
}
#ifndef SSRS_HOST_EMU
extern "C" __attribute__((visibility("default")))
int ssrs_potential_solve_f64(const float* K, int rows, int cols, const int64_t* bnodes_host, const double* bvalues_host,
                             int64_t n_bnodes, double rtol, int max_iter, float* phi, double* phi64, ssrs_solve_stats* stats,
                             void* stream) {
    return ssrs::amg::solve_impl(K, rows, cols, bnodes_host, bvalues_host, n_bnodes, rtol, max_iter, phi, stats, nullptr, stream, phi64);
}
extern "C" __attribute__((visibility("default")))
int ssrs_reserve_workspace(int rows, int cols) {
    // what the first solve of this size would otherwise allocate inside its timed path (two cudaMalloc of ~10 and ~5 GB at
    // 5000 x 6000: 0.2-1.2 s)
    if (rows <= 0 || cols <= 0) { set_error("ssrs_reserve_workspace: bad grid %dx%d", rows, cols); return SSRS_ERR_INVALID; }
    if (!ssrs::par::arena(false).idle() || !ssrs::par::arena(true).idle()) { set_error("ssrs_reserve_workspace: a solve is in progress"); return SSRS_ERR_INVALID; }
    const size_t n = (size_t)rows * (size_t)cols;
    ssrs::par::arena(false).reserve(n * 330 + ((size_t)64 << 20));
    ssrs::par::arena(true).reserve(n * 170 + ((size_t)64 << 20));
    return SSRS_OK;
}
extern "C" __attribute__((visibility("default")))
int ssrs_release_workspace(void) {
    if (!ssrs::par::arena(false).idle() || !ssrs::par::arena(true).idle()) { set_error("ssrs_release_workspace: a solve is in progress"); return SSRS_ERR_INVALID; }
    ssrs::par::arena(false).free_all();
    ssrs::par::arena(true).free_all();
    return SSRS_OK;
}
#endif
extern "C" __attribute__((visibility("default")))
int SOLVE_SHARDED_NAME(const float* K, int rows, int cols, const int64_t* bnodes_host, const double* bvalues_host,
                       int64_t n_bnodes, double rtol, int max_iter, float* phi, ssrs_solve_stats* stats,
                       const ssrs_comm* comm, void* stream) {
    return ssrs::amg::solve_impl(K, rows, cols, bnodes_host, bvalues_host, n_bnodes, rtol, max_iter, phi, stats, comm, stream);
}
