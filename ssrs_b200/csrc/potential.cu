// Stage 2 — Brandes–Ombalski fluid-flow potential on the conductivity raster (sm_100a).
//
// Replaces MovModel.assemble_sparse_linear_system + solve_sparse_linear_system
// (ssrs/movmodel.py:59-84, :86-128): the reference builds the 8-neighbour conductance graph entry by
// entry in Python, row-normalises it and hands it to SuperLU.  Here the same linear system
//     sum_j g_ij (phi_i - phi_j) = 0  at free nodes,   phi = given at Dirichlet nodes,
//     g_ij = hm(K_i, K_j) / f_ij,  hm = 1e-8 if K_i == 0 or K_j == 0 else 2/(1/K_i + 1/K_j),
//     f_ij = 1 (axial) | float32(sqrt 2) (diagonal), with the reference's last-column S/SW factor swap
//  is solved without ever forming the fine matrix:
//   * fine level: matrix-free operator evaluated from K (4 B/cell) in float64;
//   * preconditioner: aggregation AMG.  The conductances span ten orders of magnitude (half the cells sit
//     on the 1e-8 floor), so coarsening must follow the strong couplings: per level one pass of pairwise
//     matching by mutual strongest connection (handshake rounds), then unmatched nodes join the aggregate
//     of their strongest neighbour.  Prolongation is piecewise constant, coarse operators are Galerkin
//     sums (CSR), smoothing is weighted Jacobi, V(2,2) cycle, dense inverse on the coarsest level;
//   * outer iteration: right-preconditioned BiCGStab in float64 (the last-column quirk makes the operator
//     non-symmetric), true-residual restarts; result rounded to float32 like the reference (:128).
// Geometric multigrid (also operator-dependent/BoxMG interpolation) stalls on this problem: conducting
// islands that contain no coarse point lose their constant mode.  See DESIGN.md.
//
// All kernels are lambdas over pfor()/preduce() (pfor.cuh).
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

// ---- graph providers ----------------------------------------------------------------------------
// entry(i, k, j, a): k-th stored entry of row i -> kind, column j, value a (off-diagonals are negative).
struct FineGraph {
    const float* kd;     // conductivity with the sign bit marking Dirichlet nodes
    int rows, cols;
    int interior_dirichlet;   // 1 if some Dirichlet node is more than one cell away from the border
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
    SSRS_HD i64 size() const { return n; }
    SSRS_HD bool excluded(i64) const { return false; }
    SSRS_HD void range(i64 i, i64& k0, i64& k1) const { k0 = rowptr[i]; k1 = rowptr[i + 1]; }
    SSRS_HD int entry(i64 i, i64 k, i64& j, double& a) const {
        j = col[k];
        a = val[k];
        return j == i ? E_DIAG : E_OFF;
    }
};

// (A x)_i = excess * x_i + sum_off a_ij (x_j - x_i), excess = a_ii + sum_off a_ij  (difference form keeps
// the cancellation exact on the near-constant potentials).  with_dir: include Dirichlet neighbours (true
// operator); otherwise they only load the diagonal (error equation, e = 0 there).
template <class G>
SSRS_HD inline void row_eval(const G& g, i64 i, const double* x, bool with_dir, double& ax, double& diag) {
    i64 k0, k1;
    g.range(i, k0, k1);
    const double xi = x[i];
    double s = 0.0, offsum = 0.0, d = 0.0, dirsum = 0.0;
    bool have_diag = false;
    for (i64 k = k0; k < k1; ++k) {
        i64 j; double a;
        const int kind = g.entry(i, k, j, a);
        if (kind == E_OFF) { s += a * (x[j] - xi); offsum += a; }
        else if (kind == E_DIR) { if (with_dir) s += a * (x[j] - xi); dirsum += a; }
        else if (kind == E_DIAG) { d = a; have_diag = true; }
    }
    if (!have_diag) d = -(offsum + dirsum);          // fine level: zero row sum over all 8 links
    diag = d;
    const double excess = with_dir ? (d + offsum + dirsum) : (d + offsum);
    ax = excess * xi + s;
}

// ---- fine level, specialised ------------------------------------------------------------------------
// The generic providers above are used for setup.  The cycle and Krylov kernels of the fine level (most of
// the solve time) evaluate a cell's eight links with compile-time offsets instead.  FAST selects float32
// link weights (one MUFU reciprocal per link) for the preconditioner, where exactness is not required;
// the outer iteration always uses the float64 weights, so the solution is that of the exact operator.
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

// ax = (A x)_i and diag = a_ii for cell i of the fine grid.  WITH_DIR: Dirichlet neighbours contribute their
// value (true operator); otherwise they only load the diagonal (error equation).
template <bool FAST, bool WITH_DIR>
SSRS_HD inline void fine_row(const FineGraph& g, i64 i, const double* x, double& ax, double& diag) {
    const int cols = g.cols, rows = g.rows;
    const int r = (int)(i / cols), c = (int)(i - (i64)r * cols);
    const float kc = g.kd[i];
    const double xi = x[i];
    const bool quirk = (c == cols - 1) && (r >= 1) && (r <= rows - 2);     // movmodel.py:73-79
    double s = 0.0, d = 0.0;
#define SSRS_LINK(DR, DC)                                                                       \
    if ((DR < 0 ? r > 0 : (DR > 0 ? r < rows - 1 : true)) && (DC < 0 ? c > 0 : (DC > 0 ? c < cols - 1 : true))) { \
        const i64 j = i + (i64)(DR) * cols + (DC);                                              \
        const float kj = g.kd[j];                                                               \
        bool dg = (DR != 0) && (DC != 0);                                                       \
        if (DR == -1 && DC == 0 && quirk) dg = true;                                            \
        if (DR == -1 && DC == -1 && quirk) dg = false;                                          \
        const double w = link_weight<FAST>(kc, kj, dg);                                         \
        d += w;                                                                                 \
        if (WITH_DIR || !sign_set(kj)) s += w * (xi - x[j]);                                    \
    }
    SSRS_LINK(-1, -1) SSRS_LINK(-1, 0) SSRS_LINK(-1, 1)
    SSRS_LINK(0, -1)                    SSRS_LINK(0, 1)
    SSRS_LINK(1, -1)  SSRS_LINK(1, 0)  SSRS_LINK(1, 1)
#undef SSRS_LINK
    // error equation: links to Dirichlet neighbours act on (x_i - 0)
    if (!WITH_DIR) {
        double ddir = 0.0;
#define SSRS_DIRLINK(DR, DC)                                                                    \
        if ((DR < 0 ? r > 0 : (DR > 0 ? r < rows - 1 : true)) && (DC < 0 ? c > 0 : (DC > 0 ? c < cols - 1 : true))) { \
            const float kj = g.kd[i + (i64)(DR) * cols + (DC)];                                 \
            if (sign_set(kj)) {                                                                 \
                bool dg = (DR != 0) && (DC != 0);                                               \
                if (DR == -1 && DC == 0 && quirk) dg = true;                                    \
                if (DR == -1 && DC == -1 && quirk) dg = false;                                  \
                ddir += link_weight<FAST>(kc, kj, dg);                                          \
            }                                                                                   \
        }
        // Dirichlet nodes sit on the border only when produced by get_boundary_nodes; test cheaply
        if (r <= 1 || c <= 1 || r >= rows - 2 || c >= cols - 2 || g.interior_dirichlet) {
            SSRS_DIRLINK(-1, -1) SSRS_DIRLINK(-1, 0) SSRS_DIRLINK(-1, 1)
            SSRS_DIRLINK(0, -1)                       SSRS_DIRLINK(0, 1)
            SSRS_DIRLINK(1, -1)  SSRS_DIRLINK(1, 0)  SSRS_DIRLINK(1, 1)
        }
#undef SSRS_DIRLINK
        s += ddir * xi;
    }
    ax = s;
    diag = d;
}

// Precomputed forward link weights (E, N, NE, NW of every cell; 0 where the neighbour is outside the grid) in
// float32 (preconditioner) and float64 (exact operator).  A cell's other four links are its neighbours'
// forward links — except on the reference's quirk column (last column, interior rows) whose S and SW links
// carry swapped distance factors and are evaluated from K directly.  All eight neighbours are included: for
// error-equation vectors the caller keeps x = 0 at Dirichlet nodes, which is exactly "only load the diagonal".
struct FineWeights {
    const float* wf;    // [4][n]
    const double* wd;   // [4][n]
};

template <bool FAST>
SSRS_HD inline void fine_row_w(const FineGraph& g, const FineWeights& W, i64 i, const double* x, double& ax, double& diag) {
    const i64 n = (i64)g.rows * g.cols;
    const int cols = g.cols;
    const int r = (int)(i / cols), c = (int)(i - (i64)r * cols);
    const double xi = x[i];
    double s = 0.0, d = 0.0;
    // backward neighbours index below i, forward above; clamp so that zero-weight out-of-grid links stay in bounds
    const i64 jW = i - 1 < 0 ? 0 : i - 1, jS = i - cols < 0 ? 0 : i - cols;
    const i64 jSW = i - cols - 1 < 0 ? 0 : i - cols - 1, jSE = i - cols + 1 < 0 ? 0 : i - cols + 1;
    const i64 jE = i + 1 >= n ? n - 1 : i + 1, jN = i + cols >= n ? n - 1 : i + cols;
    const i64 jNE = i + cols + 1 >= n ? n - 1 : i + cols + 1, jNW = i + cols - 1 >= n ? n - 1 : i + cols - 1;
    double wE, wN, wNE, wNW, wW, wS, wSW, wSE;
    if (FAST) {
        const float* w = W.wf;
        wE = w[i]; wN = w[n + i]; wNE = w[2 * n + i]; wNW = w[3 * n + i];
        wW = (i - 1 >= 0) ? w[jW] : 0.0f; wS = (i - cols >= 0) ? w[n + jS] : 0.0f;
        wSW = (i - cols - 1 >= 0) ? w[2 * n + jSW] : 0.0f; wSE = (i - cols + 1 >= 0) ? w[3 * n + jSE] : 0.0f;
    } else {
        const double* w = W.wd;
        wE = w[i]; wN = w[n + i]; wNE = w[2 * n + i]; wNW = w[3 * n + i];
        wW = (i - 1 >= 0) ? w[jW] : 0.0; wS = (i - cols >= 0) ? w[n + jS] : 0.0;
        wSW = (i - cols - 1 >= 0) ? w[2 * n + jSW] : 0.0; wSE = (i - cols + 1 >= 0) ? w[3 * n + jSE] : 0.0;
    }
    if (c == cols - 1 && r >= 1 && r <= g.rows - 2) {                       // movmodel.py:73-79
        wS = link_weight<FAST>(g.kd[i], g.kd[i - cols], true);
        wSW = link_weight<FAST>(g.kd[i], g.kd[i - cols - 1], false);
    }
    // a cell in column 0 must not see its "W" neighbour's wrapped weight: forward weights of out-of-grid links
    // are stored as 0, and W of column 0 reads the E weight of the previous row's last cell, which is 0.
    d = ((wE + wW) + (wN + wS)) + ((wNE + wSW) + (wNW + wSE));
    s = wE * (xi - x[jE]) + wW * (xi - x[jW]) + wN * (xi - x[jN]) + wS * (xi - x[jS]) +
        wNE * (xi - x[jNE]) + wSW * (xi - x[jSW]) + wNW * (xi - x[jNW]) + wSE * (xi - x[jSE]);
    ax = s;
    diag = d;
}

// ---- storage ------------------------------------------------------------------------------------
struct Pool {
    std::vector<void*> ptrs;
    size_t bytes = 0;
    stream_t st;
    explicit Pool(stream_t s) : st(s) {}
    template <class T> T* get(i64 count) {
        void* p = nullptr;
        if (dev_alloc(&p, sizeof(T) * (size_t)(count > 0 ? count : 1), st) != 0) return nullptr;
        ptrs.push_back(p);
        bytes += sizeof(T) * (size_t)count;
        return static_cast<T*>(p);
    }
    void release(void* p) {
        for (size_t k = 0; k < ptrs.size(); ++k)
            if (ptrs[k] == p) { dev_free(p, st); ptrs.erase(ptrs.begin() + k); return; }
    }
    ~Pool() { for (void* p : ptrs) dev_free(p, st); }
};

struct Level {
    i64 n = 0, nnz = 0;
    i64* rowptr = nullptr; int* col = nullptr; double* val = nullptr;       // CSR (levels >= 1)
    int* agg = nullptr; i64 nc = 0; i64* memptr = nullptr; int* mem = nullptr;  // map to the next level
    double *x = nullptr, *b = nullptr, *t = nullptr;                            // cycle vectors (levels >= 1)
};

#define AMG_TRY(expr) do { if ((expr) != 0) { set_error("ssrs_potential_solve: device operation failed: %s", #expr); return SSRS_ERR_CUDA; } } while (0)
#define AMG_ALLOC(var, T, count) do { var = pool.get<T>(count); if (!var) { set_error("ssrs_potential_solve: out of device memory (%lld x %zu B)", (long long)(count), sizeof(T)); return SSRS_ERR_CUDA; } } while (0)

// ---- coarsening: pairwise matching + joins ---------------------------------------------------------
template <class G>
int coarsen(const G g, Level& L, Pool& pool, double theta, int rounds, int join_rounds, stream_t st) {
    const i64 n = g.size();
    float* rowmax; int *mate, *best, *root, *root2;
    Pool tmp(st);
    rowmax = tmp.get<float>(n); mate = tmp.get<int>(n); best = tmp.get<int>(n); root = tmp.get<int>(n); root2 = tmp.get<int>(n);
    if (!rowmax || !mate || !best || !root || !root2) { set_error("ssrs_potential_solve: out of device memory in coarsen"); return SSRS_ERR_CUDA; }
    const float th = (float)theta;
    AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
        float m = 0.0f;
        if (!g.excluded(i)) {
            i64 k0, k1; g.range(i, k0, k1);
            for (i64 k = k0; k < k1; ++k) { i64 j; double a; if (g.entry(i, k, j, a) == E_OFF) { float w = (float)(-a); if (w > m) m = w; } }
        }
        rowmax[i] = m;
        mate[i] = g.excluded(i) ? -2 : -1;
    }));
    for (int rnd = 0; rnd < rounds; ++rnd) {
        AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
            int bj = -1;
            if (mate[i] == -1) {
                float bw = 0.0f; unsigned bh = 0;
                const float rmi = rowmax[i];
                i64 k0, k1; g.range(i, k0, k1);
                for (i64 k = k0; k < k1; ++k) {
                    i64 j; double a;
                    if (g.entry(i, k, j, a) != E_OFF) continue;
                    const float w = (float)(-a);
                    if (!(w > 0.0f) || mate[j] != -1) continue;
                    const float rmj = rowmax[j];
                    if (w < th * (rmi > rmj ? rmi : rmj)) continue;           // strong from both sides
                    const unsigned h = pair_hash((unsigned)i, (unsigned)j);
                    if (bj < 0 || w > bw || (w == bw && (h > bh || (h == bh && (int)j > bj)))) { bw = w; bh = h; bj = (int)j; }
                }
            }
            best[i] = bj;
        }));
        AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
            const int b = best[i];
            if (b >= 0 && best[b] == (int)i) mate[i] = b;
        }));
    }
    AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
        const int m = mate[i];
        root[i] = m >= 0 ? ((int)i < m ? (int)i : m) : (m == -2 ? -2 : -1);
    }));
    for (int jr = 0; jr < join_rounds; ++jr) {
        AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
            int out = root[i];
            if (out == -1) {
                int bj = -1, btwo = 0; float bw = 0.0f; unsigned bh = 0;
                const float rmi = rowmax[i];
                i64 k0, k1; g.range(i, k0, k1);
                for (i64 k = k0; k < k1; ++k) {
                    i64 j; double a;
                    if (g.entry(i, k, j, a) != E_OFF) continue;
                    const float w = (float)(-a);
                    if (!(w > 0.0f) || root[j] < 0 || w < th * rmi) continue;  // strong for i, target already aggregated
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
        AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
            if (rt[i] == -1) rt[i] = (int)i;                  // singleton
            flag[i] = (rt[i] == (int)i) ? 1 : 0;
        }));
    }
    i64 nc = 0;
    AMG_TRY(exclusive_scan_i64(flag, n, &nc, st));
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
        AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
            const int r = rt[i];
            const int c = r >= 0 ? (int)flag[r] : -1;
            agg[i] = c;
            if (c >= 0) atomic_add_int(cnt + c, 1);
        }));
    }
    AMG_TRY(pfor(nc + 1, st, [=] SSRS_HD(i64 I) { memptr[I] = I < nc ? (i64)cnt[I] : 0; }));
    i64 total = 0;
    AMG_TRY(exclusive_scan_i64(memptr, nc + 1, &total, st));
    AMG_TRY(dev_zero(cnt, sizeof(int) * (size_t)(nc + 1), st));
    AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
        const int c = agg[i];
        if (c >= 0) { const int slot = atomic_add_int(cnt + c, 1); mem[memptr[c] + slot] = (int)i; }
    }));
    // members in ascending order -> deterministic Galerkin sums
    AMG_TRY(pfor(nc, st, [=] SSRS_HD(i64 I) {
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
template <class G>
int galerkin(const G g, const Level& L, Level& C, Pool& pool, stream_t st) {
    const i64 nc = L.nc;
    const int* agg = L.agg; const i64* memptr = L.memptr; const int* mem = L.mem;
    Pool tmp(st);
    i64* off = tmp.get<i64>(nc + 1);
    int* len = tmp.get<int>(nc);
    if (!off || !len) { set_error("ssrs_potential_solve: out of device memory in galerkin"); return SSRS_ERR_CUDA; }
    AMG_TRY(pfor(nc + 1, st, [=] SSRS_HD(i64 I) {
        i64 ub = 0;
        if (I < nc)
            for (i64 p = memptr[I]; p < memptr[I + 1]; ++p) { i64 k0, k1; g.range(mem[p], k0, k1); ub += (k1 - k0) + 1; }
        off[I] = ub;
    }));
    i64 scratch_n = 0;
    AMG_TRY(exclusive_scan_i64(off, nc + 1, &scratch_n, st));
    int* scol = tmp.get<int>(scratch_n);
    double* sval = tmp.get<double>(scratch_n);
    if (!scol || !sval) { set_error("ssrs_potential_solve: out of device memory in galerkin (%lld entries)", (long long)scratch_n); return SSRS_ERR_CUDA; }
    AMG_TRY(pfor(nc, st, [=] SSRS_HD(i64 I) {
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
    AMG_TRY(pfor(nc + 1, st, [=] SSRS_HD(i64 I) { rowptr[I] = I < nc ? (i64)len[I] : 0; }));
    i64 nnz = 0;
    AMG_TRY(exclusive_scan_i64(rowptr, nc + 1, &nnz, st));
    AMG_ALLOC(C.col, int, nnz);
    AMG_ALLOC(C.val, double, nnz);
    int* col = C.col; double* val = C.val;
    AMG_TRY(pfor(nc, st, [=] SSRS_HD(i64 I) {
        const i64 s = off[I], d = rowptr[I];
        for (int q = 0; q < len[I]; ++q) { col[d + q] = scol[s + q]; val[d + q] = sval[s + q]; }
    }));
    C.n = nc; C.nnz = nnz;
    AMG_ALLOC(C.x, double, nc);
    AMG_ALLOC(C.b, double, nc);
    AMG_ALLOC(C.t, double, nc);
    AMG_TRY(sync(st));
    return SSRS_OK;
}

// ---- cycle kernels ----------------------------------------------------------------------------------
template <class G>
int jacobi_first(const G g, const double* b, double* x, double omega, stream_t st) {
    return pfor(g.size(), st, [=] SSRS_HD(i64 i) {
        if (g.excluded(i)) { x[i] = 0.0; return; }
        i64 k0, k1; g.range(i, k0, k1);
        double d = 0.0, links = 0.0; bool have = false;
        for (i64 k = k0; k < k1; ++k) {
            i64 j; double a; const int kind = g.entry(i, k, j, a);
            if (kind == E_DIAG) { d = a; have = true; }
            else if (kind == E_OFF || kind == E_DIR) links += a;
        }
        if (!have) d = -links;
        x[i] = omega * b[i] / d;
    });
}
template <class G>
int jacobi(const G g, const double* b, const double* x, double* xn, double omega, stream_t st) {
    return pfor(g.size(), st, [=] SSRS_HD(i64 i) {
        if (g.excluded(i)) { xn[i] = 0.0; return; }
        double ax, d;
        row_eval(g, i, x, false, ax, d);
        xn[i] = x[i] + omega * (b[i] - ax) / d;
    });
}
template <class G>
int restrict_residual(const G g, const Level& L, const double* b, const double* x, double* bc, stream_t st) {
    const i64* memptr = L.memptr; const int* mem = L.mem;
    return pfor(L.nc, st, [=] SSRS_HD(i64 I) {
        double s = 0.0;
        for (i64 p = memptr[I]; p < memptr[I + 1]; ++p) {
            const i64 i = mem[p];
            double ax, d;
            row_eval(g, i, x, false, ax, d);
            s += b[i] - ax;
        }
        bc[I] = s;
    });
}
inline int prolong_add(const Level& L, i64 n, double* x, const double* xc, stream_t st) {
    const int* agg = L.agg;
    return pfor(n, st, [=] SSRS_HD(i64 i) { const int c = agg[i]; if (c >= 0) x[i] += xc[c]; });
}

struct Hierarchy {
    FineGraph fine;
    std::vector<Level> lv;     // lv[0] = fine level (agg/mem only), lv[l>=1] CSR
    double* cinv = nullptr;    // dense inverse of the coarsest operator
    i64 cn = 0;
    int coarse_sweeps = 0;     // > 0: coarsest level too large for a dense inverse, Jacobi sweeps instead
    double omega = 0.7;
    int nu = 2;
    double* fres = nullptr;    // fine-level residual scratch
    FineWeights fw = {nullptr, nullptr};
};

inline CsrGraph csr_of(const Level& L) { CsrGraph g; g.rowptr = L.rowptr; g.col = L.col; g.val = L.val; g.n = L.n; return g; }

template <class G>
int smooth(const G g, const double* b, double*& x, double*& t, int sweeps, bool zero_guess, double omega, stream_t st) {
    for (int s = 0; s < sweeps; ++s) {
        if (s == 0 && zero_guess) { AMG_TRY(jacobi_first(g, b, x, omega, st)); }
        else { AMG_TRY(jacobi(g, b, x, t, omega, st)); double* sw = x; x = t; t = sw; }
    }
    return SSRS_OK;
}

// fine-level specialisations of the cycle kernels (float32 link weights: preconditioner only)
inline int fine_jacobi_first(const FineGraph g, const FineWeights W, const double* b, double* x, double omega, stream_t st) {
    return pfor(g.size(), st, [=] SSRS_HD(i64 i) {
        if (g.excluded(i)) { x[i] = 0.0; return; }
        double ax, d;
        fine_row_w<true>(g, W, i, b, ax, d);          // only the diagonal is used
        x[i] = omega * b[i] / d;
    });
}
inline int fine_jacobi(const FineGraph g, const FineWeights W, const double* b, const double* x, double* xn, double omega, stream_t st) {
    return pfor(g.size(), st, [=] SSRS_HD(i64 i) {
        if (g.excluded(i)) { xn[i] = 0.0; return; }
        double ax, d;
        fine_row_w<true>(g, W, i, x, ax, d);
        xn[i] = x[i] + omega * (b[i] - ax) / d;
    });
}
// residual per cell (coalesced) into `res`, then a deterministic per-aggregate sum
inline int fine_restrict_residual(const FineGraph g, const FineWeights W, const Level& L, const double* b, const double* x,
                                  double* res, double* bc, stream_t st) {
    if (pfor(g.size(), st, [=] SSRS_HD(i64 i) {
            if (g.excluded(i)) { res[i] = 0.0; return; }
            double ax, d;
            fine_row_w<true>(g, W, i, x, ax, d);
            res[i] = b[i] - ax;
        }) != 0) return -1;
    const i64* memptr = L.memptr; const int* mem = L.mem;
    return pfor(L.nc, st, [=] SSRS_HD(i64 I) {
        double s = 0.0;
        for (i64 p = memptr[I]; p < memptr[I + 1]; ++p) s += res[mem[p]];
        bc[I] = s;
    });
}
inline int fine_smooth(const FineGraph g, const FineWeights W, const double* b, double*& x, double*& t, int sweeps,
                       bool zero_guess, double omega, stream_t st) {
    for (int s = 0; s < sweeps; ++s) {
        if (s == 0 && zero_guess) { AMG_TRY(fine_jacobi_first(g, W, b, x, omega, st)); }
        else { AMG_TRY(fine_jacobi(g, W, b, x, t, omega, st)); double* sw = x; x = t; t = sw; }
    }
    return SSRS_OK;
}

int coarse_solve(Hierarchy& H, Level& C, stream_t st) {
    if (H.coarse_sweeps > 0) return smooth(csr_of(C), C.b, C.x, C.t, H.coarse_sweeps, true, H.omega, st);
    const double* inv = H.cinv; const double* b = C.b; double* x = C.x; const i64 n = C.n;
    return pfor(n, st, [=] SSRS_HD(i64 i) {
        double s = 0.0;
        for (i64 j = 0; j < n; ++j) s += inv[i * n + j] * b[j];
        x[i] = s;
    });
}

// out = M^-1 rhs (one V-cycle from a zero guess).  `out`/`tmp` are fine-level buffers; on return `out`
// holds the result (the two may have been swapped).
int vcycle(Hierarchy& H, const double* rhs, double*& out, double*& tmp, stream_t st) {
    const int nl = (int)H.lv.size();
    if (nl == 1) {          // no coarse level: plain Jacobi sweeps
        return fine_smooth(H.fine, H.fw, rhs, out, tmp, 2 * H.nu, true, H.omega, st);
    }
    int rc = fine_smooth(H.fine, H.fw, rhs, out, tmp, H.nu, true, H.omega, st);
    if (rc) return rc;
    AMG_TRY(fine_restrict_residual(H.fine, H.fw, H.lv[0], rhs, out, H.fres, H.lv[1].b, st));
    for (int l = 1; l < nl - 1; ++l) {
        Level& L = H.lv[l];
        rc = smooth(csr_of(L), L.b, L.x, L.t, H.nu, true, H.omega, st);
        if (rc) return rc;
        AMG_TRY(restrict_residual(csr_of(L), L, L.b, L.x, H.lv[l + 1].b, st));
    }
    rc = coarse_solve(H, H.lv[nl - 1], st);
    if (rc) return rc;
    for (int l = nl - 2; l >= 1; --l) {
        Level& L = H.lv[l];
        AMG_TRY(prolong_add(L, L.n, L.x, H.lv[l + 1].x, st));
        rc = smooth(csr_of(L), L.b, L.x, L.t, H.nu, false, H.omega, st);
        if (rc) return rc;
    }
    AMG_TRY(prolong_add(H.lv[0], H.fine.size(), out, H.lv[1].x, st));
    return fine_smooth(H.fine, H.fw, rhs, out, tmp, H.nu, false, H.omega, st);
}

int dense_inverse(Hierarchy& H, const Level& C, Pool& pool, stream_t st) {
    const i64 n = C.n;
    double *D, *I, *colk, *rowD, *rowI;
    Pool tmp(st);
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

// out = b - A x at free nodes (b = 0 there: the Dirichlet values live in x), 0 at Dirichlet nodes; *nrm2 = |out|^2
int fine_residual(const FineGraph fg, const FineWeights W, const double* x, double* out, double* nrm2, stream_t st) {
    const i64 n = fg.size();
    AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
        if (fg.excluded(i)) { out[i] = 0.0; return; }
        double ax, d; fine_row_w<false>(fg, W, i, x, ax, d);
        out[i] = -ax;
    }));
    AMG_TRY(preduce_sum(n, st, nrm2, [=] SSRS_HD(i64 i) { return out[i] * out[i]; }));
    return SSRS_OK;
}
// out = A in for a correction vector `in` (zero at Dirichlet nodes)
int fine_apply(const FineGraph fg, const FineWeights W, const double* in, double* out, stream_t st) {
    AMG_TRY(pfor(fg.size(), st, [=] SSRS_HD(i64 i) {
        if (fg.excluded(i)) { out[i] = 0.0; return; }
        double ax, d; fine_row_w<false>(fg, W, i, in, ax, d);
        out[i] = ax;
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
extern "C" __attribute__((visibility("default"))) const char* ssrs_emu_last_error(void) { return ssrs::g_emu_err; }
#else
#define SOLVE_NAME ssrs_potential_solve
#endif

namespace ssrs { namespace amg {
int solve_impl(const float* K, int rows, int cols, const int64_t* bnodes_host, const double* bvalues_host,
               int64_t n_bnodes, double rtol, int max_iter, float* phi, ssrs_solve_stats* stats, void* stream) {
    if (K == nullptr || phi == nullptr) { set_error("ssrs_potential_solve: NULL raster"); return SSRS_ERR_INVALID; }
    if (rows < 3 || cols < 3) { set_error("ssrs_potential_solve: grid %dx%d too small", rows, cols); return SSRS_ERR_INVALID; }
    if ((int64_t)rows * cols > 2000000000LL) { set_error("ssrs_potential_solve: more than 2e9 cells"); return SSRS_ERR_UNSUPPORTED; }
    if (n_bnodes <= 0 || bnodes_host == nullptr || bvalues_host == nullptr) { set_error("ssrs_potential_solve: no Dirichlet nodes"); return SSRS_ERR_INVALID; }
    if (!(rtol > 0.0)) rtol = 1e-9;
    if (max_iter <= 0) max_iter = 300;
    stream_t st = (stream_t)stream;
    const i64 n = (i64)rows * cols;
    const double t_begin = now_ms();
#ifndef SSRS_HOST_EMU
    {   // keep freed workspace cached in the device's default pool between solves
        int dev = 0; cudaMemPool_t mp;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) {
            uint64_t thresh = ~0ULL;
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thresh);
        }
    }
#endif
    Pool pool(st);
    Hierarchy H;

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
    double *x, *r, *rh, *p, *v, *s, *t, *y, *y2, *z, *z2;
    AMG_ALLOC(x, double, n); AMG_ALLOC(r, double, n); AMG_ALLOC(rh, double, n); AMG_ALLOC(p, double, n);
    AMG_ALLOC(v, double, n); AMG_ALLOC(s, double, n); AMG_ALLOC(t, double, n);
    AMG_ALLOC(y, double, n); AMG_ALLOC(y2, double, n); AMG_ALLOC(z, double, n); AMG_ALLOC(z2, double, n);
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
    AMG_ALLOC(H.fres, double, n);
    {
        float* wf; double* wd;
        AMG_ALLOC(wf, float, 4 * n);
        AMG_ALLOC(wd, double, 4 * n);
        const FineGraph fgw = H.fine;
        AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) {
            const int r = (int)(i / cols), c = (int)(i - (i64)r * cols);
            const float kc = fgw.kd[i];
            const bool hasE = c < cols - 1, hasN = r < rows - 1, hasW = c > 0;
            const double e = hasE ? link_weight<false>(kc, fgw.kd[i + 1], false) : 0.0;
            const double nn = hasN ? link_weight<false>(kc, fgw.kd[i + cols], false) : 0.0;
            const double ne = (hasN && hasE) ? link_weight<false>(kc, fgw.kd[i + cols + 1], true) : 0.0;
            const double nw = (hasN && hasW) ? link_weight<false>(kc, fgw.kd[i + cols - 1], true) : 0.0;
            wd[i] = e; wd[n + i] = nn; wd[2 * n + i] = ne; wd[3 * n + i] = nw;
            wf[i] = (float)e; wf[n + i] = (float)nn; wf[2 * n + i] = (float)ne; wf[3 * n + i] = (float)nw;
        }));
        H.fw.wf = wf; H.fw.wd = wd;
    }

    // ---- setup: hierarchy ----
    const double theta = 0.5;
    const i64 coarse_target = 400, dense_cap = 2048;
    H.lv.emplace_back();
    H.lv[0].n = n;
    i64 total_nnz = 9 * n;
    for (int l = 0; l < 40; ++l) {
        Level& L = H.lv[(size_t)l];
        if (l > 0 && L.n <= coarse_target) break;
        int rc = (l == 0) ? coarsen(H.fine, L, pool, theta, 8, 3, st) : coarsen(csr_of(L), L, pool, theta, 8, 3, st);
        if (rc) return rc;
        if (L.nc < 1 || (double)L.nc > 0.9 * (double)L.n) {        // stalled: stop here
            pool.release(L.agg); pool.release(L.memptr); pool.release(L.mem);
            L.agg = nullptr; L.memptr = nullptr; L.mem = nullptr; L.nc = 0;
            break;
        }
        Level C;
        rc = (l == 0) ? galerkin(H.fine, L, C, pool, st) : galerkin(csr_of(L), L, C, pool, st);
        if (rc) return rc;
        total_nnz += C.nnz;
        H.lv.push_back(C);
    }
    {
        Level& C = H.lv.back();
        if (H.lv.size() > 1) {
            if (C.n <= dense_cap) { int rc = dense_inverse(H, C, pool, st); if (rc) return rc; }
            else H.coarse_sweeps = 60;
        }
    }
    const double t_setup = now_ms();

    // ---- BiCGStab, right preconditioned ----
    const FineGraph fg = H.fine;
    double r0n2 = 0.0;
    { int rc = fine_residual(fg, H.fw, x, r, &r0n2, st); if (rc) return rc; }
    const double r0 = sqrt(r0n2);
    int iters = 0, restarts = 0, converged = (r0 == 0.0);
    double best_true = 1.0;
    double rel = (r0 == 0.0) ? 0.0 : 1.0;
    while (!converged && iters < max_iter && restarts <= 6) {
        AMG_TRY(copy_d2d(rh, r, sizeof(double) * (size_t)n, st));
        AMG_TRY(dev_zero(p, sizeof(double) * (size_t)n, st));
        AMG_TRY(dev_zero(v, sizeof(double) * (size_t)n, st));
        double rho = 1.0, alpha = 1.0, om = 1.0;
        bool breakdown = false;
        while (iters < max_iter) {
            double rho_new = 0.0;
            { const double *a_ = rh, *b_ = r; AMG_TRY(preduce_sum(n, st, &rho_new, [=] SSRS_HD(i64 i) { return a_[i] * b_[i]; })); }
            if (rho_new == 0.0 || !(fabs(rho_new) < 1e300)) { breakdown = true; break; }
            const double beta = (rho_new / rho) * (alpha / om);
            rho = rho_new;
            { double *pp = p; const double *rr = r, *vv = v; const double om_ = om;
              AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) { pp[i] = rr[i] + beta * (pp[i] - om_ * vv[i]); })); }
            { int rc = vcycle(H, p, y, y2, st); if (rc) return rc; }
            { int rc = fine_apply(fg, H.fw, y, v, st); if (rc) return rc; }
            double rhv = 0.0;
            { const double *a_ = rh, *b_ = v; AMG_TRY(preduce_sum(n, st, &rhv, [=] SSRS_HD(i64 i) { return a_[i] * b_[i]; })); }
            if (rhv == 0.0) { breakdown = true; break; }
            alpha = rho / rhv;
            { double* ss = s; const double *rr = r, *vv = v; const double al = alpha;
              AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) { ss[i] = rr[i] - al * vv[i]; })); }
            { int rc = vcycle(H, s, z, z2, st); if (rc) return rc; }
            { int rc = fine_apply(fg, H.fw, z, t, st); if (rc) return rc; }
            double ts = 0.0, tt = 0.0;
            { const double *a_ = t, *b_ = s;
              AMG_TRY(preduce_sum2(n, st, &ts, &tt, [=] SSRS_HD(i64 i, double& u0, double& u1) { u0 = a_[i] * b_[i]; u1 = a_[i] * a_[i]; })); }
            om = (tt > 0.0) ? ts / tt : 0.0;
            double rn2 = 0.0;
            { double *xx = x, *rr = r; const double *yy = y, *zz = z, *ss = s, *tv = t; const double al = alpha, om_ = om;
              AMG_TRY(preduce_sum(n, st, &rn2, [=] SSRS_HD(i64 i) {
                  xx[i] += al * yy[i] + om_ * zz[i];
                  const double rv = ss[i] - om_ * tv[i];
                  rr[i] = rv;
                  return rv * rv;
              })); }
            ++iters;
            rel = sqrt(rn2) / r0;
            if (!(rel == rel)) { breakdown = true; break; }
            if (rel <= rtol || om == 0.0) break;
        }
        // true residual: accept, or restart from the current iterate
        double tn2 = 0.0;
        { int rc = fine_residual(fg, H.fw, x, r, &tn2, st); if (rc) return rc; }
        rel = sqrt(tn2) / r0;
        if (!(rel == rel)) { set_error("ssrs_potential_solve: NaN residual (NaN in the conductivity raster?)"); return SSRS_ERR_NOT_CONVERGED; }
        if (rel <= 4.0 * rtol) converged = 1;
        else if (!breakdown && rel > 0.5 * best_true && rel <= 1e-6) converged = 2;   // float64 attainable accuracy reached
        else ++restarts;
        if (rel < best_true) best_true = rel;
    }
    { const double* xx = x; AMG_TRY(pfor(n, st, [=] SSRS_HD(i64 i) { phi[i] = (float)xx[i]; })); }     // movmodel.py:128
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
    return ssrs::amg::solve_impl(K, rows, cols, bnodes_host, bvalues_host, n_bnodes, rtol, max_iter, phi, stats, stream);
}
