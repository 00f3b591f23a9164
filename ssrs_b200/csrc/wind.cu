// "Next" rows f-2 and f-3 of the hot-path scope table: the producers feeding stage 1.
//
// f-2  ssrs_interp_wind — scattered wind points (WTK sites) -> per-cell wind speed / direction rasters.
//      Replaces Simulator._get_interpolated_wind_conditions + _interpolate_wtk_vardata
//      (ssrs/simulator.py:765-792): u/v components interpolated with scipy griddata(method='linear'), i.e.
//      barycentric interpolation on the Qhull Delaunay triangulation of the points, NaN outside the hull, then
//      speed = hypot(u, v), direction = mod(atan2(u, v) + 2 pi, 2 pi) in degrees.  The triangulation (a few
//      hundred points) stays on the host (scipy.spatial.Delaunay = the same Qhull); the per-cell work — 3e7 to
//      1.2e8 cells per wind case — runs here: triangles are rasterised over their bounding boxes (a cell on a
//      shared edge goes to the lower triangle index, both sides interpolate the same value there), then every
//      cell evaluates its triangle in float64.  HBM-bound: 4 B index + 8 B output per cell.
//      wtk_interp_type = 'cubic' (griddata's Clough-Tocher scheme, clough_tocher.cuh) shares the rasterisation:
//      gradients at the sites by the same Gauss-Seidel sweeps as scipy (one warp per wind component; the sweep is
//      sequential over the sites by definition), 19 Bezier ordinates per triangle and component, then one cubic
//      per cell and component.
//
// f-3  ssrs_thermal_seeds + ssrs_gaussian_blur — compute_thermals (ssrs/layers.py:188-214): per cell inside the
//      10 % border a thermal is seeded with probability 1/(int(wtfactor) - 1), wtfactor = 1000 + |aspect-180|/180
//      * 2000 (np.random.randint(1, int(wtfactor)) == 5), with strength lognormal(scale + 3, 0.5); the seeds are
//      then smoothed by scipy.ndimage.gaussian_filter(sigma=4, mode='constant').  The reference draws from numpy's
//      global stream cell by cell, so only the distribution can be reproduced: draws here are
//      Philox4x32-10 keyed by (seed, cell).  The blur is deterministic and matches scipy's to float32 rounding.
#include "common.cuh"
#include "clough_tocher.cuh"

#include <math.h>

namespace ssrs {
namespace {

constexpr int NO_TRIANGLE = 0x7f7f7f7f;

struct Tri {
    double x2, y2;            // third vertex
    double a00, a01, a10, a11;   // inverse of [p0 - p2, p1 - p2]
    int i0, i1, i2;
    int ok;
};

__device__ __forceinline__ Tri load_tri(const double* px, const double* py, const int* tri, int t) {
    Tri T;
    T.i0 = tri[3 * t]; T.i1 = tri[3 * t + 1]; T.i2 = tri[3 * t + 2];
    const double x0 = px[T.i0], y0 = py[T.i0], x1 = px[T.i1], y1 = py[T.i1];
    T.x2 = px[T.i2]; T.y2 = py[T.i2];
    const double m00 = x0 - T.x2, m01 = x1 - T.x2, m10 = y0 - T.y2, m11 = y1 - T.y2;
    const double det = m00 * m11 - m01 * m10;
    T.ok = det != 0.0;
    const double inv = T.ok ? 1.0 / det : 0.0;
    T.a00 = m11 * inv; T.a01 = -m01 * inv; T.a10 = -m10 * inv; T.a11 = m00 * inv;
    return T;
}
__device__ __forceinline__ void barycentric(const Tri& T, double x, double y, double& c0, double& c1, double& c2) {
    const double dx = x - T.x2, dy = y - T.y2;
    c0 = T.a00 * dx + T.a01 * dy;
    c1 = T.a10 * dx + T.a11 * dy;
    c2 = 1.0 - c0 - c1;
}

// one CTA per triangle: mark the cells of its bounding box that lie inside (scipy's find_simplex tolerance)
__global__ void __launch_bounds__(256) rasterise_triangles_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                                                  const int* __restrict__ tri, int ntri, double x0, double y0,
                                                                  double res, int rows, int cols, int* __restrict__ owner) {
    const double eps = 100.0 * 2.220446049250313e-16;
    for (int t = blockIdx.x; t < ntri; t += gridDim.x) {
        const Tri T = load_tri(px, py, tri, t);
        if (!T.ok) continue;
        const double xa = px[T.i0], xb = px[T.i1], ya = py[T.i0], yb = py[T.i1];
        const double xmin = fmin(fmin(xa, xb), T.x2), xmax = fmax(fmax(xa, xb), T.x2);
        const double ymin = fmin(fmin(ya, yb), T.y2), ymax = fmax(fmax(ya, yb), T.y2);
        int c_lo = (int)floor((xmin - x0) / res) - 1, c_hi = (int)ceil((xmax - x0) / res) + 1;
        int r_lo = (int)floor((ymin - y0) / res) - 1, r_hi = (int)ceil((ymax - y0) / res) + 1;
        c_lo = max(c_lo, 0); r_lo = max(r_lo, 0); c_hi = min(c_hi, cols - 1); r_hi = min(r_hi, rows - 1);
        if (c_lo > c_hi || r_lo > r_hi) continue;
        const int w = c_hi - c_lo + 1;
        const long long cells = (long long)w * (r_hi - r_lo + 1);
        for (long long k = threadIdx.x; k < cells; k += blockDim.x) {
            const int r = r_lo + (int)(k / w), c = c_lo + (int)(k % w);
            double c0, c1, c2;
            barycentric(T, x0 + c * res, y0 + r * res, c0, c1, c2);
            if (c0 >= -eps && c1 >= -eps && c2 >= -eps) atomicMin(owner + (long long)r * cols + c, t);
        }
    }
}

__global__ void __launch_bounds__(256) interp_wind_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                                          const double* __restrict__ east, const double* __restrict__ north,
                                                          const int* __restrict__ tri, double x0, double y0, double res,
                                                          int rows, int cols, const int* __restrict__ owner,
                                                          float* __restrict__ wspeed, float* __restrict__ wdirn) {
    const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
    if (r >= rows || c >= cols) return;
    const long long i = (long long)r * cols + c;
    const int t = owner[i];
    float s = nanf(""), d = nanf("");                 // outside the convex hull: griddata's fill value
    if (t != NO_TRIANGLE) {
        const Tri T = load_tri(px, py, tri, t);
        double c0, c1, c2;
        barycentric(T, x0 + c * res, y0 + r * res, c0, c1, c2);
        const double e = c0 * east[T.i0] + c1 * east[T.i1] + c2 * east[T.i2];
        const double n = c0 * north[T.i0] + c1 * north[T.i1] + c2 * north[T.i2];
        const double two_pi = 6.283185307179586;
        s = (float)sqrt(e * e + n * n);                                         // simulator.py:787-788
        d = (float)(fmod(atan2(e, n) + two_pi, two_pi) * (180.0 / 3.141592653589793));   // :789-791
    }
    wspeed[i] = s;
    wdirn[i] = d;
}

// ---- Clough-Tocher ('cubic') ---------------------------------------------------------------------------
struct WarpLanes {                                 // the lanes of one warp share a vertex's neighbour loop
    __device__ int lane() const { return threadIdx.x; }
    __device__ int lanes() const { return 32; }
    __device__ double sum(double v) const {        // butterfly: every lane ends with the same bits (a + b == b + a)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ void sync() const { __syncwarp(); }
};

// CTA 0: easterly component, CTA 1: northerly.  grad: [2][npoints][2]; sweeps: [2] (0 = maxiter reached)
__global__ void __launch_bounds__(32) ct_gradients_kernel(const double* px, const double* py, const double* east,
                                                          const double* north, int npoints, const int* nb_indptr,
                                                          const int* nb_indices, int maxiter, double tol, double* grad,
                                                          int* sweeps) {
    WarpLanes w;
    const int comp = blockIdx.x;
    const int n = ct::estimate_gradients(w, px, py, comp == 0 ? east : north, npoints, nb_indptr, nb_indices, maxiter, tol,
                                         grad + (size_t)comp * 2 * npoints);
    if (threadIdx.x == 0) sweeps[comp] = n;
}

// one thread per (triangle, component): the triangle's 19 Bezier ordinates
__global__ void __launch_bounds__(128) ct_coefficients_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                                              const double* __restrict__ east, const double* __restrict__ north,
                                                              int npoints, const int* __restrict__ tri,
                                                              const int* __restrict__ neighbors, int ntri,
                                                              const double* __restrict__ grad, double* __restrict__ coef) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * ntri) return;
    const int t = k >> 1, comp = k & 1;
    const Tri T = load_tri(px, py, tri, t);
    const int v[3] = {T.i0, T.i1, T.i2};
    const double* val = comp == 0 ? east : north;
    const double* g = grad + (size_t)comp * 2 * npoints;
    double p[3][2], f[3], df[3][2], c[3][3];
    int has_nb[3];
    for (int j = 0; j < 3; ++j) {
        p[j][0] = px[v[j]]; p[j][1] = py[v[j]];
        f[j] = val[v[j]];
        df[j][0] = g[2 * v[j]]; df[j][1] = g[2 * v[j] + 1];
        const int nb = neighbors[3 * t + j];                 // the triangle across the side opposite vertex j
        has_nb[j] = nb >= 0;
        c[j][0] = c[j][1] = c[j][2] = 0.0;
        if (nb >= 0) {
            const int a = tri[3 * nb], b = tri[3 * nb + 1], d = tri[3 * nb + 2];
            barycentric(T, (px[a] + px[b] + px[d]) / 3, (py[a] + py[b] + py[d]) / 3, c[j][0], c[j][1], c[j][2]);
        }
    }
    ct::coefficients(p, f, df, has_nb, c, coef + (size_t)k * ct::COEF_STRIDE);
}

__global__ void __launch_bounds__(256) interp_wind_cubic_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                                                const int* __restrict__ tri, double x0, double y0, double res,
                                                                int rows, int cols, const int* __restrict__ owner,
                                                                const double* __restrict__ coef, float* __restrict__ wspeed,
                                                                float* __restrict__ wdirn) {
    const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
    if (r >= rows || c >= cols) return;
    const long long i = (long long)r * cols + c;
    const int t = owner[i];
    float s = nanf(""), d = nanf("");                 // outside the convex hull: griddata's fill value
    if (t != NO_TRIANGLE) {
        const Tri T = load_tri(px, py, tri, t);
        double b0, b1, b2;
        barycentric(T, x0 + c * res, y0 + r * res, b0, b1, b2);
        const double* ce = coef + (size_t)t * 2 * ct::COEF_STRIDE;       // neighbouring cells share the triangle: L1 hits
        const double e = ct::evaluate(ce, b0, b1, b2);
        const double n = ct::evaluate(ce + ct::COEF_STRIDE, b0, b1, b2);
        const double two_pi = 6.283185307179586;
        s = (float)sqrt(e * e + n * n);                                         // simulator.py:787-788
        d = (float)(fmod(atan2(e, n) + two_pi, two_pi) * (180.0 / 3.141592653589793));   // :789-791
    }
    wspeed[i] = s;
    wdirn[i] = d;
}

// ---- nearest-site interpolation (wtk_interp_type = 'nearest') -------------------------------------------
// scipy's griddata(method='nearest') is a k-d tree query: the value of the closest site in Euclidean distance,
// defined everywhere (no hull).  A CTA owns 32 x 8 cells.  With c the CTA's centre, R the largest distance from c to
// one of its cell centres and d_min the distance from c to its closest site, the closest site of ANY cell of the CTA
// lies within d_min + 2R of c (triangle inequality), so the CTA first filters the sites by that bound (a handful
// survive for sites a few hundred cells apart) and each cell then tests only the survivors, in float64 like the tree
// and with separately rounded products so that ties resolve identically; equal distances go to the lower site index.
constexpr int NEAREST_MAX_CAND = 256;

__device__ __forceinline__ double dist2(double ax, double ay, double bx, double by) {
    const double dx = ax - bx, dy = ay - by;
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

__global__ void __launch_bounds__(256) nearest_wind_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                                           const double* __restrict__ east, const double* __restrict__ north,
                                                           int npoints, double x0, double y0, double res, int rows, int cols,
                                                           float* __restrict__ wspeed, float* __restrict__ wdirn) {
    __shared__ double s_red[8];
    __shared__ double s_bound2;
    __shared__ int s_cnt;
    __shared__ int s_cand[NEAREST_MAX_CAND];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const double cx = x0 + (blockIdx.x * 32 + 15.5) * res, cy = y0 + (blockIdx.y * 8 + 3.5) * res;
    double m = 1.0e300;
    for (int j = tid; j < npoints; j += 256) m = fmin(m, dist2(cx, cy, px[j], py[j]));
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = m;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (tid == 0) {
        double mm = s_red[0];
        for (int w = 1; w < 8; ++w) mm = fmin(mm, s_red[w]);
        const double R = res * 15.890248582070704;                    // sqrt(15.5^2 + 3.5^2) cells
        const double b = (sqrt(mm) + 2.0 * R) * (1.0 + 1e-12);
        s_bound2 = b * b;
    }
    __syncthreads();
    const double bound2 = s_bound2;
    for (int j = tid; j < npoints; j += 256)
        if (dist2(cx, cy, px[j], py[j]) <= bound2) {
            const int k = atomicAdd(&s_cnt, 1);
            if (k < NEAREST_MAX_CAND) s_cand[k] = j;
        }
    __syncthreads();
    const int cnt = s_cnt;
    const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
    if (r >= rows || c >= cols) return;
    const double x = x0 + c * res, y = y0 + r * res;
    double best = 1.0e300;
    int bj = 0x7fffffff;
    if (cnt <= NEAREST_MAX_CAND) {
        for (int k = 0; k < cnt; ++k) {
            const int j = s_cand[k];
            const double d = dist2(x, y, px[j], py[j]);
            if (d < best || (d == best && j < bj)) { best = d; bj = j; }
        }
    } else {                                                          // too many survivors (very dense sites): scan all
        for (int j = 0; j < npoints; ++j) {
            const double d = dist2(x, y, px[j], py[j]);
            if (d < best) { best = d; bj = j; }
        }
    }
    const double e = east[bj], n = north[bj];
    const double two_pi = 6.283185307179586;
    const long long i = (long long)r * cols + c;
    wspeed[i] = (float)sqrt(e * e + n * n);                                              // simulator.py:787-788
    wdirn[i] = (float)(fmod(atan2(e, n) + two_pi, two_pi) * (180.0 / 3.141592653589793));   // :789-791
}

// ---- thermals -----------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned& o0, unsigned& o1, unsigned& o2, unsigned& o3) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

__global__ void __launch_bounds__(256) thermal_seeds_kernel(const float* __restrict__ aspect, int rows, int cols, float mean,
                                                            float sigma, unsigned long long seed, float* __restrict__ out) {
    const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
    if (r >= rows || c >= cols) return;
    const long long i = (long long)r * cols + c;
    const int bx = (int)(0.1 * cols), by = (int)(0.1 * rows);                   // layers.py:194-195
    float v = 0.0f;
    if (r >= by && r < rows - by && c >= bx && c < cols - bx) {
        const double wtfactor = 1000.0 + (fabs((double)aspect[i] - 180.0) / 180.0) * 2000.0;   // :199-200
        const unsigned outcomes = (unsigned)((int)wtfactor - 1);                // randint(1, int(wtfactor)): values 1..int-1
        unsigned w0, w1, w2, w3;
        philox4x32_10((unsigned)i, (unsigned)(i >> 32), 0u, 0x7468726du, (unsigned)seed, (unsigned)(seed >> 32), w0, w1, w2, w3);
        // one outcome in `outcomes` is a hit (the reference tests == 5): Lemire's multiply-shift, bias < 2^-20
        const unsigned draw = (unsigned)(((unsigned long long)w0 * outcomes) >> 32);
        if (draw == 4u) {
            // lognormal(mean, sigma) = exp(mean + sigma * N(0,1)), Box-Muller on two 32-bit words
            const float u1 = ((float)(w1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float u2 = ((float)(w2 >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
            v = expf(mean + sigma * z);
        }
    }
    out[i] = v;
}

// one separable pass of scipy.ndimage.gaussian_filter (mode='constant', cval=0): weights[k] for offsets -R..R
template <bool ALONG_ROWS>
__global__ void __launch_bounds__(256) blur_pass_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols,
                                                        int radius, const float* __restrict__ weights) {
    extern __shared__ float wsh[];
    for (int k = threadIdx.y * 32 + threadIdx.x; k <= 2 * radius; k += 256) wsh[k] = weights[k];
    __syncthreads();
    const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
    if (r >= rows || c >= cols) return;
    float acc = 0.0f;
    if (ALONG_ROWS) {        // axis 0
        const int lo = max(r - radius, 0), hi = min(r + radius, rows - 1);
        for (int y = lo; y <= hi; ++y) acc += wsh[y - r + radius] * __ldg(in + (long long)y * cols + c);
    } else {
        const int lo = max(c - radius, 0), hi = min(c + radius, cols - 1);
        const float* row = in + (long long)r * cols;
        for (int x = lo; x <= hi; ++x) acc += wsh[x - c + radius] * __ldg(row + x);
    }
    out[(long long)r * cols + c] = acc;
}

}  // namespace
}  // namespace ssrs

using namespace ssrs;

extern "C" int ssrs_interp_wind(const double* px, const double* py, const double* east, const double* north, int npoints,
                                const int32_t* triangles, int ntriangles, double x0, double y0, double resolution,
                                int rows, int cols, int32_t* owner_scratch, float* wspeed, float* wdirn, void* stream) {
    SSRS_REQUIRE(px && py && east && north && triangles && owner_scratch && wspeed && wdirn, "ssrs_interp_wind: NULL buffer");
    SSRS_REQUIRE(npoints >= 3 && ntriangles >= 1 && rows > 0 && cols > 0 && resolution > 0.0, "ssrs_interp_wind: bad sizes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SSRS_CUDA_TRY(cudaMemsetAsync(owner_scratch, 0x7f, sizeof(int32_t) * (size_t)rows * cols, st));
    const int blocks = ntriangles < sm_count() * 8 ? ntriangles : sm_count() * 8;
    rasterise_triangles_kernel<<<blocks, 256, 0, st>>>(px, py, triangles, ntriangles, x0, y0, resolution, rows, cols, owner_scratch);
    SSRS_CUDA_TRY(cudaGetLastError());
    dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 8));
    interp_wind_kernel<<<grid, dim3(32, 8), 0, st>>>(px, py, east, north, triangles, x0, y0, resolution, rows, cols,
                                                     owner_scratch, wspeed, wdirn);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int64_t ssrs_interp_wind_cubic_scratch_bytes(int npoints, int ntriangles) {
    if (npoints < 0 || ntriangles < 0) return 0;
    // gradients [2][npoints][2], ordinates [ntriangles][2][20], sweep counts [2] (int32)
    return (int64_t)sizeof(double) * (4 * (int64_t)npoints + 2 * (int64_t)ct::COEF_STRIDE * ntriangles + 1);
}

extern "C" int ssrs_interp_wind_cubic(const double* px, const double* py, const double* east, const double* north, int npoints,
                                      const int32_t* triangles, const int32_t* neighbors, int ntriangles,
                                      const int32_t* vertex_nb_indptr, const int32_t* vertex_nb_indices, double x0, double y0,
                                      double resolution, int rows, int cols, int32_t* owner_scratch, void* ct_scratch,
                                      float* wspeed, float* wdirn, void* stream) {
    SSRS_REQUIRE(px && py && east && north && triangles && neighbors && vertex_nb_indptr && vertex_nb_indices &&
                 owner_scratch && ct_scratch && wspeed && wdirn, "ssrs_interp_wind_cubic: NULL buffer");
    SSRS_REQUIRE(npoints >= 3 && ntriangles >= 1 && rows > 0 && cols > 0 && resolution > 0.0, "ssrs_interp_wind_cubic: bad sizes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* grad = static_cast<double*>(ct_scratch);
    double* coef = grad + 4 * (size_t)npoints;
    int* sweeps = reinterpret_cast<int*>(coef + 2 * (size_t)ct::COEF_STRIDE * ntriangles);
    ct_gradients_kernel<<<2, 32, 0, st>>>(px, py, east, north, npoints, vertex_nb_indptr, vertex_nb_indices, 400, 1e-6, grad,
                                          sweeps);                                  // griddata's maxiter and tol
    SSRS_CUDA_TRY(cudaGetLastError());
    ct_coefficients_kernel<<<(unsigned)cdiv(2 * ntriangles, 128), 128, 0, st>>>(px, py, east, north, npoints, triangles, neighbors,
                                                                      ntriangles, grad, coef);
    SSRS_CUDA_TRY(cudaGetLastError());
    SSRS_CUDA_TRY(cudaMemsetAsync(owner_scratch, 0x7f, sizeof(int32_t) * (size_t)rows * cols, st));
    const int blocks = ntriangles < sm_count() * 8 ? ntriangles : sm_count() * 8;
    rasterise_triangles_kernel<<<blocks, 256, 0, st>>>(px, py, triangles, ntriangles, x0, y0, resolution, rows, cols, owner_scratch);
    SSRS_CUDA_TRY(cudaGetLastError());
    dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 8));
    interp_wind_cubic_kernel<<<grid, dim3(32, 8), 0, st>>>(px, py, triangles, x0, y0, resolution, rows, cols, owner_scratch, coef,
                                                           wspeed, wdirn);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int ssrs_interp_wind_nearest(const double* px, const double* py, const double* east, const double* north, int npoints,
                                        double x0, double y0, double resolution, int rows, int cols, float* wspeed,
                                        float* wdirn, void* stream) {
    SSRS_REQUIRE(px && py && east && north && wspeed && wdirn, "ssrs_interp_wind_nearest: NULL buffer");
    SSRS_REQUIRE(npoints >= 1 && rows > 0 && cols > 0 && resolution > 0.0, "ssrs_interp_wind_nearest: bad sizes");
    dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 8));
    nearest_wind_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(px, py, east, north, npoints, x0, y0,
                                                                                 resolution, rows, cols, wspeed, wdirn);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int ssrs_thermal_seeds(const float* aspect, int rows, int cols, float thermal_intensity_scale, uint64_t seed,
                                  float* seeds, void* stream) {
    SSRS_REQUIRE(aspect && seeds, "ssrs_thermal_seeds: NULL buffer");
    SSRS_REQUIRE(rows > 0 && cols > 0, "ssrs_thermal_seeds: bad sizes");
    dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 8));
    thermal_seeds_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
        aspect, rows, cols, thermal_intensity_scale + 3.0f, 0.5f, seed, seeds);      // layers.py:203-204
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int ssrs_gaussian_blur(const float* in, float* out, float* tmp, int rows, int cols, float sigma, float truncate,
                                  float* weights_scratch, void* stream) {
    SSRS_REQUIRE(in && out && tmp && weights_scratch, "ssrs_gaussian_blur: NULL buffer");
    SSRS_REQUIRE(rows > 0 && cols > 0 && sigma > 0.0f && truncate > 0.0f, "ssrs_gaussian_blur: bad arguments");
    const int radius = (int)(truncate * sigma + 0.5f);                               // scipy _gaussian_kernel1d radius
    SSRS_REQUIRE(radius <= 2048, "ssrs_gaussian_blur: kernel radius above 2048 cells");
    float w[2 * 2048 + 1];
    double sum = 0.0;
    for (int k = -radius; k <= radius; ++k) { const double e = exp(-0.5 / ((double)sigma * sigma) * k * k); sum += e; }
    for (int k = -radius; k <= radius; ++k) w[k + radius] = (float)(exp(-0.5 / ((double)sigma * sigma) * k * k) / sum);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SSRS_CUDA_TRY(cudaMemcpyAsync(weights_scratch, w, sizeof(float) * (2 * radius + 1), cudaMemcpyHostToDevice, st));
    SSRS_CUDA_TRY(cudaStreamSynchronize(st));          // `w` lives on this stack frame
    dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 8));
    const size_t sh = sizeof(float) * (2 * radius + 1);
    blur_pass_kernel<true><<<grid, dim3(32, 8), sh, st>>>(in, tmp, rows, cols, radius, weights_scratch);     // axis 0 first, like scipy
    blur_pass_kernel<false><<<grid, dim3(32, 8), sh, st>>>(tmp, out, rows, cols, radius, weights_scratch);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}
