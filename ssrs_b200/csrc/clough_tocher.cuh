// Clough-Tocher piecewise-cubic interpolation on a Delaunay triangulation — the arithmetic behind
// scipy.interpolate.griddata(method='cubic') in two dimensions, which is what the reference calls when
// Config.wtk_interp_type = 'cubic' (ssrs/config.py:60, ssrs/simulator.py:765-776).  scipy is a third-party
// dependency of the reference (setup.py lists it unpinned; the image has 1.18.1); its published algorithm
// (scipy/interpolate/interpnd.pyx: CloughTocher2DInterpolator, Alfeld 1984 / Farin 1986 with Nielson's and Renka's
// global curvature-minimising gradient estimate) is restated here in three pieces:
//   ct_estimate_gradients   the Gauss-Seidel iteration of estimate_gradients_2d_global (tol 1e-6, <= 400 sweeps)
//   ct_coefficients         the 19 Bezier ordinates of one triangle's three micro-triangles
//   ct_evaluate             the cubic at barycentric coordinates b
// Everything is `__host__ __device__` so that tests/ctemu.py can compile the same lines with g++ and compare them with
// scipy on the CPU (test infrastructure only; the product library runs them in kernels, csrc/wind.cu).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SSRS_CT_HD __host__ __device__ __forceinline__
#else
#define SSRS_CT_HD inline
#endif

namespace ssrs {
namespace ct {

constexpr int N_COEF = 19;
constexpr int COEF_STRIDE = 20;      // padded: a triangle's ordinates start on a 32-byte boundary

// One vertex update of the Gauss-Seidel sweep: the 2 x 2 normal equations of the edge-curvature functional around
// vertex i, accumulated over its neighbours.  `W` provides the lanes that share the neighbour loop: lane(), lanes(),
// sum(v) (same total on every lane), sync().  One lane = scipy's summation order.
template <class W>
SSRS_CT_HD double gs_update_vertex(W& w, const double* px, const double* py, const double* f, double* grad, int i,
                                   const int* nb_indptr, const int* nb_indices) {
    double q0 = 0.0, q1 = 0.0, q3 = 0.0, s0 = 0.0, s1 = 0.0;
    const double xi = px[i], yi = py[i], f1 = f[i];
    for (int j = nb_indptr[i] + w.lane(); j < nb_indptr[i + 1]; j += w.lanes()) {
        const int p2 = nb_indices[j];
        const double ex = px[p2] - xi, ey = py[p2] - yi;
        const double L = sqrt(ex * ex + ey * ey);
        const double L3 = L * L * L;
        const double f2 = f[p2];
        const double df2 = -ex * grad[2 * p2] - ey * grad[2 * p2 + 1];      // neighbour's gradient along the edge
        q0 += 4 * ex * ex / L3;
        q1 += 4 * ex * ey / L3;
        q3 += 4 * ey * ey / L3;
        s0 += (6 * (f1 - f2) - 2 * df2) * ex / L3;
        s1 += (6 * (f1 - f2) - 2 * df2) * ey / L3;
    }
    q0 = w.sum(q0); q1 = w.sum(q1); q3 = w.sum(q3); s0 = w.sum(s0); s1 = w.sum(s1);
    const double q2 = q1;
    const double det = q0 * q3 - q1 * q2;
    const double r0 = (q3 * s0 - q1 * s1) / det;
    const double r1 = (-q2 * s0 + q0 * s1) / det;
    double change = fmax(fabs(grad[2 * i] + r0), fabs(grad[2 * i + 1] + r1));
    w.sync();                                     // every lane has read the old gradient
    if (w.lane() == 0) { grad[2 * i] = -r0; grad[2 * i + 1] = -r1; }
    w.sync();                                     // ... and sees the new one: the sweep is sequential over vertices
    change /= fmax(1.0, fmax(fabs(r0), fabs(r1)));
    return change;
}

// estimate_gradients_2d_global: sweeps until the largest relative change of a sweep falls below tol.  Returns the
// number of sweeps, 0 if maxiter sweeps did not converge (scipy then warns and uses the last iterate; so do we).
template <class W>
SSRS_CT_HD int estimate_gradients(W& w, const double* px, const double* py, const double* f, int npoints,
                                  const int* nb_indptr, const int* nb_indices, int maxiter, double tol, double* grad) {
    for (int i = w.lane(); i < 2 * npoints; i += w.lanes()) grad[i] = 0.0;
    w.sync();
    for (int it = 0; it < maxiter; ++it) {
        double err = 0.0;
        for (int i = 0; i < npoints; ++i) err = fmax(err, gs_update_vertex(w, px, py, f, grad, i, nb_indptr, nb_indices));
        if (err < tol) return it + 1;
    }
    return 0;
}

// Bezier ordinates of triangle (p0, p1, p2) with values f[3] and gradients df[3][2] at its vertices.  g[k] is the
// edge parameter of the side opposite vertex k: the direction in which the cross-boundary derivative is made linear
// points at the neighbouring triangle's centroid (affine invariant; scipy's choice), given as that centroid's
// barycentric coordinates c[k][3] in THIS triangle; has_nb[k] = 0 on the hull (g = -1/2).
SSRS_CT_HD void coefficients(const double p[3][2], const double f[3], const double df[3][2], const int has_nb[3],
                             const double c[3][3], double* out) {
    const double e12x = p[1][0] - p[0][0], e12y = p[1][1] - p[0][1];
    const double e23x = p[2][0] - p[1][0], e23y = p[2][1] - p[1][1];
    const double e31x = p[0][0] - p[2][0], e31y = p[0][1] - p[2][1];
    const double f1 = f[0], f2 = f[1], f3 = f[2];
    const double df12 = +(df[0][0] * e12x + df[0][1] * e12y);
    const double df21 = -(df[1][0] * e12x + df[1][1] * e12y);
    const double df23 = +(df[1][0] * e23x + df[1][1] * e23y);
    const double df32 = -(df[2][0] * e23x + df[2][1] * e23y);
    const double df31 = +(df[2][0] * e31x + df[2][1] * e31y);
    const double df13 = -(df[0][0] * e31x + df[0][1] * e31y);
    const double c3000 = f1, c2100 = (df12 + 3 * c3000) / 3, c2010 = (df13 + 3 * c3000) / 3;
    const double c0300 = f2, c1200 = (df21 + 3 * c0300) / 3, c0210 = (df23 + 3 * c0300) / 3;
    const double c0030 = f3, c1020 = (df31 + 3 * c0030) / 3, c0120 = (df32 + 3 * c0030) / 3;
    const double c2001 = (c2100 + c2010 + c3000) / 3;
    const double c0201 = (c1200 + c0300 + c0210) / 3;
    const double c0021 = (c1020 + c0120 + c0030) / 3;
    double g[3];
    g[0] = has_nb[0] ? (2 * c[0][2] + c[0][1] - 1) / (2 - 3 * c[0][2] - 3 * c[0][1]) : -0.5;
    g[1] = has_nb[1] ? (2 * c[1][0] + c[1][2] - 1) / (2 - 3 * c[1][0] - 3 * c[1][2]) : -0.5;
    g[2] = has_nb[2] ? (2 * c[2][1] + c[2][0] - 1) / (2 - 3 * c[2][1] - 3 * c[2][0]) : -0.5;
    const double c0111 = (g[0] * (-c0300 + 3 * c0210 - 3 * c0120 + c0030) + (-c0300 + 2 * c0210 - c0120 + c0021 + c0201)) / 2;
    const double c1011 = (g[1] * (-c0030 + 3 * c1020 - 3 * c2010 + c3000) + (-c0030 + 2 * c1020 - c2010 + c2001 + c0021)) / 2;
    const double c1101 = (g[2] * (-c3000 + 3 * c2100 - 3 * c1200 + c0300) + (-c3000 + 2 * c2100 - c1200 + c2001 + c0201)) / 2;
    const double c1002 = (c1101 + c1011 + c2001) / 3;
    const double c0102 = (c1101 + c0111 + c0201) / 3;
    const double c0012 = (c1011 + c0111 + c0021) / 3;
    const double c0003 = (c1002 + c0102 + c0012) / 3;
    out[0] = c3000; out[1] = c2100; out[2] = c2010; out[3] = c2001; out[4] = c1200; out[5] = c1101; out[6] = c1020;
    out[7] = c1011; out[8] = c1002; out[9] = c0300; out[10] = c0210; out[11] = c0201; out[12] = c0120; out[13] = c0111;
    out[14] = c0102; out[15] = c0030; out[16] = c0021; out[17] = c0012; out[18] = c0003;
}

// The cubic at barycentric coordinates (b0, b1, b2) of the macro-triangle: the smallest coordinate names the
// micro-triangle, and in the extended coordinates (b - min, 3 min) one of the four is zero.
SSRS_CT_HD double evaluate(const double* c, double b0, double b1, double b2) {
    const double m = fmin(b0, fmin(b1, b2));
    const double a1 = b0 - m, a2 = b1 - m, a3 = b2 - m, a4 = 3 * m;
    return a1 * a1 * a1 * c[0] + 3 * a1 * a1 * a2 * c[1] + 3 * a1 * a1 * a3 * c[2] + 3 * a1 * a1 * a4 * c[3]
         + 3 * a1 * a2 * a2 * c[4] + 6 * a1 * a2 * a4 * c[5] + 3 * a1 * a3 * a3 * c[6] + 6 * a1 * a3 * a4 * c[7]
         + 3 * a1 * a4 * a4 * c[8] + a2 * a2 * a2 * c[9] + 3 * a2 * a2 * a3 * c[10] + 3 * a2 * a2 * a4 * c[11]
         + 3 * a2 * a3 * a3 * c[12] + 6 * a2 * a3 * a4 * c[13] + 3 * a2 * a4 * a4 * c[14] + a3 * a3 * a3 * c[15]
         + 3 * a3 * a3 * a4 * c[16] + 3 * a3 * a4 * a4 * c[17] + a4 * a4 * a4 * c[18];
}

}  // namespace ct
}  // namespace ssrs
