"""TEST INFRASTRUCTURE ONLY — compiles ssrs_b200/csrc/clough_tocher.cuh (the `__host__ __device__` arithmetic of the
'cubic' wind interpolation) with g++ into tests/_build/ so that `-m "not gpu"` tests can compare the very lines the
kernels run with scipy's CloughTocher2DInterpolator on the CPU.  The product package never loads this library."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "ssrs_b200", "csrc", "clough_tocher.cuh")
OUT = os.path.join(ROOT, "tests", "_build", "libssrs_ct_emu.so")

SRC = r"""
#include "clough_tocher.cuh"
using namespace ssrs;
namespace {
struct OneLane {                       // scipy's own order: one lane walks all neighbours
    int lane() const { return 0; }
    int lanes() const { return 1; }
    double sum(double v) const { return v; }
    void sync() const {}
};
// the same 2 x 2 inverse csrc/wind.cu's load_tri/barycentric use
void bary(const double* px, const double* py, const int* t, double x, double y, double b[3]) {
    const double x2 = px[t[2]], y2 = py[t[2]];
    const double m00 = px[t[0]] - x2, m01 = px[t[1]] - x2, m10 = py[t[0]] - y2, m11 = py[t[1]] - y2;
    const double inv = 1.0 / (m00 * m11 - m01 * m10);
    const double dx = x - x2, dy = y - y2;
    b[0] = (m11 * inv) * dx + (-m01 * inv) * dy;
    b[1] = (-m10 * inv) * dx + (m00 * inv) * dy;
    b[2] = 1.0 - b[0] - b[1];
}
}
extern "C" int emu_ct_gradients(const double* px, const double* py, const double* f, int n, const int* indptr,
                                const int* indices, int maxiter, double tol, double* grad) {
    OneLane w;
    return ct::estimate_gradients(w, px, py, f, n, indptr, indices, maxiter, tol, grad);
}
extern "C" void emu_ct_interp(const double* px, const double* py, const double* f, const double* grad, const int* tri,
                              const int* nbr, const double* xq, const double* yq, const int* simplex, int nq, double* out) {
    for (int q = 0; q < nq; ++q) {
        const int t = simplex[q];
        if (t < 0) { out[q] = NAN; continue; }
        const int* v = tri + 3 * t;
        double p[3][2], fv[3], df[3][2], c[3][3] = {{0}};
        int has_nb[3];
        for (int j = 0; j < 3; ++j) {
            p[j][0] = px[v[j]]; p[j][1] = py[v[j]]; fv[j] = f[v[j]];
            df[j][0] = grad[2 * v[j]]; df[j][1] = grad[2 * v[j] + 1];
            const int nb = nbr[3 * t + j];
            has_nb[j] = nb >= 0;
            if (nb >= 0) {
                const int* u = tri + 3 * nb;
                bary(px, py, v, (px[u[0]] + px[u[1]] + px[u[2]]) / 3, (py[u[0]] + py[u[1]] + py[u[2]]) / 3, c[j]);
            }
        }
        double coef[ct::COEF_STRIDE];
        ct::coefficients(p, fv, df, has_nb, c, coef);
        double b[3];
        bary(px, py, v, xq[q], yq[q], b);
        out[q] = ct::evaluate(coef, b[0], b[1], b[2]);
    }
}
"""

_lib = None


def lib():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        if not os.path.exists(OUT) or max(os.path.getmtime(HDR), os.path.getmtime(__file__)) > os.path.getmtime(OUT):
            tmp = f"{OUT}.{os.getpid()}.tmp"
            subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.dirname(HDR), "-x", "c++", "-",
                            "-o", tmp], input=SRC.encode(), check=True)
            os.replace(tmp, OUT)
        _lib = C.CDLL(OUT)
        _lib.emu_ct_gradients.restype = C.c_int
        _lib.emu_ct_interp.restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gradients(points, values, indptr, indices, maxiter=400, tol=1e-6):
    px = np.ascontiguousarray(points[:, 0], dtype=np.float64)
    py = np.ascontiguousarray(points[:, 1], dtype=np.float64)
    f = np.ascontiguousarray(values, dtype=np.float64)
    ip = np.ascontiguousarray(indptr, dtype=np.int32)
    ix = np.ascontiguousarray(indices, dtype=np.int32)
    grad = np.empty((len(px), 2), dtype=np.float64)
    sweeps = lib().emu_ct_gradients(_p(px), _p(py), _p(f), C.c_int(len(px)), _p(ip), _p(ix), C.c_int(maxiter), C.c_double(tol),
                                    _p(grad))
    return grad, sweeps


def interpolate(points, values, grad, triangles, neighbors, xq, yq, simplex):
    px = np.ascontiguousarray(points[:, 0], dtype=np.float64)
    py = np.ascontiguousarray(points[:, 1], dtype=np.float64)
    f = np.ascontiguousarray(values, dtype=np.float64)
    g = np.ascontiguousarray(grad, dtype=np.float64)
    tri = np.ascontiguousarray(triangles, dtype=np.int32)
    nbr = np.ascontiguousarray(neighbors, dtype=np.int32)
    xq = np.ascontiguousarray(xq, dtype=np.float64)
    yq = np.ascontiguousarray(yq, dtype=np.float64)
    sx = np.ascontiguousarray(simplex, dtype=np.int32)
    out = np.empty(len(xq), dtype=np.float64)
    lib().emu_ct_interp(_p(px), _p(py), _p(f), _p(g), _p(tri), _p(nbr), _p(xq), _p(yq), _p(sx), C.c_int(len(xq)), _p(out))
    return out
