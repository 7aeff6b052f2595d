// ORACLE — TEST INFRASTRUCTURE ONLY.
// "Strict" arithmetic: the reference's own expression order, operation by operation, with no
// FMA, plain left-to-right sums and libm's log1p — i.e. what the unmodified reference sources
// compute when built against the plain-loop Armadillo stand-in of oracle/shim.  The oracle can be
// switched to this mode (orc_set_strict) so that its CONTROL FLOW (RANSAC decisions, L-BFGS,
// backtracking, momentum loop, convergence tests) can be compared with the reference bit for bit;
// the default "spec" arithmetic of oracle_math.hpp (explicit FMA, double-double sums, own log1p)
// is a different rounding of the same formulas, chosen so a GPU can reproduce it exactly.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace orc_strict {

// spline::operator() (minispline.cpp:48-55), one component
static inline double spline1(const double* rec, long n, int c, double x) {
    double fl = std::floor(x);
    double idxf = std::max(std::min(fl, (double)n), 0.);
    size_t idx = (size_t)idxf;
    double h = x - idx;
    const double* r0 = rec;
    const double* rl = rec + (n - 1) * 16;
    if (x < idx) return (r0[8 + c] * h + r0[4 + c]) * h + r0[c];
    if (x > n - 1) return (rl[8 + c] * h + rl[4 + c]) * h + rl[c];
    const double* p = rec + idx * 16;
    return ((p[12 + c] * h + p[8 + c]) * h + p[4 + c]) * h + p[c];
}

// quat_prod (quat.cpp:33-38)
static inline void qprod(const double p[4], const double q[4], double r[4]) {
    r[0] = p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3];
    r[1] = p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2];
    r[2] = p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1];
    r[3] = p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0];
}

static inline double norm_seq(const double* v, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i] * v[i];
    return std::sqrt(s);
}

// one row of opt_compute_problem (core_private.cpp:19-28)
static inline void problem_row(const double* rec, long n, double q0, double sr, double delay,
                               double ts_a, double ts_b, const double ra[3], const double rb[3],
                               double row[3]) {
    const double xa = (ts_a - q0 + delay) * sr;
    const double xb = (ts_b - q0 + delay) * sr;
    double a[4], b[4];
    for (int c = 0; c < 4; ++c) { a[c] = spline1(rec, n, c, xa); b[c] = spline1(rec, n, c, xb); }
    const double na = norm_seq(a, 4), nb = norm_seq(b, 4);
    for (int c = 0; c < 4; ++c) { a[c] = a[c] / na; b[c] = b[c] / nb; }  // arma::normalise
    // quat_rotate_point(quat_conj(a), p) = conj(a) (x) ((0,p) (x) a)   (quat.cpp:45-47)
    double ca[4] = {a[0], -a[1], -a[2], -a[3]}, cb[4] = {b[0], -b[1], -b[2], -b[3]};
    double pa[4] = {0, ra[0], ra[1], ra[2]}, pb[4] = {0, rb[0], rb[1], rb[2]};
    double t[4], ar[4], br[4];
    qprod(pa, a, t); qprod(ca, t, ar);
    qprod(pb, b, t); qprod(cb, t, br);
    row[0] = ar[2] * br[3] - ar[3] * br[2];
    row[1] = ar[3] * br[1] - ar[1] * br[3];
    row[2] = ar[1] * br[2] - ar[2] * br[1];
}

static inline double dot3(const double* a, const double* b) {
    double s = 0.0;
    s += a[0] * b[0];
    s += a[1] * b[1];
    s += a[2] * b[2];
    return s;
}

// safe_normalize (inline_utils.hpp:5-11): m / norm
static inline void safe_normalize3(const double v[3], double out[3]) {
    double nrm = norm_seq(v, 3);
    if (nrm < 1e-12) { out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; return; }
    out[0] = v[0] / nrm; out[1] = v[1] / nrm; out[2] = v[2] / nrm;
}

static inline void cross3(const double* a, const double* b, double* r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

// arma::norm(P * M)
static inline double norm_PM(const double* P, int n, const double m[3]) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        double pm = dot3(P + 3 * i, m);
        s += pm * pm;
    }
    return std::sqrt(s);
}

// FrameState::Loss 3-arg (core_private.cpp:117-123)
static inline double loss3_P(const double* P, int n, const double m[3], double k) {
    const double scale = k / norm_seq(m, 3);
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        double r = dot3(P + 3 * i, m) * scale;
        acc += std::log1p(r * r);
    }
    return acc;
}

// FrameState::Loss 5-arg (core_private.cpp:99-114), the literal forward-mode chain with its
// diagonal matrix products multiplied out (adding exact zeros does not change the values)
static inline double loss5_P(const double* P, int n, const double m[3], double k, double grad[3]) {
    double v4 = 0.0;
    for (int c = 0; c < 3; ++c) v4 += m[c] * m[c];       // sum_jac(sqr_jac(m))
    const double y = k * k;
    const double den = v4 / y;                             // div_jac(v4, k*k)
    const double j5 = 1.0 / y;
    double L = 0.0, g[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < n; ++i) {
        const double* p = P + 3 * i;
        const double v1 = dot3(p, m);
        const double v2 = v1 * v1;
        const double u = v2 / den;                         // div_jac(v2, v5)
        L += std::log1p(u);
        const double j7 = 1. / (1. + u);
        const double a = (1.0 / den) * (2. * v1);         // j6a * j2
        const double b = (-v2 / (den * den)) * j5;        // j6b * j5 (* j4 = 1)
        for (int c = 0; c < 3; ++c) {
            const double t = a * p[c] + b * (2. * m[c]);  // (j6a j2 j1 + j6b j5 j4 j3)
            g[c] += (1.0 * j7) * t;                       // j8 * j7 * (...)
        }
    }
    grad[0] = g[0]; grad[1] = g[1]; grad[2] = g[2];
    return L;
}

}  // namespace orc_strict
