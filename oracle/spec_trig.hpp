// TEST INFRASTRUCTURE (CPU oracle).  sin, cos, acos of the arithmetic contract:
// fixed expression trees over IEEE binary64, so that the oracle, the engine's host code and its
// device kernels produce the same bits in the gyro ingest (quat_from_aa, quat.cpp:5-17; quat_slerp,
// quat.cpp:55-74, where the reference calls libm).  Plain host C++; compile with -ffp-contract=off.
// Cody-Waite reduction by pi/2 (two fma steps), degree-13 / 14 minimax kernels on [-pi/4, pi/4];
// rational asin approximation on [0, 1/2] with the half-angle identity for acos.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

inline double trig_sin_kernel(double r) {
    const double S[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                         2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
    const double z = r * r;
    double p = S[5];
    for (int i = 4; i >= 0; --i) p = std::fma(z, p, S[i]);
    return std::fma(z * r, p, r);
}
inline double trig_cos_kernel(double r) {
    const double C[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                         -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
    const double z = r * r;
    double p = C[5];
    for (int i = 4; i >= 0; --i) p = std::fma(z, p, C[i]);
    const double hz = 0.5 * z, w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * z) * p);
}
inline int trig_reduce(double x, double& r) {
    const double kd = std::nearbyint(x * 6.36619772367581382433e-01);
    r = std::fma(-kd, 6.12323399573676603587e-17, std::fma(-kd, 1.57079632679489655800e+00, x));
    return (int)((long long)kd & 3);
}
inline double spec_sin(double x) {
    if (!(x > -1073741824.0 && x < 1073741824.0)) return (x - x) / (x - x);
    double r;
    const int q = trig_reduce(x, r);
    const double s = trig_sin_kernel(r), c = trig_cos_kernel(r);
    return q == 0 ? s : q == 1 ? c : q == 2 ? -s : -c;
}
inline double spec_cos(double x) {
    if (!(x > -1073741824.0 && x < 1073741824.0)) return (x - x) / (x - x);
    double r;
    const int q = trig_reduce(x, r);
    const double s = trig_sin_kernel(r), c = trig_cos_kernel(r);
    return q == 0 ? c : q == 1 ? -s : q == 2 ? -c : s;
}
inline double spec_acos(double x) {
    const double hi = 1.57079632679489655800e+00, lo = 6.12323399573676603587e-17, pi = 3.14159265358979311600e+00;
    const double P[6] = {1.66666666666666657415e-01, -3.25565818622400915405e-01, 2.01212532134862925881e-01,
                         -4.00555345006794114027e-02, 7.91534994289814532176e-04, 3.47933107596021167570e-05};
    const double Q[5] = {1.0, -2.40339491173441421878e+00, 2.02094576023350569471e+00, -6.88283971605453293030e-01,
                         7.70381505559019352791e-02};
    auto R = [&](double z) {
        double p = P[5], q = Q[4];
        for (int i = 4; i >= 0; --i) p = std::fma(z, p, P[i]);
        for (int i = 3; i >= 0; --i) q = std::fma(z, q, Q[i]);
        return (z * p) / q;
    };
    const double ax = std::fabs(x);
    if (!(ax <= 1.0)) return (x - x) / (x - x);
    if (ax == 1.0) return x > 0 ? 0.0 : pi + 2.0 * lo;
    if (ax < 0.5) return hi - (x - (lo - x * R(x * x)));
    const double z = (1.0 - ax) * 0.5, s = std::sqrt(z), r = R(z);
    if (x < 0) return pi - 2.0 * (s + (r * s - lo));
    uint64_t u;
    double df = s;
    std::memcpy(&u, &df, 8);
    u &= 0xffffffff00000000ULL;
    std::memcpy(&df, &u, 8);
    const double c = (z - df * df) / (s + df);
    return 2.0 * (df + (r * s + c));
}

}  // namespace orc
