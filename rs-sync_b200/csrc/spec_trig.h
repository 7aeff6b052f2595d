// sin, cos and acos as FIXED EXPRESSION TREES over IEEE binary64 (+, -, x, /, sqrt, fma, rint),
// shared by the host and the device side of the gyro ingest.
//
// Why: the gyro integration (quat_from_aa, quat.cpp:5-17: sin and cos of half the rotation angle) and
// the variable-rate resampling (quat_slerp, quat.cpp:55-74: acos and three sines) are the only places
// of the path that call libm, and libm's results differ between glibc and CUDA by an ulp here and
// there -- enough to move Sync's amplified gradient.  With these three functions
// the device kernels, the host code and the CPU oracle produce the same bits, so the ingest can run
// on the GPU without un-pinning the oracle comparison (the same move as log1p_nonneg, device_math.cuh).
// Accuracy: <= 1 ulp from libm on the ranges the path uses (checked by the CPU test suite); the algorithms
// are the classic ones: Cody-Waite reduction by pi/2 in two fma steps and the minimax kernels of
// degree 13 / 14 for sin / cos on [-pi/4, pi/4]; the rational approximation of asin on [0, 1/2] with
// the half-angle identity for acos.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define RS_TRIG_HD __host__ __device__ __forceinline__
#else
#define RS_TRIG_HD inline
#endif

namespace rs {

namespace trig_detail {

RS_TRIG_HD double fma_(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return ::fma(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
RS_TRIG_HD double rint_(double x) {
#ifdef __CUDA_ARCH__
    return ::rint(x);
#else
    return std::nearbyint(x);  // round-to-nearest-even is the default rounding mode
#endif
}
RS_TRIG_HD double sqrt_(double x) {
#ifdef __CUDA_ARCH__
    return ::sqrt(x);
#else
    return std::sqrt(x);
#endif
}
RS_TRIG_HD double clear_low_word(double x) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(x), 0);
#else
    uint64_t u;
    std::memcpy(&u, &x, 8);
    u &= 0xffffffff00000000ULL;
    std::memcpy(&x, &u, 8);
    return x;
#endif
}

// sin on |r| <= pi/4: r + r^3 (S1 + r^2 (S2 + ... S6))
RS_TRIG_HD double sin_kernel(double r) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = r * r;
    double p = fma_(z, S6, S5);
    p = fma_(z, p, S4);
    p = fma_(z, p, S3);
    p = fma_(z, p, S2);
    p = fma_(z, p, S1);
    return fma_(z * r, p, r);
}
// cos on |r| <= pi/4: 1 - r^2/2 + r^4 (C1 + r^2 (C2 + ... C6))
RS_TRIG_HD double cos_kernel(double r) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = r * r;
    double p = fma_(z, C6, C5);
    p = fma_(z, p, C4);
    p = fma_(z, p, C3);
    p = fma_(z, p, C2);
    p = fma_(z, p, C1);
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;
    // (1 - w) - hz is the rounding error of w, exact
    return w + (((1.0 - w) - hz) + (z * z) * p);
}
// x = k pi/2 + r, |r| <= pi/4 (+ rounding); returns k mod 4.  |x| < 2^30 (larger arguments: NaN)
RS_TRIG_HD int reduce_pio2(double x, double& r) {
    const double kTwoOverPi = 6.36619772367581382433e-01;
    const double kPio2Hi = 1.57079632679489655800e+00, kPio2Lo = 6.12323399573676603587e-17;
    const double kd = rint_(x * kTwoOverPi);
    r = fma_(-kd, kPio2Lo, fma_(-kd, kPio2Hi, x));
    return (int)((long long)kd & 3);
}

}  // namespace trig_detail

RS_TRIG_HD double spec_sin(double x) {
    using namespace trig_detail;
    if (!(x > -1073741824.0 && x < 1073741824.0)) return (x - x) / (x - x);  // NaN for inf, NaN and huge arguments
    double r;
    const int q = reduce_pio2(x, r);
    const double s = sin_kernel(r), c = cos_kernel(r);
    return (q == 0) ? s : (q == 1) ? c : (q == 2) ? -s : -c;
}
RS_TRIG_HD double spec_cos(double x) {
    using namespace trig_detail;
    if (!(x > -1073741824.0 && x < 1073741824.0)) return (x - x) / (x - x);
    double r;
    const int q = reduce_pio2(x, r);
    const double s = sin_kernel(r), c = cos_kernel(r);
    return (q == 0) ? c : (q == 1) ? -s : (q == 2) ? -c : s;
}
// acos on [-1, 1]; NaN outside (quat_slerp's dot product of two nearly equal unit quaternions can
// exceed 1 by an ulp: the reference's std::acos returns NaN there too, which its panic then catches,
// core_private.cpp:180)
RS_TRIG_HD double spec_acos(double x) {
    using namespace trig_detail;
    const double kPio2Hi = 1.57079632679489655800e+00, kPio2Lo = 6.12323399573676603587e-17;
    const double kPi = 3.14159265358979311600e+00;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    const double ax = x < 0 ? -x : x;
    if (!(ax <= 1.0)) return (x - x) / (x - x);  // |x| > 1 or NaN
    if (ax == 1.0) return x > 0 ? 0.0 : kPi + 2.0 * kPio2Lo;
    // R(z) = z P(z) / Q(z) ~ (asin(sqrt z) - sqrt z) / sqrt z
    if (ax < 0.5) {
        const double z = x * x;
        const double p = z * fma_(z, fma_(z, fma_(z, fma_(z, fma_(z, pS5, pS4), pS3), pS2), pS1), pS0);
        const double q = fma_(z, fma_(z, fma_(z, fma_(z, qS4, qS3), qS2), qS1), 1.0);
        const double r = p / q;
        return kPio2Hi - (x - (kPio2Lo - x * r));
    }
    const double z = (1.0 - ax) * 0.5;
    const double p = z * fma_(z, fma_(z, fma_(z, fma_(z, fma_(z, pS5, pS4), pS3), pS2), pS1), pS0);
    const double q = fma_(z, fma_(z, fma_(z, fma_(z, qS4, qS3), qS2), qS1), 1.0);
    const double s = sqrt_(z);
    const double r = p / q;
    if (x < 0) {
        const double w = r * s - kPio2Lo;
        return kPi - 2.0 * (s + w);
    }
    const double df = clear_low_word(s);
    const double c = (z - df * df) / (s + df);
    const double w = r * s + c;
    return 2.0 * (df + w);
}

}  // namespace rs
