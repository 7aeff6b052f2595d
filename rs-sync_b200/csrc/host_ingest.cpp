// Host-side gyro ingest of the engine: variable-rate resampling and the natural-cubic-spline
// coefficient build.  Both are short sequential recurrences over the gyro track (O(n), run once
// per SetGyroQuaternions call, ~1 ms for a minute of 1 kHz data), and their results feed the
// finite-difference delay gradient of Sync, which amplifies 1-ulp differences by 1/(2h) = 5e5 —
// so they are evaluated in one fixed order on the host and only the finished 128-byte spline
// records travel to the GPU.  Compile with -ffp-contract=off.
#include "host_ingest.h"

#include <algorithm>
#include <cmath>
#include <memory>
#include <mutex>
#include <thread>

namespace rs {

namespace {

// ndspline::make (ndspline.cpp:13-19) = spline::set_points (minispline.cpp:3-46) per component:
// unit-spaced knots; interior rows [1/3, 4/3, 1/3] with rhs y[i+1] - 2 y[i] + y[i-1], natural
// boundary rows [., 2, 0] / [0, 2, .] with rhs 0; eliminate downwards (:22-26), then upwards
// (:28-32), divide (:34), then d and b per interval (:38-44).
//
// The matrix entries and both elimination factors do not depend on the data, only on n, so they are
// computed once for the four components -- by the same operations on the same values as the
// reference's per-component loop, hence the same bits.  The interior rows start out identical, so
// the elimination state reaches a floating-point fixed point after a few dozen rows; from there on
// every row repeats it exactly and the (division-bound) recurrence is not re-evaluated.  What
// remains sequential per component is rhs[i+1] -= rhs[i] * f[i] and its mirror image, which run
// for the four components side by side.
struct Elimination {
    std::vector<double> f_down;  // f_down[i]: factor with which row i is subtracted from row i + 1
    std::vector<double> f_up;    // f_up[i]:   factor with which row i is subtracted from row i - 1
    std::vector<double> diag;    // diagonal after both sweeps
};

Elimination eliminate(size_t n) {
    Elimination e;
    e.f_down.assign(n, 0.0);
    e.f_up.assign(n, 0.0);
    e.diag.assign(n, 2.0 / 3.0 * 2.0);
    std::vector<double> sub(n, 1.0 / 3.0), sup(n, 1.0 / 3.0);
    std::vector<double>& diag = e.diag;
    sub[0] = 0.0; diag[0] = 2.0; sup[0] = 0.0;
    sub[n - 1] = 0.0; diag[n - 1] = 2.0; sup[n - 1] = 0.0;
    // downwards, minispline.cpp:22-26: row i clears the sub-diagonal of row i + 1, for i + 2 < n
    size_t steady = n;  // rows steady .. n-2 leave this sweep in the same state
    for (size_t i = 0; i + 2 < n; ++i) {
        const double f = 1. / diag[i] * sub[i + 1];
        sub[i + 1] -= diag[i] * f;
        diag[i + 1] -= sup[i] * f;
        e.f_down[i] = f;
        // row i + 1 came out exactly like row i (both interior): the rows below start out like
        // row i + 1 did, so each repeats this step with the same operands
        if (i >= 1 && sub[i + 1] == sub[i] && diag[i + 1] == diag[i] && f == e.f_down[i - 1]) {
            for (size_t j = i + 1; j + 2 < n; ++j) {
                sub[j + 1] = sub[i];
                diag[j + 1] = diag[i];
                e.f_down[j] = f;
            }
            steady = i;
            break;
        }
    }
    // upwards, :28-32: row i clears the super-diagonal of row i - 1, for i > 1
    for (size_t i = n - 1; i > 1; --i) {
        const double f = 1. / diag[i] * sup[i - 1];
        diag[i - 1] -= sub[i] * f;
        sup[i - 1] -= diag[i] * f;
        e.f_up[i] = f;
        // rows i - 1 and i entered this sweep in the same state (both in the steady range, as is
        // row i + 1, whose sub-diagonal entered step i + 1) and row i - 1 left it with row i's
        // diagonal: step i - 1 therefore sees the operands of step i, and so on up to `steady`
        if (i + 1 <= n - 2 && i - 1 > steady && diag[i - 1] == diag[i]) {
            for (size_t j = i - 1; j > steady; --j) {  // step j writes row j - 1 >= steady
                diag[j - 1] = diag[i - 1];
                sup[j - 1] = sup[i - 1];
                e.f_up[j] = f;
            }
            i = steady + 1;  // the loop continues with step `steady`
        }
    }
    return e;
}

}  // namespace

void build_spline_system(const double* quats, size_t n, double* rhs, double* diag) {
    // the elimination only depends on n: the orientation search builds 48 splines of one length
    static std::mutex mu;
    static std::shared_ptr<const Elimination> cached;
    static size_t cached_n = 0;
    std::shared_ptr<const Elimination> ep;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cached && cached_n == n) ep = cached;
    }
    if (!ep) {
        ep = std::make_shared<const Elimination>(eliminate(n));
        std::lock_guard<std::mutex> lk(mu);
        cached = ep;
        cached_n = n;
    }
    const Elimination& e = *ep;
    std::copy(e.diag.begin(), e.diag.end(), diag);
    // right-hand sides of the four components, interleaved like the input; the downward sweep (:25)
    // consumes each row right after it is formed
    for (int c = 0; c < 4; ++c) rhs[c] = rhs[4 * (n - 1) + c] = 0.0;
    for (size_t i = 1; i + 1 < n; ++i) {
        const double f = e.f_down[i - 1];
        for (int c = 0; c < 4; ++c) {
            const double r = (quats[4 * (i + 1) + c] - 2 * quats[4 * i + c]) + quats[4 * (i - 1) + c];
            rhs[4 * i + c] = r - rhs[4 * (i - 1) + c] * f;
        }
    }
    for (size_t i = n - 1; i > 1; --i) {  // :31
        const double f = e.f_up[i];
        for (int c = 0; c < 4; ++c) rhs[4 * (i - 1) + c] -= rhs[4 * i + c] * f;
    }
}

namespace {

// quat_slerp, quat.cpp:55-74
void slerp4(const double* p, const double* q_in, double t, double* out) {
    double q[4] = {q_in[0], q_in[1], q_in[2], q_in[3]};
    double cosang = ((p[0] * q[0] + p[1] * q[1]) + p[2] * q[2]) + p[3] * q[3];
    if (cosang < 0) {
        q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3];
        cosang = ((p[0] * q[0] + p[1] * q[1]) + p[2] * q[2]) + p[3] * q[3];
    }
    const double ang = std::acos(cosang);
    double wp, wq;
    if (ang > 1e-9) {
        const double s = std::sin(ang);
        wp = std::sin((1 - t) * ang) / s;
        wq = std::sin(t * ang) / s;
    } else {
        wp = 1 - t;
        wq = t;
    }
    for (int c = 0; c < 4; ++c) out[c] = wp * p[c] + wq * q[c];
}

}  // namespace

bool integrate_gyro(const double* ts, const double* gyro, size_t count, const char* orient,
                    double* out) {
    int src[3] = {0, 1, 2};
    double sgn[3] = {1.0, 1.0, 1.0};
    if (orient) {
        for (int i = 0; i < 3; ++i) {
            const char ch = orient[i];
            const char lo = (char)(ch | 0x20);
            if (lo < 'x' || lo > 'z') return false;
            src[i] = lo - 'x';
            sgn[i] = (ch == lo) ? -1.0 : 1.0;
        }
        if (orient[3] != 0) return false;
    }
    if (count == 0) return true;
    double q[4] = {1.0, 0.0, 0.0, 0.0};
    out[0] = 1.0; out[1] = 0.0; out[2] = 0.0; out[3] = 0.0;
    for (size_t i = 1; i < count; ++i) {
        const double dt = ts[i] - ts[i - 1];
        const double a0 = sgn[0] * gyro[3 * i + src[0]] * dt, a1 = sgn[1] * gyro[3 * i + src[1]] * dt,
                     a2 = sgn[2] * gyro[3 * i + src[2]] * dt;
        // quat_from_aa, quat.cpp:5-17
        const double th2 = (a0 * a0 + a1 * a1) + a2 * a2;
        double d[4];
        if (th2 > 0.) {
            const double th = std::sqrt(th2), half = th * 0.5, k = std::sin(half) / th;
            d[0] = std::cos(half); d[1] = a0 * k; d[2] = a1 * k; d[3] = a2 * k;
        } else {
            d[0] = 1.; d[1] = a0 * 0.5; d[2] = a1 * 0.5; d[3] = a2 * 0.5;
        }
        // quat_prod(d, q), quat.cpp:33-38, then arma::normalise
        double r[4];
        r[0] = ((d[0] * q[0] - d[1] * q[1]) - d[2] * q[2]) - d[3] * q[3];
        r[1] = ((d[0] * q[1] + d[1] * q[0]) + d[2] * q[3]) - d[3] * q[2];
        r[2] = ((d[0] * q[2] - d[1] * q[3]) + d[2] * q[0]) + d[3] * q[1];
        r[3] = ((d[0] * q[3] + d[1] * q[2]) - d[2] * q[1]) + d[3] * q[0];
        const double nrm = std::sqrt(((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]) + r[3] * r[3]);
        for (int c = 0; c < 4; ++c) {
            q[c] = r[c] / nrm;
            out[4 * i + c] = q[c];
        }
    }
    return true;
}

// SyncProblemPrivate::SetGyroQuaternions(const int64_t*, const double*, size_t),
// core_private.cpp:142-190, integer quirks included.
IngestStatus resample_variable_rate(const int64_t* ts_us, const double* quats, size_t count,
                                    std::vector<double>& out_quats, double& sample_rate,
                                    double& first_timestamp, std::string& err) {
    constexpr uint64_t kMicro = 1000000ULL;
    if (count < 2) { err = "set-gyro-quaternions: need at least 2 samples"; return IngestStatus::Invalid; }
    const uint64_t span = (uint64_t)(ts_us[count - 1] - ts_us[0]);
    if (span == 0) { err = "set-gyro-quaternions: zero time span"; return IngestStatus::Invalid; }
    const uint64_t rate_uhz = kMicro * kMicro * (uint64_t)count / span;                   // :146-147
    const int rate_hz = int(std::round((double)rate_uhz / 50. / (double)kMicro) * 50);   // :148-149
    if (rate_hz <= 0) { err = "set-gyro-quaternions: sample rate rounds to zero"; return IngestStatus::Invalid; }
    const uint64_t t_last = (uint64_t)ts_us[count - 1];
    std::vector<uint64_t> grid;
    int tick = (int)std::ceil((double)((uint64_t)(ts_us[0] * (int64_t)rate_hz) / kMicro));  // :152
    while (kMicro * (uint64_t)tick / (uint64_t)rate_hz < t_last) {                           // :153
        grid.push_back(kMicro * (uint64_t)tick / (uint64_t)rate_hz);
        ++tick;
    }
    for (size_t i = 1; i < count; ++i) {  // :157-164
        if (ts_us[i - 1] > ts_us[i]) {
            err = "set-gyro-quaternions:  timestamps out of order at pos " + std::to_string(i) + " (" +
                  std::to_string(ts_us[i - 1]) + " > " + std::to_string(ts_us[i]) + ")";
            return IngestStatus::OutOfOrder;
        }
    }
    if (grid.size() < 2) { err = "set-gyro-quaternions: fewer than 2 resampled samples"; return IngestStatus::Invalid; }
    out_quats.resize(grid.size() * 4);
    for (size_t j = 0; j < grid.size(); ++j) {  // :166-182
        const uint64_t t = grid[j];
        const int64_t* it = std::lower_bound(ts_us, ts_us + count, t,
                                             [](int64_t a, uint64_t b) { return (uint64_t)a < b; });
        const size_t hi = (size_t)(it - ts_us);
        double* dst = &out_quats[4 * j];
        if (hi > 0) {
            const double frac =
                1. * (double)(t - (uint64_t)ts_us[hi - 1]) / (double)(ts_us[hi] - ts_us[hi - 1]);
            slerp4(quats + 4 * (hi - 1), quats + 4 * hi, frac, dst);
        } else {
            for (int c = 0; c < 4; ++c) dst[c] = quats[4 * hi + c];
        }
        if (!(std::isfinite(dst[0]) && std::isfinite(dst[1]) && std::isfinite(dst[2]) && std::isfinite(dst[3]))) {
            err = "set-gyro-quaternions: non-finite sample after interpolation";  // :180-181
            return IngestStatus::NonFinite;
        }
    }
    sample_rate = 1. * rate_hz;                              // :183
    first_timestamp = 1. * (double)grid[0] / (double)kMicro;  // :184
    if (!std::isfinite(sample_rate)) { err = "set-gyro-quaternions: non-finite sample rate. wtf?"; return IngestStatus::NonFinite; }
    if (!std::isfinite(first_timestamp)) { err = "set-gyro-quaternions: non-finite first timestamp. wtf?"; return IngestStatus::NonFinite; }
    return IngestStatus::Ok;
}

}  // namespace rs
