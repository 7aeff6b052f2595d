// Host-side gyro ingest of the engine: variable-rate resampling and the natural-cubic-spline
// coefficient build.  Both are short sequential recurrences over the gyro track (O(n), run once
// per SetGyroQuaternions call, ~1 ms for a minute of 1 kHz data), and their results feed the
// finite-difference delay gradient of Sync, which amplifies 1-ulp differences by 1/(2h) = 5e5 —
// so they are evaluated in one fixed order on the host and only the finished 128-byte spline
// records travel to the GPU.  Compile with -ffp-contract=off.
#include "host_ingest.h"

#include <algorithm>
#include <cmath>
#include <thread>

namespace rs {

namespace {

struct TriRow {
    double sub, diag, sup, rhs;
};

// One component of ndspline::make (ndspline.cpp:13-19): spline::set_points, minispline.cpp:3-46.
// Unit-spaced knots; interior rows [1/3, 4/3, 1/3] with rhs y[i+1] - 2 y[i] + y[i-1], natural
// boundary rows [., 2, 0] / [0, 2, .] with rhs 0; eliminate downwards, then upwards, divide.
void solve_component(const double* y, size_t n, size_t stride, int comp, double* rec) {
    std::vector<TriRow> row(n);
    row[0] = TriRow{0.0, 2.0, 0.0, 0.0};
    row[n - 1] = TriRow{0.0, 2.0, 0.0, 0.0};
    for (size_t i = 1; i + 1 < n; ++i) {
        const double yi = y[i * stride];
        row[i].sub = 1.0 / 3.0;
        row[i].diag = 2.0 / 3.0 * 2.0;
        row[i].sup = 1.0 / 3.0;
        row[i].rhs = (y[(i + 1) * stride] - 2 * yi) + y[(i - 1) * stride];
    }
    for (size_t i = 0; i + 2 < n; ++i) {  // minispline.cpp:22-26
        TriRow& cur = row[i];
        TriRow& nxt = row[i + 1];
        const double f = 1. / cur.diag * nxt.sub;
        nxt.sub -= cur.diag * f;
        nxt.diag -= cur.sup * f;
        nxt.rhs -= cur.rhs * f;
    }
    for (size_t i = n - 1; i > 1; --i) {  // :28-32
        TriRow& cur = row[i];
        TriRow& prv = row[i - 1];
        const double f = 1. / cur.diag * prv.sup;
        prv.diag -= cur.sub * f;
        prv.sup -= cur.diag * f;
        prv.rhs -= cur.rhs * f;
    }
    // second-derivative coefficients c, then d and b per interval (:34-44)
    double c_prev = 0.0, b_prev = 0.0, d_prev = 0.0;
    double c_i = row[0].rhs / row[0].diag;
    for (size_t i = 0; i < n; ++i) {
        // group g of record i (y, b, c, d = 0..3) lives at group position g ^ (i & 3): see
        // rec_groups in device_math.cuh
        const size_t sw = (i & 3) * 4;
        double* out = rec + i * 16 + comp;
        const double yi = y[i * stride];
        double b, d;
        if (i + 1 < n) {
            const double c_next = row[i + 1].rhs / row[i + 1].diag;
            d = 1.0 / 3.0 * (c_next - c_i);
            b = (y[(i + 1) * stride] - yi) - 1.0 / 3.0 * (2.0 * c_i + c_next);
            out[0 ^ sw] = yi; out[4 ^ sw] = b; out[8 ^ sw] = c_i; out[12 ^ sw] = d;
            c_prev = c_i; b_prev = b; d_prev = d;
            c_i = c_next;
        } else {
            d = 0.0;
            b = (3.0 * d_prev + 2.0 * c_prev) + b_prev;
            out[0 ^ sw] = yi; out[4 ^ sw] = b; out[8 ^ sw] = c_i; out[12 ^ sw] = d;
        }
    }
}

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

void build_spline_records(const double* quats, size_t n, double* rec, bool one_thread_per_component) {
    if (!one_thread_per_component) {
        for (int comp = 0; comp < 4; ++comp) solve_component(quats + comp, n, 4, comp, rec);
        return;
    }
    // the four components are independent sequential solves: one host thread each
    std::thread th[3];
    for (int comp = 1; comp < 4; ++comp)
        th[comp - 1] = std::thread([=]() { solve_component(quats + comp, n, 4, comp, rec); });
    solve_component(quats, n, 4, 0, rec);
    for (auto& t : th) t.join();
}

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
