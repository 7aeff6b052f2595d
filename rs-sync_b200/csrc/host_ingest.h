// Host-side gyro ingest (see host_ingest.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace rs {

enum class IngestStatus { Ok = 0, Invalid = 1, NonFinite = 2, OutOfOrder = 3 };

// ndspline::make (ndspline.cpp:13-19), the sequential half: the tridiagonal system of the natural
// cubic spline on unit knots (minispline.cpp:3-32) for the four quaternion components, eliminated
// downwards and upwards.  rhs (4 n doubles, interleaved like the input) and diag (n doubles) are what
// is left of it: the second-derivative coefficient of sample i, component c is rhs[4 i + c] / diag[i]
// (:34).  The divisions and the b / d coefficients (:38-44) are independent per sample and are
// finished on the device (launch_spline_finish, engine.h).
void build_spline_system(const double* quats, size_t n, double* rhs /* 4 n */, double* diag /* n */);

// variable-rate SetGyroQuaternions (core_private.cpp:142-190): resample onto the uniform
// integer-microsecond grid by slerp.
IngestStatus resample_variable_rate(const int64_t* ts_us, const double* quats, size_t count,
                                    std::vector<double>& out_quats, double& sample_rate,
                                    double& first_timestamp, std::string& err);

// optdata_fill_gyro (core_testcode.cpp:37-53), the step before SetGyroQuaternions: q_0 = identity,
// q_i = normalise(quat_from_aa(w_i (t_i - t_{i-1})) (x) q_{i-1}) (quat.cpp:5-17, 33-38).  `orient`
// is a 3-character gyro_orientation string (core_testcode.cpp:186-190) or null for "XYZ":
// character i names the input axis routed to output axis i, lower case flips its sign.
// Returns false for a malformed orientation string.
bool integrate_gyro(const double* timestamps_s, const double* gyro_xyz, size_t count,
                    const char* orient, double* quats_out /* count * 4 */);

}  // namespace rs
