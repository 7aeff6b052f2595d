// Host-side gyro ingest (see host_ingest.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace rs {

enum class IngestStatus { Ok = 0, Invalid = 1, NonFinite = 2, OutOfOrder = 3 };

// ndspline::make (ndspline.cpp:13-19): n quaternions (w,x,y,z) -> n records of 16 doubles
// {y[4], b[4], c[4], d[4]}.
void build_spline_records(const double* quats, size_t n, double* rec /* n * 16 doubles */);

// variable-rate SetGyroQuaternions (core_private.cpp:142-190): resample onto the uniform
// integer-microsecond grid by slerp.
IngestStatus resample_variable_rate(const int64_t* ts_us, const double* quats, size_t count,
                                    std::vector<double>& out_quats, double& sample_rate,
                                    double& first_timestamp, std::string& err);

}  // namespace rs
