// rssync.h — C++ interface of the synchronisation loss engine, binary-compatible with
// src/core/public/rssync.h of VladimirP1/rs-sync (same class name, same virtual-member order,
// same factory signature with C++ linkage), so a caller compiled against the reference header
// (core_testcode, GyroFlow-side C++) links against librssync_b200.so unchanged.
//
// Units: seconds unless a name ends in _us.  PreSync and Sync return {cost, delay}.
// Errors follow the reference's convention (src/core_support/panic.cpp:7-15): the reason is
// written to ./panic.txt and the process exits with status 1.
#pragma once

#include <cstddef>
#include <cstdint>
#include <utility>

#if defined(_WIN32)
#if defined(RSSYNC_EXPORTS)
#define RSSYNC_API __declspec(dllexport)
#else
#define RSSYNC_API __declspec(dllimport)
#endif
#else
#define RSSYNC_API
#endif

class ISyncProblem {
   public:
    virtual ~ISyncProblem();

    // fixed-rate gyro orientation track: count x (w,x,y,z)
    virtual void SetGyroQuaternions(const double* data, size_t count, double sample_rate,
                                    double first_timestamp) = 0;
    // variable-rate track with integer microsecond timestamps; resampled internally
    virtual void SetGyroQuaternions(const int64_t* timestamps_us, const double* quats,
                                    size_t count) = 0;
    // rays of features tracked from `frame` to the next frame, with per-ray timestamps
    virtual void SetTrackResult(int64_t frame, const double* ts_a, const double* ts_b,
                                const double* rays_a, const double* rays_b, size_t count) = 0;
    // brute-force search over frames [frame_begin, frame_end)
    virtual std::pair<double, double> PreSync(double initial_delay, int64_t frame_begin,
                                              int64_t frame_end, double search_step,
                                              double search_radius) = 0;
    // refinement over frames [frame_begin, frame_end]
    virtual std::pair<double, double> Sync(double initial_delay, int64_t frame_begin,
                                           int64_t frame_end, double search_center,
                                           double search_radius) = 0;
    // PreSync's loss curve on point_count equally spaced delays
    virtual void DebugPreSync(double initial_delay, int64_t frame_begin, int64_t frame_end,
                              double search_radius, double* delays, double* costs,
                              int point_count) = 0;
};

RSSYNC_API ISyncProblem* CreateSyncProblem();
