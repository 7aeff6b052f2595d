// ISyncProblem / CreateSyncProblem over the C ABI: the C++ face a core_testcode-style caller
// links against (reference: src/core/public/rssync.h:9-31, implementation class
// SyncProblemPrivate core_private.hpp:44-61).  A non-zero status from the C ABI becomes the
// reference's panic convention (src/core_support/panic.cpp:7-15).
#include "rssync.h"

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "rssync_b200.h"

namespace {

[[noreturn]] void panic_to_file(const char* reason) {
    if (FILE* f = std::fopen("panic.txt", "w")) {
        std::fprintf(f, "%s\n", reason);
        std::fclose(f);
    }
    std::fprintf(stderr, "rssync panic: %s\n", reason);
    std::exit(1);
}

class SyncProblemB200 final : public ISyncProblem {
   public:
    SyncProblemB200() {
        // RSSYNC_DEVICES="0,1,2,3": spread the problem over those GPUs (rssync_create_multi), so that an
        // unmodified core_testcode-style caller uses all of them; unset: the current device
        std::vector<int> devs;
        if (const char* e = std::getenv("RSSYNC_DEVICES")) {
            for (const char* c = e; *c;) {
                char* end = nullptr;
                const long v = std::strtol(c, &end, 10);
                if (end == c) break;
                devs.push_back((int)v);
                c = (*end == ',') ? end + 1 : end;
            }
        }
        int rc = devs.size() > 1 ? rssync_create_multi(devs.data(), (int)devs.size(), &p_) : rssync_create(&p_);
        if (rc != RSSYNC_OK)
            panic_to_file(p_ ? rssync_last_error(p_) : "rssync: no usable CUDA device (there is no CPU fallback)");
    }
    ~SyncProblemB200() override { rssync_destroy(p_); }

    void SetGyroQuaternions(const double* data, size_t count, double sample_rate,
                            double first_timestamp) override {
        check(rssync_set_gyro_fixed(p_, data, count, sample_rate, first_timestamp));
    }
    void SetGyroQuaternions(const int64_t* timestamps_us, const double* quats, size_t count) override {
        check(rssync_set_gyro_var(p_, timestamps_us, quats, count));
    }
    void SetTrackResult(int64_t frame, const double* ts_a, const double* ts_b, const double* rays_a,
                        const double* rays_b, size_t count) override {
        check(rssync_set_track(p_, frame, ts_a, ts_b, rays_a, rays_b, count));
    }
    std::pair<double, double> PreSync(double initial_delay, int64_t frame_begin, int64_t frame_end,
                                      double search_step, double search_radius) override {
        double cost = 0, delay = 0;
        check(rssync_presync(p_, initial_delay, frame_begin, frame_end, search_step, search_radius, &cost, &delay));
        return {cost, delay};
    }
    std::pair<double, double> Sync(double initial_delay, int64_t frame_begin, int64_t frame_end,
                                   double search_center, double search_radius) override {
        double cost = 0, delay = 0;
        check(rssync_sync(p_, initial_delay, frame_begin, frame_end, search_center, search_radius, &cost, &delay));
        // progress lines of core_private.cpp:330 (printed for every iteration that did not
        // leave the loop through one of its two `break`s)
        int n = rssync_last_sync_trace(p_, nullptr, nullptr, 0);
        std::vector<double> d(n), s(n);
        rssync_last_sync_trace(p_, d.data(), s.data(), n);
        int printed = (n < 400) ? n - 1 : n;
        for (int i = 0; i < printed; ++i) std::cerr << d[i] << " " << s[i] << std::endl;
        return {cost, delay};
    }
    void DebugPreSync(double initial_delay, int64_t frame_begin, int64_t frame_end, double search_radius,
                      double* delays, double* costs, int point_count) override {
        check(rssync_debug_presync(p_, initial_delay, frame_begin, frame_end, search_radius, delays, costs,
                                   point_count));
    }

   private:
    void check(int rc) {
        if (rc != RSSYNC_OK) panic_to_file(rssync_last_error(p_));
    }
    rssync_problem* p_ = nullptr;
};

}  // namespace

ISyncProblem* CreateSyncProblem() { return new SyncProblemB200(); }

ISyncProblem::~ISyncProblem() {}
