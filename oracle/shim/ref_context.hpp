// oracle/shim/ref_context.hpp — TEST INFRASTRUCTURE ONLY.
// Thread-local context that lets the unmodified reference code draw from the pinned RNG of
// oracle_math.hpp instead of its random_device-seeded mt19937 (inline_utils.hpp:13-17).
#pragma once
#include <cstdint>
#include <vector>

namespace rssync_ref {

struct Context {
    uint64_t seed = 100, stream = 1, call_no = 0;
    // PreSync / DebugPreSync: set per item by the shadow std::for_each (oracle/shim/execution)
    uint64_t offset_idx = 0;
    int64_t frame_id = 0;
    // Sync initialisation (serial loop, core_private.cpp:218-223): frames in iteration order
    bool sync_mode = false;
    const std::vector<int64_t>* sync_frames = nullptr;
    long sync_pos = -1;
    uint32_t k = 0;
};
Context& ctx();                    // this thread's context
extern Context g_call;             // what the driver set for the current API call
extern uint64_t g_foreach_count;   // par for_each invocations since the driver reset it
extern int g_threads;

int pinned_mtrand(int iter, int line, int lo, int hi);

}  // namespace rssync_ref
