// Pinned counter-based RNG of the randomised translation estimator (replaces mtrand,
// inline_utils.hpp:13-17 of the reference, which is seeded from std::random_device and therefore
// irreproducible).  Stateless: every draw is a hash of (seed, stream, call_no, offset_idx,
// frame_id, iter, k).  Shared by host (key prefixes) and device (draws).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define RS_HD __host__ __device__ __forceinline__
#else
#define RS_HD inline
#endif

namespace rs {

RS_HD uint64_t mix64(uint64_t z) {
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
// (seed, stream, call_no, offset_idx) part of the key; the frame id is folded in per task
RS_HD uint64_t rng_prefix(uint64_t seed, uint64_t stream, uint64_t call_no, uint64_t offset_idx) {
    uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ULL);
    h = mix64(h ^ (stream + (call_no << 8)));
    h = mix64(h ^ offset_idx);
    return h;
}
RS_HD uint64_t rng_task_key(uint64_t prefix, int64_t frame_id) {
    return mix64(prefix ^ (uint64_t)frame_id);
}

}  // namespace rs
