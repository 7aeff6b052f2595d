#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <thread>
#include <vector>
#include <atomic>
#include <cstring>
inline bool copy_checked(double* dst, const double* src, size_t n, double* lo, double* hi) {
    double acc = 0.0, l = lo ? *lo : 0.0, h = hi ? *hi : 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double v = src[i];
        dst[i] = v;
        acc += v * 0.0;
        if (lo) { l = v < l ? v : l; h = v > h ? v : h; }
    }
    if (lo) { *lo = l; *hi = h; }
    return acc == 0.0;
}
template <bool NT>
__attribute__((target("avx2"))) bool copy_checked_avx2(double* dst, const double* src, size_t n, double* lo, double* hi) {
    double acc = 0.0, l = lo ? *lo : 0.0, h = hi ? *hi : 0.0;
    size_t i = 0;
    for (; i < n && ((uintptr_t)(dst + i) & 31); ++i) {
        const double v = src[i]; dst[i] = v; acc += v * 0.0; l = v < l ? v : l; h = v > h ? v : h;
    }
    __m256d va = _mm256_setzero_pd(), vl = _mm256_set1_pd(l), vh = _mm256_set1_pd(h);
    const __m256d zero = _mm256_setzero_pd();
    for (; i + 4 <= n; i += 4) {
        const __m256d v = _mm256_loadu_pd(src + i);
        if (NT) _mm256_stream_pd(dst + i, v); else _mm256_store_pd(dst + i, v);
        va = _mm256_add_pd(va, _mm256_mul_pd(v, zero));
        vl = _mm256_min_pd(v, vl);
        vh = _mm256_max_pd(v, vh);
    }
    double ta[4], tl[4], th[4];
    _mm256_storeu_pd(ta, va); _mm256_storeu_pd(tl, vl); _mm256_storeu_pd(th, vh);
    for (int k = 0; k < 4; ++k) { acc += ta[k]; l = tl[k] < l ? tl[k] : l; h = th[k] > h ? th[k] : h; }
    for (; i < n; ++i) {
        const double v = src[i]; dst[i] = v; acc += v * 0.0; l = v < l ? v : l; h = v > h ? v : h;
    }
    if (NT) _mm_sfence();
    if (lo) { *lo = l; *hi = h; }
    return acc == 0.0;
}
int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 7;
    const size_t F = 3300, N = 200, total = F * N;
    std::vector<double> tsa(total, 1.0), tsb(total, 2.0), ra(3 * total, 0.5), rb(3 * total, 0.25);
    double* st = (double*)aligned_alloc(4096, (8 * total + 64) * 8);
    for (size_t i = 0; i < 8 * total; ++i) st[i] = 0;
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 5; ++rep) {
            std::atomic<size_t> next{0};
            std::atomic<int> bad{0};
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back([&] {
                for (;;) {
                    size_t lo = next.fetch_add(16);
                    if (lo >= F) break;
                    for (size_t f = lo; f < std::min(F, lo + 16); ++f) {
                        double l = tsa[f * N], h = l;
                        bool ok;
                        double *d0 = st + f * N, *d1 = st + total + f * N, *d2 = st + 2 * total + 3 * f * N, *d3 = st + 5 * total + 3 * f * N;
                        if (mode == 0) ok = copy_checked(d2, &ra[3 * f * N], 3 * N, 0, 0) & copy_checked(d3, &rb[3 * f * N], 3 * N, 0, 0) & copy_checked(d0, &tsa[f * N], N, &l, &h) & copy_checked(d1, &tsb[f * N], N, &l, &h);
                        else if (mode == 1) ok = copy_checked_avx2<false>(d2, &ra[3 * f * N], 3 * N, 0, 0) & copy_checked_avx2<false>(d3, &rb[3 * f * N], 3 * N, 0, 0) & copy_checked_avx2<false>(d0, &tsa[f * N], N, &l, &h) & copy_checked_avx2<false>(d1, &tsb[f * N], N, &l, &h);
                        else if (mode == 3) { memcpy(d2, &ra[3 * f * N], 24 * N); memcpy(d3, &rb[3 * f * N], 24 * N); memcpy(d0, &tsa[f * N], 8 * N); memcpy(d1, &tsb[f * N], 8 * N); ok = true; }
                        else ok = copy_checked_avx2<true>(d2, &ra[3 * f * N], 3 * N, 0, 0) & copy_checked_avx2<true>(d3, &rb[3 * f * N], 3 * N, 0, 0) & copy_checked_avx2<true>(d0, &tsa[f * N], N, &l, &h) & copy_checked_avx2<true>(d1, &tsb[f * N], N, &l, &h);
                        if (!ok) bad++;
                    }
                }
            });
            for (auto& x : th) x.join();
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (rep == 4) printf("mode %d threads %d: %.3f ms (%.1f GB/s) bad %d\n", mode, T, ms, 8 * total * 8 / ms / 1e6, bad.load());
        }
    }
}
