// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
// C driver around the UNMODIFIED reference translation units (compiled from /root/reference by
// oracle/Makefile `ref`, against oracle/shim).  Exposes the reference's ISyncProblem plus a few
// of its non-static free functions (opt_compute_problem, opt_guess_translational_motion,
// core_private.cpp:15,34) so tests can pin the oracle stage by stage against the reference's own
// code, and bench.py --impl reference can time it.  The only behavioural substitution is the RNG
// (see shim/inline_utils.hpp).
#include <core_private.hpp>
#include <quat.hpp>

#include "oracle_math.hpp"
#include "shim/ref_context.hpp"

arma::mat opt_compute_problem(int64_t frame, double gyro_delay, const OptData& data);
arma::vec3 opt_guess_translational_motion(const arma::mat& problem, int max_iters);

namespace rssync_ref {
Context g_call;
uint64_t g_foreach_count = 0;
int g_threads = 1;
Context& ctx() {
    static thread_local Context c;
    return c;
}
int pinned_mtrand(int iter, int line, int lo, int hi) {
    Context& c = ctx();
    // the first of the two call sites (core_private.cpp:42) starts a new hypothesis
    const bool first_site = (line == 42);
    if (first_site) {
        c.k = 0;
        if (c.sync_mode && iter == 0) c.sync_pos++;
    }
    int64_t frame = c.frame_id;
    uint64_t off = c.offset_idx;
    if (c.sync_mode) {
        frame = (*c.sync_frames)[(size_t)c.sync_pos];
        off = 0;
    }
    const uint64_t key = orc::rng_task_key(c.seed, c.stream, c.call_no, off, frame);
    const uint32_t n = (uint32_t)(hi - lo + 1);
    return lo + (int)orc::rng_index(key, (uint32_t)iter, c.k++, n);
}
}  // namespace rssync_ref

using rssync_ref::g_call;

namespace {
struct Ref {
    SyncProblemPrivate* sp;
    std::vector<int64_t> sync_frames;
};
void begin_call(uint64_t seed, uint64_t stream, uint64_t call_no) {
    g_call = rssync_ref::Context();
    g_call.seed = seed;
    g_call.stream = stream;
    g_call.call_no = call_no;
    rssync_ref::g_foreach_count = 0;
    rssync_ref::ctx() = g_call;
}
}  // namespace

extern "C" {

void* ref_create() {
    Ref* r = new Ref();
    r->sp = static_cast<SyncProblemPrivate*>(CreateSyncProblem());
    return r;
}
void ref_destroy(void* h) {
    Ref* r = (Ref*)h;
    delete r->sp;
    delete r;
}
void ref_set_threads(int t) { rssync_ref::g_threads = t < 1 ? 1 : t; }

void ref_set_gyro_fixed(void* h, const double* q, size_t n, double sr, double t0) {
    ((Ref*)h)->sp->SetGyroQuaternions(q, n, sr, t0);
}
void ref_set_gyro_var(void* h, const int64_t* ts, const double* q, size_t n) {
    ((Ref*)h)->sp->SetGyroQuaternions(ts, q, n);
}
void ref_set_track(void* h, int64_t frame, const double* ts_a, const double* ts_b, const double* ra,
                   const double* rb, size_t n) {
    ((Ref*)h)->sp->SetTrackResult(frame, ts_a, ts_b, ra, rb, n);
}
// frame ids in the iteration order of the reference's unordered_map (core_private.cpp:65,218)
int ref_frame_order(void* h, int64_t* out, int cap) {
    int n = 0;
    for (auto& kv : ((Ref*)h)->sp->problem.frame_data) {
        if (out && n < cap) out[n] = kv.first;
        ++n;
    }
    return n;
}
double ref_gyro_rate(void* h) { return ((Ref*)h)->sp->problem.sample_rate; }
double ref_gyro_start(void* h) { return ((Ref*)h)->sp->problem.quats_start; }

void ref_presync(void* h, uint64_t seed, uint64_t call_no, double initial, int64_t fb, int64_t fe,
                 double step, double radius, double* cost, double* delay) {
    begin_call(seed, orc::kStreamPreSync, call_no);
    auto r = ((Ref*)h)->sp->PreSync(initial, fb, fe, step, radius);
    *cost = r.first;
    *delay = r.second;
}
void ref_debug_presync(void* h, uint64_t seed, uint64_t stream, uint64_t call_no, double initial,
                       int64_t fb, int64_t fe, double radius, double* delays, double* costs, int n) {
    begin_call(seed, stream, call_no);
    ((Ref*)h)->sp->DebugPreSync(initial, fb, fe, radius, delays, costs, n);
}
void ref_sync(void* h, uint64_t seed, uint64_t call_no, double initial, int64_t fb, int64_t fe,
              double center, double radius, double* cost, double* delay) {
    Ref* r = (Ref*)h;
    begin_call(seed, orc::kStreamSyncInit, call_no);
    r->sync_frames.clear();
    for (auto& kv : r->sp->problem.frame_data)  // same iteration order as core_private.cpp:218
        if (!(kv.first < fb || kv.first > fe)) r->sync_frames.push_back(kv.first);
    g_call.sync_mode = true;
    g_call.sync_frames = &r->sync_frames;
    g_call.sync_pos = -1;
    rssync_ref::ctx() = g_call;
    auto res = r->sp->Sync(initial, fb, fe, center, radius);
    *cost = res.first;
    *delay = res.second;
}

// ---- stage probes ---------------------------------------------------------------------------
void ref_spline_eval(void* h, const double* x, int n, double* out) {
    const OptData& d = ((Ref*)h)->sp->problem;
    for (int i = 0; i < n; ++i) {
        arma::mat v = d.quats.eval(x[i]);
        for (int c = 0; c < 4; ++c) out[4 * i + c] = v[c];
    }
}
void ref_problem_matrix(void* h, int64_t frame, double delay, double* P) {
    arma::mat m = opt_compute_problem(frame, delay, ((Ref*)h)->sp->problem);
    for (arma::uword i = 0; i < m.n_rows; ++i)
        for (int c = 0; c < 3; ++c) P[3 * i + c] = m(i, c);
}
void ref_guess_motion(void* h, uint64_t seed, uint64_t stream, uint64_t call_no, uint64_t offset_idx,
                      int64_t frame, double delay, int iters, double* m3) {
    begin_call(seed, stream, call_no);
    rssync_ref::Context& c = rssync_ref::ctx();
    c.offset_idx = offset_idx;
    c.frame_id = frame;
    arma::mat P = opt_compute_problem(frame, delay, ((Ref*)h)->sp->problem);
    arma::vec3 v = opt_guess_translational_motion(P, iters);
    for (int i = 0; i < 3; ++i) m3[i] = v[i];
}
void ref_loss(void* h, int64_t frame, double delay, const double* m3, double k, double* loss3,
              double* loss5, double* ddelay, double* grad3) {
    FrameState fs(frame, &((Ref*)h)->sp->problem);
    fs.var_k = k;
    arma::mat d(1, 1), m(3, 1), l3, l5, jd, jm;
    d[0] = delay;
    for (int i = 0; i < 3; ++i) m[i] = m3[i];
    fs.Loss(d, m, l3);
    fs.Loss(d, m, l5, jd, jm);
    *loss3 = l3[0];
    *loss5 = l5[0];
    *ddelay = jd[0];
    for (int i = 0; i < 3; ++i) grad3[i] = jm[i];
}
void ref_slerp(const double* p, const double* q, double t, double* out) {
    arma::vec4 r = quat_slerp(arma::vec4(p), arma::vec4(q), t);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
}

}  // extern "C"
