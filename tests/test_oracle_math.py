"""Known-answer tests that pin the CPU oracle (SURVEY.md §8(c) "pins that can be constructed").

The reference ships no tests or golden vectors, so the oracle is pinned against SciPy / analytic
answers here and against the unmodified reference sources in test_oracle_vs_reference.py.
"""
import numpy as np
import pytest
from scipy.interpolate import CubicSpline
from scipy.spatial.transform import Rotation, Slerp

from conftest import rel_err, workload


def test_spline_equals_scipy_natural_cubic(oracle_loader, w_tiny):
    """pin 1: minispline.cpp:3-46 == natural CubicSpline on unit knots"""
    o = oracle_loader.OracleProblem().load(w_tiny)
    rec = o.spline()
    n = rec.shape[0]
    cs = CubicSpline(np.arange(n), w_tiny.quats, bc_type="natural")
    x = np.random.default_rng(0).uniform(0, n - 1, 2000)
    assert np.max(np.abs(o.spline_eval(x) - cs(x))) < 1e-13
    # coefficient layout: y, b, c, d per knot and component
    assert np.array_equal(rec[:, 0:4], w_tiny.quats)
    assert np.max(np.abs(rec[:-1, 12:16] - cs.c[0][: n - 1])) < 1e-12  # cubic term
    assert np.all(rec[0, 8:12] == 0) and np.all(rec[-1, 12:16] == 0)   # natural ends


def test_spline_extrapolation_quirks(oracle_loader, w_tiny):
    """minispline.cpp:48-55: linear on the left; on the right h = x - min(floor(x), n), so the
    value jumps back to y[n-1] at x = n (a6 in SURVEY §8a) — replicated, not fixed."""
    o = oracle_loader.OracleProblem().load(w_tiny)
    rec = o.spline()
    n = rec.shape[0]
    y0, b0 = rec[0, 0:4], rec[0, 4:8]
    yl, bl = rec[-1, 0:4], rec[-1, 4:8]
    v = o.spline_eval(np.array([-2.5, n - 0.5, float(n), n + 3.25]))
    assert np.allclose(v[0], y0 - 2.5 * b0, atol=1e-15)
    assert np.allclose(v[1], yl + 0.5 * bl, atol=1e-15)
    assert np.allclose(v[2], yl, atol=0)            # sawtooth: h = 0 at x = n
    assert np.allclose(v[3], yl + 3.25 * bl, atol=1e-15)  # h = x - n beyond the end


def test_log1p_accuracy(oracle_loader):
    rng = np.random.default_rng(1)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 50000), rng.uniform(0, 4, 50000), [0.0, np.inf]])
    got = oracle_loader.log1p(x)
    ref = np.log1p(x)
    fin = np.isfinite(ref) & (ref > 0)
    ulp = np.spacing(ref[fin])
    assert np.max(np.abs(got[fin] - ref[fin]) / ulp) <= 2.0
    assert got[-2] == 0.0 and np.isinf(got[-1])
    assert np.isnan(oracle_loader.log1p(np.array([np.nan]))[0])


def test_slerp_equals_scipy(oracle_loader):
    """pin 2: quat_slerp (quat.cpp:55-74) == scipy Slerp (w-first <-> scalar-last)"""
    import ctypes as C
    L = oracle_loader.lib()
    rng = np.random.default_rng(2)
    for _ in range(50):
        q = rng.normal(size=(2, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        t = float(rng.uniform())
        out = np.empty(4)
        L.orc_slerp(q[0].ctypes.data_as(oracle_loader.c_double_p), q[1].ctypes.data_as(oracle_loader.c_double_p),
                    t, out.ctypes.data_as(oracle_loader.c_double_p))
        r = Rotation.from_quat(q[:, [1, 2, 3, 0]])
        ref = Slerp([0, 1], r)([t]).as_quat()[0][[3, 0, 1, 2]]
        if np.dot(ref, out) < 0:
            ref = -ref
        assert np.max(np.abs(out - ref)) < 1e-12


def test_problem_matrix_matches_textbook(oracle_loader, w_tiny, synth_mod):
    """opt_compute_problem (core_private.cpp:15-32) written the reference's way in numpy:
    normalise(spline) -> conj(a) (x) p (x) a via two Hamilton products -> cross."""
    w = w_tiny
    o = oracle_loader.OracleProblem().load(w)
    fid, delay = int(w.frame_ids[3]), 0.0123
    P = o.problem_matrix(fid, delay, w.n_rays)
    i = 3
    xa = (w.ts_a[i] - w.gyro_t0 + delay) * w.gyro_rate
    xb = (w.ts_b[i] - w.gyro_t0 + delay) * w.gyro_rate
    qa = o.spline_eval(xa)
    qb = o.spline_eval(xb)
    qa /= np.linalg.norm(qa, axis=1, keepdims=True)
    qb /= np.linalg.norm(qb, axis=1, keepdims=True)
    ar = synth_mod.quat_rotate(synth_mod.quat_conj(qa), w.rays_a[i])
    br = synth_mod.quat_rotate(synth_mod.quat_conj(qb), w.rays_b[i])
    ref = np.cross(ar, br)
    assert np.max(np.abs(P - ref)) < 5e-15


def test_pure_rotation_scene_has_zero_rows(oracle_loader, synth_mod):
    """pin 5: without translation every row of P vanishes at the true delay"""
    w = synth_mod.make_workload("tiny", noise_px=0.0, outlier_frac=0.0)
    # re-generate rays_b for a camera that does not translate: b ray = R(q_b) R(q_a)^-1 a ray
    track = synth_mod.QuatTrack(w.quats, w.gyro_rate, w.gyro_t0)
    qa = track(w.ts_a + w.true_delay[:, None])
    qb = track(w.ts_b + w.true_delay[:, None])
    world = synth_mod.quat_rotate(synth_mod.quat_conj(qa), w.rays_a)
    rays_b = synth_mod.quat_rotate(qb, world)
    o = oracle_loader.OracleProblem()
    o.SetGyroQuaternions(w.quats, w.quats.shape[0], w.gyro_rate, w.gyro_t0)
    o.SetTrackResult(7, w.ts_a[2], w.ts_b[2], w.rays_a[2], rays_b[2], w.n_rays)
    P = o.problem_matrix(7, float(w.true_delay[2]), w.n_rays)
    assert np.max(np.abs(P)) < 1e-9
    Pw = o.problem_matrix(7, float(w.true_delay[2]) + 0.02, w.n_rays)
    assert np.max(np.abs(Pw)) > 1e-4


def test_loss_gradient_closed_form(oracle_loader, w_small):
    """pin 4: closed form of the forward-mode chain == numeric differentiation; grad . m == 0"""
    o = oracle_loader.OracleProblem().load(w_small)
    fid = int(w_small.frame_ids[9])
    m, k = o.guess_motion(fid, 0.03, 200, 3, 0, 0)
    m = m + np.array([0.01, -0.02, 0.015])
    f0, g = o.loss5(fid, 0.03, m, k)
    num = np.empty(3)
    for c in range(3):
        e = np.zeros(3)
        e[c] = 1e-6
        num[c] = (o.loss5(fid, 0.03, m + e, k)[0] - o.loss5(fid, 0.03, m - e, k)[0]) / 2e-6
    assert np.max(np.abs(num - g)) < 1e-4 * max(1.0, np.max(np.abs(g)))
    assert abs(np.dot(g, m)) < 1e-8 * np.linalg.norm(g) * np.linalg.norm(m)
    # the two loss formulas (core_private.cpp:92-115 vs :117-123) agree to rounding
    assert rel_err(o.loss3(fid, 0.03, m, k), f0) < 1e-12


def test_presync_grid_sizes_and_delay_bits(oracle_loader):
    """pin 3: `for (d = c - r; d < c + r; d += step)` yields 200 / 2000 points, not 201 / 2001"""
    d = oracle_loader.presync_delays(0.0, 0.002, 0.2)
    assert len(d) == 200
    assert d[0] == -0.2 and d[-1] == pytest.approx(0.198, abs=1e-12) and d[-1] != 0.198
    x = -0.2
    for i in range(200):
        assert d[i] == x
        x += 0.002
    assert len(oracle_loader.presync_delays(0.0, 0.001, 1.0)) == 2000
    o = oracle_loader.OracleProblem().load(workload("tiny"))
    dd, _ = o.DebugPreSync(0.01, 5, 8, 0.2, 201)
    assert dd[0] == 0.01 - 0.2 and dd[-1] == 0.01 - 0.2 + 2 * 0.2 * 200 / 200 and len(dd) == 201


def test_variable_rate_ingest_known_answers(oracle_loader):
    """pin 6: uniform 5000 us stamps from a multiple of 5000 -> 200 Hz, n-1 samples,
    sample 0 copied, last input dropped; 197.3 Hz -> 200; 449 Hz -> 450 with the 2222/4444/6666 grid"""
    rng = np.random.default_rng(4)
    n = 400
    aa = np.cumsum(rng.normal(scale=0.01, size=(n, 3)), axis=0)
    q = Rotation.from_rotvec(aa).as_quat()[:, [3, 0, 1, 2]]
    ts = 1_000_000 + 5000 * np.arange(n, dtype=np.int64)
    o = oracle_loader.OracleProblem()
    o.SetGyroQuaternions(ts, q, n)
    rq, sr, q0 = o.resampled()
    assert sr == 200.0 and q0 == 1.0 and rq.shape[0] == n - 1
    assert np.array_equal(rq[0], q[0])
    sign = np.sign(np.sum(rq[1:] * q[1:-1], axis=1))[:, None]
    assert np.max(np.abs(rq[1:] * sign - q[1:-1])) < 1e-12
    # unaligned start: the grid starts before t0 (floor division) and holds q[0]
    ts2 = ts + 1234
    o.SetGyroQuaternions(ts2, q, n)
    rq2, sr2, q02 = o.resampled()
    assert sr2 == 200.0 and q02 == 1.0 and np.array_equal(rq2[0], q[0])
    # rate rounding to the nearest 50 Hz
    for hz, want in ((197.3, 200.0), (449.0, 450.0), (1000.4, 1000.0)):
        t = (1e6 * np.arange(n) / hz).astype(np.int64) + 7_000_000
        o.SetGyroQuaternions(t, q, n)
        assert o.resampled()[1] == want
    t = (1e6 * np.arange(n) / 449.0).astype(np.int64)
    o.SetGyroQuaternions(t, q, n)
    assert o.resampled()[2] == 0.0
    # out-of-order timestamps: the reference's panic message
    bad = ts.copy()
    bad[10], bad[11] = bad[11], bad[10]
    with pytest.raises(oracle_loader.OracleError) as e:
        o.SetGyroQuaternions(bad, q, n)
    assert "timestamps out of order at pos 11" in e.value.message


def test_rng_draws_are_pinned(oracle_loader):
    """golden draws of the counter-based RNG: any change breaks parity silently"""
    L = oracle_loader.lib()
    got = [L.orc_rng_index(100, 1, 0, d, 3900, it, k, 200) for d in (0, 7) for it in (0, 19) for k in (0, 1)]
    import json, os
    path = os.path.join(os.path.dirname(__file__), "golden", "rng_draws.json")
    with open(path) as f:
        want = json.load(f)["draws"]
    assert got == want
    assert all(0 <= g < 200 for g in got)


def test_known_delay_recovered(oracle_loader, w_small):
    """pin 7: synthetic scene with known delay: PreSync within a step, Sync within 1 ms"""
    w = w_small
    o = oracle_loader.OracleProblem(threads=8).load(w)
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[0]) + 60
    cost, d = o.PreSync(0.0, fb, fe, w.presync_step, w.presync_radius)
    assert abs(d - 0.037) <= 1.5 * w.presync_step
    for _ in range(2):
        cost, d = o.Sync(d, fb, fe, 0.0, w.presync_radius)
    assert abs(d - 0.037) < 1e-3


def test_double_double_sum_is_order_independent(oracle_loader):
    L = oracle_loader.lib()
    rng = np.random.default_rng(9)
    x = rng.uniform(0, 5, 4000) * 10.0 ** rng.integers(-8, 3, 4000)
    a = L.orc_ddsum(oracle_loader._dp(np.ascontiguousarray(x)), x.size)
    xs = np.ascontiguousarray(rng.permutation(x))
    b = L.orc_ddsum(oracle_loader._dp(xs), xs.size)
    import math
    assert a == b == math.fsum(x)


def test_oracle_reproduces_committed_golden_curve(oracle_loader, w_tiny):
    """the committed fixture (tests/golden/make_golden.py) pins the oracle against drift"""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tiny_curve.json")))
    w = w_tiny
    o = oracle_loader.OracleProblem(threads=2, seed=100).load(w)
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    delays = oracle_loader.presync_delays(0.0, w.presync_step, w.presync_radius)
    assert [float(d).hex() for d in delays] == g["delays_hex"]
    costs = o.presync_grid(fb, fe, delays, call_no=0)
    want = np.array([float.fromhex(h) for h in g["costs_hex"]])
    assert rel_err(costs, want) <= 1e-12
    o.set_rng(100, 1)
    sc, sd = o.Sync(0.035, fb, fe - 1, 0.0, 0.2)
    assert rel_err(sd, float.fromhex(g["sync_delay_hex"])) <= 1e-9
    assert rel_err(sc, float.fromhex(g["sync_cost_hex"])) <= 1e-9


def test_simplified_loss_mode_recovers_the_delay(oracle_loader):
    """the thesis' simplified (no-translation) variant as the oracle defines it: the loss curve has its
    minimum near the true delay on a scene with little translation, and Sync refines it"""
    import importlib
    synth = importlib.import_module("rs-sync_b200.synth")
    w = synth.make_workload("small")
    o = oracle_loader.OracleProblem(threads=4, seed=100).load(w)
    o.set_loss_mode(True)
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    d, c = o.DebugPreSync(0.0, fb, fe, 0.1, 41)
    assert abs(d[int(np.argmin(c))] - w.true_delay[0]) <= 0.005
    cost, delay = o.Sync(float(d[int(np.argmin(c))]), fb, fb + 40, 0.0, 0.2)
    assert abs(delay - w.true_delay[0]) < 3e-3 and cost > 0
    o.set_loss_mode(False)
    assert not np.allclose(o.DebugPreSync(0.0, fb, fe, 0.1, 41)[1], c)


def _ulps(a, b):
    a, b = np.asarray(a), np.asarray(b)
    spacing = np.spacing(np.abs(b))
    return np.max(np.abs(a - b) / spacing)


def test_spec_trig_is_pinned_and_accurate(rsb, oracle_loader):
    """sin / cos / acos of the arithmetic contract (the gyro ingest's only transcendental functions):
    the engine's host code and the oracle's restatement give the same bits, within 1 ulp of libm on
    the ranges the path uses, and the special cases are the reference's (acos of a dot product an ulp
    above 1 is NaN, quat.cpp:61 / core_private.cpp:180)"""
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.uniform(-3.3, 3.3, 200000), rng.uniform(0, 0.2, 50000), rng.uniform(-1e4, 1e4, 50000),
                         np.array([0.0, -0.0, np.pi / 4, np.pi / 2, np.pi, 1e-300, 5e-324])])
    for which, ref in (("sin", np.sin), ("cos", np.cos)):
        a, b = rsb.probe_spec_trig(xs, which), oracle_loader.spec_trig(xs, which)
        assert np.array_equal(a, b)
        ok = np.abs(ref(xs)) > 1e-3  # (ulps of a result near a zero of the function measure the reduction)
        assert _ulps(a[ok], ref(xs)[ok]) <= 1.0
    ys = np.concatenate([rng.uniform(-1, 1, 200000), 1 - rng.uniform(0, 1e-6, 50000) ** 2, np.array([0.0, 0.5, -0.5, 1.0, -1.0])])
    a, b = rsb.probe_spec_trig(ys, "acos"), oracle_loader.spec_trig(ys, "acos")
    assert np.array_equal(a, b)
    assert _ulps(a, np.arccos(ys)) <= 1.0
    bad = rsb.probe_spec_trig(np.array([1.0000000000000002, -1.5, np.nan, np.inf]), "acos")
    assert np.all(np.isnan(bad))
    assert np.all(np.isnan(rsb.probe_spec_trig(np.array([np.inf, -np.inf, np.nan, 1e300]), "sin")))
