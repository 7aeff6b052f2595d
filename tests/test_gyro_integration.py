"""Gyro integration (optdata_fill_gyro, core_testcode.cpp:37-53) — the caller-side step before
SetGyroQuaternions and the per-variant work of the orientation search (core_testcode.cpp:184-233).
Host code on both sides: the product's C ABI function against the oracle's restatement and
against analytic answers."""
import numpy as np
import pytest


def test_constant_rate_about_one_axis(rsb, oracle_loader):
    n, rate, w = 2000, 400.0, 1.7
    ts = 12.5 + np.arange(n) / rate
    gyro = np.zeros((n, 3))
    gyro[:, 2] = w
    for q in (rsb.integrate_gyro(ts, gyro), oracle_loader.integrate_gyro(ts, gyro)):
        theta = w * (ts - ts[0])
        want = np.stack([np.cos(theta / 2), 0 * theta, 0 * theta, np.sin(theta / 2)], axis=1)
        assert np.max(np.abs(q - want)) < 1e-12
        assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-15)


def test_product_equals_oracle_and_numpy_restatement(rsb, oracle_loader, synth_mod, w_tiny):
    w = w_tiny
    rng = np.random.default_rng(2)
    ts = w.gyro_t0 + np.arange(w.quats.shape[0]) / w.gyro_rate + rng.uniform(-2e-4, 2e-4, w.quats.shape[0])
    ts = np.sort(ts)
    for orient in [None] + synth_mod.ORIENTATIONS:
        a = rsb.integrate_gyro(ts, w.omega, orient)
        b = oracle_loader.integrate_gyro(ts, w.omega, orient)
        assert np.array_equal(a, b), orient
    om = synth_mod.orient_omega(w.omega, "zXy")
    c = synth_mod.integrate_gyro(om, np.concatenate([[0.0], np.diff(ts)]))
    assert np.max(np.abs(rsb.integrate_gyro(ts, w.omega, "zXy") - c)) < 1e-13


def test_zero_rate_and_bad_orientation(rsb, oracle_loader):
    ts = np.arange(5) * 0.01
    q = rsb.integrate_gyro(ts, np.zeros((5, 3)))  # theta^2 = 0 branch of quat_from_aa (quat.cpp:13-16)
    assert np.array_equal(q, np.tile([1.0, 0, 0, 0], (5, 1)))
    for bad in ("XY", "XYZW", "XYQ"):
        with pytest.raises(rsb.RsSyncError):
            rsb.integrate_gyro(ts, np.zeros((5, 3)), bad)
        with pytest.raises(oracle_loader.OracleError):
            oracle_loader.integrate_gyro(ts, np.zeros((5, 3)), bad)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 16, 17, 18, 19, 20, 33, 64, 257, 4097, 57503])
def test_host_spline_elimination_matches_the_oracle_bit_for_bit(rsb, oracle_loader, n):
    """SetGyroQuaternions' host half (host_ingest.cpp: shared elimination factors, the fixed-point
    shortcut of the sweeps) against the oracle's per-component restatement of minispline.cpp; the
    device half (c = rhs / diag, b, d; spline_finish_kernel) is restated here with numpy's IEEE
    operations, the GPU test compares the finished records themselves"""
    rng = np.random.default_rng(n)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    rhs, diag = rsb.probe_spline_system(q)
    c = rhs / diag[:, None]
    rec = np.empty((n, 16))
    rec[:, 0:4] = q
    rec[:, 8:12] = c
    third = 1.0 / 3.0
    rec[:-1, 12:16] = third * (c[1:] - c[:-1])                                  # minispline.cpp:40
    rec[:-1, 4:8] = (q[1:] - q[:-1]) - third * (2.0 * c[:-1] + c[1:])          # :41
    rec[-1, 12:16] = 0.0                                                         # :43
    rec[-1, 4:8] = (3.0 * rec[-2, 12:16] + 2.0 * c[-2]) + rec[-2, 4:8]          # :44
    o = oracle_loader.OracleProblem()
    o.SetGyroQuaternions(q, n, 1000.0, 0.0)
    assert np.array_equal(rec, o.spline())
