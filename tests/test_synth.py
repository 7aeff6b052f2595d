"""The synthetic workload generator is deterministic and shaped like BASELINE.json's configs."""
import numpy as np

from conftest import workload


def test_deterministic_and_frame_local(synth_mod):
    a = synth_mod.make_workload("tiny")
    b = synth_mod.make_workload("tiny")
    assert np.array_equal(a.rays_b, b.rays_b) and np.array_equal(a.quats, b.quats)
    # a frame's rays do not depend on which range was asked for
    c = synth_mod.make_workload("tiny", frames=6, first_frame=8)
    i = int(np.where(a.frame_ids == 9)[0][0])
    j = int(np.where(c.frame_ids == 9)[0][0])
    assert np.allclose(a.rays_a[i], c.rays_a[j]) and np.allclose(a.ts_a[i], c.ts_a[j])


def test_shapes_and_units(synth_mod):
    w = workload("tiny")
    assert w.ts_a.shape == (12, 40) and w.rays_a.shape == (12, 40, 3)
    assert np.allclose(np.linalg.norm(w.rays_a, axis=-1), 1.0) and np.allclose(np.linalg.norm(w.rays_b, axis=-1), 1.0)
    assert np.allclose(np.linalg.norm(w.quats, axis=-1), 1.0)
    # ts_a inside the frame's readout window, ts_b one frame later (core_testcode.cpp:144-145)
    t0 = w.frame_ids[:, None] / w.fps
    assert np.all(w.ts_a >= t0) and np.all(w.ts_a <= t0 + synth_mod.READOUT)
    assert np.all(w.ts_b - w.ts_a > 0.5 / w.fps)
    # the gyro track covers every timestamp +- (radius + true delay)
    tend = w.gyro_t0 + (w.quats.shape[0] - 1) / w.gyro_rate
    assert w.gyro_t0 < w.ts_a.min() - w.presync_radius - 0.05 and tend > w.ts_b.max() + w.presync_radius + 0.05


def test_undistort_roundtrip(synth_mod):
    rng = np.random.default_rng(0)
    px = rng.uniform(100, synth_mod.WIDTH - 100, 500)
    py = rng.uniform(100, synth_mod.HEIGHT - 100, 500)
    qx, qy = synth_mod.ray_to_pixel(synth_mod.pixel_to_ray(px, py))
    assert np.max(np.abs(qx - px)) < 1e-6 and np.max(np.abs(qy - py)) < 1e-6


def test_syncpoints_auto_format(synth_mod):
    """`auto` syncpoints of core_testcode.cpp:270-273: C2 gives 27"""
    w = synth_mod.make_workload("C2", frames=3300, rays=2)
    assert len(w.syncpoints()) == 27 and w.syncpoints()[0] == 3900
