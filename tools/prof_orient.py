"""Timing of the orientation search (C5 shape) for 1, 8 and 48 variants, and of a variable-rate
SetGyroQuaternions.  usage: python tools/prof_orient.py"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("rs-sync_b200")
synth = importlib.import_module("rs-sync_b200.synth")
w = synth.make_workload("C2")
p = pkg.SyncProblem(seed=100).load(w, bulk=True)
p.flush()
ts = w.gyro_t0 + np.arange(w.quats.shape[0]) / w.gyro_rate
fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
for n in (1, 1, 8, 48, 48):
    t = time.perf_counter()
    c, d = p.orientation_search(ts, w.omega, synth.ORIENTATIONS[:n], 0.0, fb, fe, w.presync_step, w.presync_radius)
    print(f"{n:2d} variants: {(time.perf_counter() - t) * 1e3:8.2f} ms")
# the same with no frames in range: everything but the grids
for n in (1, 48):
    t = time.perf_counter()
    p.orientation_search(ts, w.omega, synth.ORIENTATIONS[:n], 0.0, 10 ** 7, 10 ** 7 + 1, w.presync_step, w.presync_radius)
    print(f"{n:2d} variants, no frames (gyro pipeline only): {(time.perf_counter() - t) * 1e3:8.2f} ms")
tsu = (ts * 1e6).astype(np.int64)
for _ in range(3):
    t = time.perf_counter()
    p.SetGyroQuaternions(tsu, w.quats, len(tsu))
    t1 = time.perf_counter()
    p.flush()
    print(f"variable-rate SetGyroQuaternions: call {(t1 - t) * 1e3:.2f} ms, + flush {(time.perf_counter() - t1) * 1e3:.2f} ms")
