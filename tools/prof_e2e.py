"""Wall-clock breakdown of the end-to-end path (host buffers in, loss curve out).
usage: python tools/prof_e2e.py [workload]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("rs-sync_b200")
synth = importlib.import_module("rs-sync_b200.synth")
w = synth.make_workload(sys.argv[1] if len(sys.argv) > 1 else "C2")
p = pkg.SyncProblem(seed=100)
counts = np.full(w.n_frames, w.n_rays)
fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
delays = np.linspace(-w.presync_radius, w.presync_radius, 201)
FLUSH = "noflush" not in sys.argv  # explicit flush between ingest and grid (separates their times)
for rep in range(4):
    t = [time.perf_counter()]
    p.SetGyroQuaternions(w.quats, w.quats.shape[0], w.gyro_rate, w.gyro_t0); t.append(time.perf_counter())
    p.set_track_batch(w.frame_ids, counts, w.ts_a, w.ts_b, w.rays_a, w.rays_b); t.append(time.perf_counter())
    if FLUSH: p.flush()
    t.append(time.perf_counter())
    c = p.presync_grid(fb, fe, delays, stream=2, call_no=rep); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"rep {rep}: gyro {d[0]:.2f} ms, tracks {d[1]:.2f} ms, flush(H2D) {d[2]:.2f} ms, grid {d[3]:.2f} ms, total {sum(d):.2f} ms")

# the same from tracked pixels (rssync_set_track_pixels): no host sort / transpose, half the bytes
p2 = pkg.SyncProblem(seed=100)
ta, tb = w.frame_ids / w.fps, (w.frame_ids + 1) / w.fps
for rep in range(4):
    t = [time.perf_counter()]
    p2.SetGyroQuaternions(w.quats, w.quats.shape[0], w.gyro_rate, w.gyro_t0); t.append(time.perf_counter())
    p2.set_track_pixels(w.frame_ids, counts, ta, tb, w.px_a, w.px_b, synth.LENS, synth.HEIGHT); t.append(time.perf_counter())
    if FLUSH: p2.flush()
    t.append(time.perf_counter())
    c2 = p2.presync_grid(fb, fe, delays, stream=2, call_no=rep); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"pixels rep {rep}: gyro {d[0]:.2f} ms, tracks {d[1]:.2f} ms, flush {d[2]:.2f} ms, grid {d[3]:.2f} ms, total {sum(d):.2f} ms; "
          f"max rel diff of the curve vs the ray path {np.max(np.abs(c2 - c) / c):.2e}")
