"""Small end-to-end pass over every kernel of the engine, for `compute-sanitizer` (memcheck,
racecheck, synccheck, initcheck): bulk ingest on the device (ingest_rays_kernel), pixel front end
(ingest_pixels_kernel), spline records (spline_finish_kernel), the PreSync grid through the staged
path and the global path (presync_kernel: TMA bulk copies, mbarrier hand-off, the `arrived` counter),
the windows form, Sync (sync_init / sync_motion_fgrad / sync_trials / reduce kernels; the L-BFGS
history in shared memory) and the probes.  Results are compared with the CPU oracle so that a run
under the sanitizer is also a correctness run.
usage: compute-sanitizer --tool racecheck python tools/sanitize_workload.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("rs-sync_b200")
synth = importlib.import_module("rs-sync_b200.synth")
from oracle import loader

w = synth.make_workload("small", frames=66, rays=70)   # >= 64 frames: device-side bulk ingest; 3 slots, ragged tail
g = pkg.SyncProblem(seed=100).load(w, bulk=True)
o = loader.OracleProblem(threads=4, seed=100).load(w)
fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
delays = np.linspace(-0.02, 0.02, 19)
cg = g.presync_grid(fb, fe, delays, stream=2, call_no=1)          # staged path
co = o.presync_grid(fb, fe, delays, stream=2, call_no=1)
assert np.max(np.abs(cg - co) / co) <= 1e-9, "staged grid"
wide = np.array([-3.0, -0.5, 0.0, 0.4, 2.5])                       # window does not fit: global path
cw, cwo = g.presync_grid(fb, fb + 9, wide, call_no=2), o.presync_grid(fb, fb + 9, wide, call_no=2)
assert np.max(np.abs(cw - cwo) / cwo) <= 1e-9, "global-path grid"
fbs = np.array([fb, fb + 20, fb + 40])
g.set_rng(100, 10)
pc, pd = g.presync_windows(0.0, fbs, fbs + 20, 0.004, 0.04)
g.set_rng(100, 20); o.set_rng(100, 20)
sg = g.Sync(0.038, fb, fb + 12, 0.0, 0.2)
so = o.Sync(0.038, fb, fb + 12, 0.0, 0.2)
assert abs(sg[1] - so[1]) <= 1e-9 * abs(so[1]), ("sync", sg, so)
g.set_rng(100, 30)
g.sync_batch(np.array([0.038, 0.036, 0.02]), np.array([fb, fb + 30, 10 ** 6]), np.array([fb + 10, fb + 40, 10 ** 6 + 5]), 0.0, 0.2)
p2 = pkg.SyncProblem(seed=100)
p2.SetGyroQuaternions(w.gyro_timestamps_us(), w.quats, w.quats.shape[0])    # variable-rate ingest
ta, tb = w.frame_ids / w.fps, (w.frame_ids + 1) / w.fps
p2.set_track_pixels(w.frame_ids, np.full(w.n_frames, w.n_rays), ta, tb, w.px_a, w.px_b, synth.LENS, synth.HEIGHT)
p2.presync_grid(fb, fe, delays[:4], stream=2, call_no=1)
g.probe_problem_matrix(fb + 3, 0.03, w.n_rays); g.probe_guess_motion(fb + 3, 0.03, 200, 3, 0, 0)
g.probe_loss(fb + 3, 0.03, np.array([0.1, 0.2, 0.9]), 50.0); g.probe_lbfgs(fb + 3, 0.03, np.array([0.1, 0.2, 0.9]), 50.0)
print("sanitize workload ok; kernel launches", g.stats()["kernel_launches"])
