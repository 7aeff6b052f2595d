"""in-library multi-GPU problem on a BASELINE-sized workload: load, grid, re-ingest, grid (debug aid)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("rs-sync_b200")
synth = importlib.import_module("rs-sync_b200.synth")
import torch
w = synth.make_workload(sys.argv[1] if len(sys.argv) > 1 else "C2")
n = torch.cuda.device_count()
one = pkg.SyncProblem(seed=100).load(w, bulk=True)
mp = pkg.SyncProblem(seed=100, devices=list(range(n)))
counts = np.full(w.n_frames, w.n_rays)
fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
delays = np.linspace(-0.2, 0.2, 201)
want = one.presync_grid(fb, fe, delays, stream=2, call_no=7)
for rep in range(4):
    t0 = time.perf_counter()
    mp.SetGyroQuaternions(w.quats, w.quats.shape[0], w.gyro_rate, w.gyro_t0)
    mp.set_track_batch(w.frame_ids, counts, w.ts_a, w.ts_b, w.rays_a, w.rays_b)
    print("ingested", flush=True)
    got = mp.presync_grid(fb, fe, delays, stream=2, call_no=7)
    print(f"rep {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms, equal {np.array_equal(got, want)}", flush=True)
