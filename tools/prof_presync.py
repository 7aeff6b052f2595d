"""Profiling driver: one workload, a few PreSync grid launches (for `ncu -k regex:presync_kernel`).
usage: python tools/prof_presync.py [workload] [n_launches]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("rs-sync_b200")
synth = importlib.import_module("rs-sync_b200.synth")
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = synth.make_workload(name)
p = pkg.SyncProblem(seed=100).load(w, bulk=True)
p.flush()  # inputs resident: every grid call below is ONE launch (a grid that follows a bulk ingest
           # directly is launched chunk by chunk behind the upload)
p.set_kernel_timing(True)
fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
delays = np.linspace(-w.presync_radius, w.presync_radius, 201)
for i in range(reps):
    c = p.presync_grid(fb, fe, delays, stream=2, call_no=i)
    st = p.stats()
    print(f"launch {i}: kernel {st['last_grid_kernel_ms']:.3f} ms, argmin {delays[int(np.argmin(c))]:.4f}, "
          f"exact-estimator tasks {st['last_grid_exact_tasks']} of {st['last_grid_tasks']}")
