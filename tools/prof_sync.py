"""Timing of the Sync path on C2's 27 syncpoints (bench.py's `sync` section), with launch counts."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("rs-sync_b200")
synth = importlib.import_module("rs-sync_b200.synth")
w = synth.make_workload(sys.argv[1] if len(sys.argv) > 1 else "C2")
p = pkg.SyncProblem(seed=100).load(w, bulk=True)
p.flush()
sps = w.syncpoints(); win = w.sync_window
fbs = np.array(sps, dtype=np.int64)
for rep in range(3):
    p.set_rng(100, 0)
    l0 = p.stats()["kernel_launches"]
    t0 = time.perf_counter()
    d = p.presync_windows(0.0, fbs, fbs + win, w.presync_step, 0.2)[1]
    t1 = time.perf_counter()
    its = []
    for i in range(4):
        ta = time.perf_counter()
        _, d = p.sync_batch(d, fbs, fbs + win, 0.0, 0.2)
        st = p.stats()
        its.append((st["sync_outer_iters"], round((time.perf_counter() - ta) * 1e3, 2)))
    t2 = time.perf_counter()
    print(f"{len(sps)} syncpoints: presync {1e3*(t1-t0):.1f} ms, 4 x sync_batch {1e3*(t2-t1):.1f} ms {its}, "
          f"launches {p.stats()['kernel_launches']-l0}, {len(sps)/(t2-t0):.0f} syncpoints/s")
