import importlib, time, sys, numpy as np
sys.path.insert(0,'.')
pkg=importlib.import_module("rs-sync_b200"); synth=importlib.import_module("rs-sync_b200.synth")
w=synth.make_workload("C2")
p=pkg.SyncProblem(seed=100).load(w,bulk=True); p.flush()
sps=w.syncpoints(); win=w.sync_window
for rep in range(2):
    p.set_rng(100,0)
    t0=time.perf_counter()
    d=np.array([p.PreSync(0.0,pos,pos+win,w.presync_step,0.2)[1] for pos in sps])
    t1=time.perf_counter()
    fbs=np.array(sps,dtype=np.int64)
    its=[]
    for i in range(4):
        ta=time.perf_counter()
        _,d=p.sync_batch(d,fbs,fbs+win,0.0,0.2)
        st=p.stats(); its.append((st["sync_outer_iters"], st["sync_lbfgs_evals"], time.perf_counter()-ta))
    t2=time.perf_counter()
    print("presync x27: %.1f ms, 4x sync_batch: %.1f ms"%((t1-t0)*1e3,(t2-t1)*1e3), its)
# single sync call timing
p.set_rng(100,0)
t=time.perf_counter(); r=p.Sync(0.037,sps[0],sps[0]+win,0.0,0.2); print("single Sync %.2f ms"%((time.perf_counter()-t)*1e3), r, p.stats()["sync_outer_iters"])
