"""Runs the BASELINE.json configurations other than the bench headline on one GPU (or, under
torchrun, sharded) and prints one JSON line per configuration.
usage: python tools/run_configs.py [C1] [C3] [C4] [C5] [--orients N]"""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

pkg = importlib.import_module("rs-sync_b200")
synth = importlib.import_module("rs-sync_b200.synth")
driver = importlib.import_module("rs-sync_b200.driver")
sharded = importlib.import_module("rs-sync_b200.sharded")

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
dev = "cpu"
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    os.environ.setdefault("RSSYNC_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    warm = torch.zeros(1, device=dev)
    dist.all_reduce(warm)  # NCCL builds its communicator on the first collective: keep that out of the timings
    torch.cuda.synchronize()


def emit(d):
    if rank == 0:
        print(json.dumps(d), flush=True)


which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["C1", "C3", "C4", "C5"]
n_orient = int(sys.argv[sys.argv.index("--orients") + 1]) if "--orients" in sys.argv else 48

if "C1" in which:
    w = synth.make_workload("C1")
    p = pkg.SyncProblem(seed=100).load(w, bulk=True)
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    p.PreSync(0.0, fb, fe, w.presync_step, w.presync_radius)
    t = time.perf_counter()
    c, d = p.PreSync(0.0, fb, fe, w.presync_step, w.presync_radius)
    t1 = time.perf_counter()
    cn = p.call_counter()
    p.Sync(d, fb, fb + 60, 0.0, w.presync_radius)  # warm-up: streams, graph, first use of the kernels
    p.set_rng(100, cn)
    t2 = time.perf_counter()
    cs, ds = p.Sync(d, fb, fb + 60, 0.0, w.presync_radius)
    t3 = time.perf_counter()
    emit({"config": "C1: 300 frames x 100 rays, PreSync radius 200 ms step 2 ms (200 offsets) + one Sync",
          "presync_ms": (t1 - t) * 1e3, "presync_delay": d, "presync_cells_per_s": 200 * 300 * 100 / (t1 - t),
          "sync_ms": (t3 - t2) * 1e3, "sync_cold_ms": (t2 - t1) * 1e3, "sync_delay": ds,
          "true_delay": float(w.true_delay[0])})

if "C3" in which:
    t0 = time.perf_counter()
    w = synth.make_workload("C3")
    t_gen = time.perf_counter() - t0
    p = pkg.SyncProblem(seed=100).load(w, bulk=True)
    p.set_kernel_timing(True)
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    delays = np.linspace(-1.0, 1.0, 2001)
    p.flush()
    lo, hi = sharded.shard_range(len(delays), rank, world)
    p.presync_grid(fb, fe, delays[lo:lo + 8], stream=2, call_no=0, offset_index_base=lo)  # warm-up
    t = time.perf_counter()
    curve = sharded.presync_grid_sharded(p, fb, fe, delays, stream=2, call_no=1, rank=rank, world=world, device=dev)
    dt = time.perf_counter() - t
    st = p.stats()
    emit({"config": f"C3: 10000 frames x 500 rays x 2001 offsets (radius 1 s, step 1 ms), offsets sharded x{world}",
          "seconds": dt, "cells_per_s": 2001 * 10000 * 500 / dt, "kernel_ms_rank0": st["last_grid_kernel_ms"],
          "argmin_delay": float(delays[int(np.argmin(curve))]), "true_delay": float(w.true_delay[0]),
          "exact_estimator_tasks_rank0": int(st["last_grid_exact_tasks"]), "tasks_rank0": int(st["last_grid_tasks"]),
          "synth_seconds": t_gen})
    del p, w

if "C4" in which:
    t0 = time.perf_counter()
    w = synth.make_workload("C4s")
    t_gen = time.perf_counter() - t0
    p = pkg.SyncProblem(seed=100).load(w, bulk=True)
    p.flush()
    cfg = driver.default_config(w)
    res = None
    for rep in range(2):
        p.set_rng(100, 0)
        t = time.perf_counter()
        res = driver.run(p, cfg, mode="batched", rank=rank, world=world, device=dev, debug_csv=None,
                         presync_delays=pkg.presync_delays)
        dt = time.perf_counter() - t
    err = [abs(d / 1000.0 - w.true_delay_at(s)) for s, d in zip(res["syncpoints"], res["delay_ms"])]
    emit({"config": f"C4: 30 min 60 fps trace, {len(res['syncpoints'])} syncpoints (window 60, every 1000 frames), "
                    f"PreSync + 4 x Sync each, syncpoints sharded x{world}",
          "seconds": dt, "syncpoints_per_s": len(res["syncpoints"]) / dt, "rmse_vs_linear_fit_ms": res["rmse_ms"],
          "mean_abs_delay_error_ms": float(np.mean(err)) * 1e3, "first_last_delay_ms": [float(res["delay_ms"][0]), float(res["delay_ms"][-1])],
          "gyro_samples": int(w.quats.shape[0]), "frames": w.n_frames, "synth_seconds": t_gen})
    del p, w

if "C5" in which:
    w = synth.make_workload("C2")
    p = pkg.SyncProblem(seed=100).load(w, bulk=True)
    p.flush()
    ts = w.gyro_t0 + np.arange(w.quats.shape[0]) / w.gyro_rate
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    orients = synth.ORIENTATIONS[:n_orient]
    search = lambda pr, ors: pr.orientation_search(ts, w.omega, ors, 0.0, fb, fe, w.presync_step, w.presync_radius)
    search(p, orients[:1])  # warm-up
    t = time.perf_counter()
    batch = lambda pr, ors, cn: pr.orientation_search(ts, w.omega, ors, 0.0, fb, fe, w.presync_step, w.presync_radius,
                                                      call_nos=cn)
    cost, delay = sharded.orientation_search_sharded(p, search, orients, seed=100, call_no_base=0, rank=rank,
                                                     world=world, device=dev, batch_fn=batch)
    dt = time.perf_counter() - t
    order = np.argsort(cost)
    emit({"config": f"C5: {len(orients)} gyro_orientation variants x PreSync on C2 (3300 x 200 x 200 offsets), variants sharded x{world}",
          "seconds": dt, "cells_per_s": len(orients) * 200 * 3300 * 200 / dt,
          "top5": [[orients[i], float(cost[i]), float(delay[i])] for i in order[:5]], "true_orientation": "XYZ"})

if world > 1:
    dist.destroy_process_group()
