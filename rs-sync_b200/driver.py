"""Syncpoint driver: the main() of the reference's demo program (core_testcode.cpp:235-318) without
the video front end.  Same JSON config keys (README.md:15-44), same outputs:

  input.initial_guess (ms), input.use_simple_presync, input.simple_presync_radius (ms),
  input.simple_presync_step (ms), input.frame_range [begin, end],
  params.sync_window, params.syncpoints_format ("auto" | "array"), params.syncpoint_distance,
  params.syncpoints_array, output.csv_path

`debug.csv` gets the 200-point DebugPreSync curve of the first window (:283-299) and the result CSV
one `pos,1000*delay` row per syncpoint (:315).  Video / gyro / lens paths are replaced by a
problem the caller has already filled (SetGyroQuaternions + SetTrackResult), e.g. from
rs-sync_b200.synth.  `rmse_vs_linear_fit` is the accuracy figure of the thesis (python/plot_sync.py:19,44).

Two execution modes give identical numbers: `sequential` issues the reference's call sequence
(DebugPreSync, then per syncpoint PreSync + 4 x Sync); `batched` evaluates the same calls with
explicit RNG call numbers, the Sync passes of all syncpoints advanced side by side on the device
(and, under torch.distributed, syncpoints sharded over ranks).
"""
from __future__ import annotations

import json
import math
import sys

import numpy as np

N_SYNC_PASSES = 4          # core_testcode.cpp:314
DEBUG_PLOT_SIZE = 200      # core_testcode.cpp:285


def syncpoint_list(config):
    """core_testcode.cpp:268-280"""
    inp, par = config["input"], config["params"]
    fmt = par["syncpoints_format"]
    if fmt == "auto":
        f0, f1 = int(inp["frame_range"][0]), int(inp["frame_range"][1])
        return list(range(f0, f1 - int(par["sync_window"]), int(par["syncpoint_distance"])))
    if fmt == "array":
        return [int(p) for p in par["syncpoints_array"]]
    raise ValueError(f"syncpoints_format must be 'auto' or 'array', got {fmt!r}")


def _fmt(x):
    """operator<< of a double with the default precision (6 significant digits)"""
    return "%g" % x


def rmse_vs_linear_fit(pos, delay_ms):
    """plot_sync.py:19,44: standard deviation of the residual of a least-squares line"""
    pos = np.asarray(pos, dtype=np.float64)
    d = np.asarray(delay_ms, dtype=np.float64)
    if len(pos) < 3:
        return float("nan")
    slope, intercept = np.polyfit(pos, d, 1)
    return float(np.std(intercept + slope * pos - d))


def run(problem, config, *, mode="batched", rank=0, world=1, device="cpu", debug_csv="debug.csv", presync_delays=None):
    """Runs the syncpoint loop on `problem`; returns {"syncpoints", "delay_ms", "cost", "rmse_ms"}.
    Rank 0 writes the CSV files.  `presync_delays(initial, step, radius)` must be the engine's delay
    grid function when mode == "batched"."""
    inp, par, out = config["input"], config["params"], config.get("output", {})
    window = int(par["sync_window"])
    initial = float(inp["initial_guess"]) / 1000.0
    use_presync = bool(inp.get("use_simple_presync", False))
    radius = float(inp["simple_presync_radius"]) / 1000.0 if use_presync else math.inf
    step = float(inp.get("simple_presync_step", 0)) / 1000.0
    sps = syncpoint_list(config)
    n = len(sps)
    base = problem.call_counter() if hasattr(problem, "call_counter") else 0

    if debug_csv and "simple_presync_radius" in inp:  # :283-299
        f0 = int(inp["frame_range"][0])
        dd, cc = problem.DebugPreSync(initial, f0, f0 + window, float(inp["simple_presync_radius"]) / 1000.0,
                                      DEBUG_PLOT_SIZE)
        if rank == 0:
            with open(debug_csv, "w") as f:
                for a, b in zip(dd, cc):
                    f.write(f"{_fmt(a)},{_fmt(b)}\n")
        base += 1

    per_sp = (1 if use_presync else 0) + N_SYNC_PASSES  # API calls per syncpoint
    delays = np.full(n, initial)
    costs = np.zeros(n)
    if mode == "sequential":
        for i, pos in enumerate(sps):  # :303-316
            d = initial
            if use_presync:
                d = problem.PreSync(d, pos, pos + window, step, radius)[1]
            for _ in range(N_SYNC_PASSES):
                costs[i], d = problem.Sync(d, pos, pos + window, initial, radius)
            delays[i] = d
    else:
        from . import sharded
        mine = list(range(rank, n, world))
        if use_presync:
            if hasattr(problem, "presync_windows") and mine:  # all of this rank's windows in one grid launch
                idx = np.asarray(mine)
                wb = np.asarray(sps, dtype=np.int64)[idx]
                _, d = problem.presync_windows(initial, wb, wb + window, step, radius,
                                               call_nos=(base + per_sp * idx).astype(np.uint64))
                delays[idx] = d
            else:
                grid = np.asarray(presync_delays(initial, step, radius))
                for i in mine:
                    curve = problem.presync_grid(sps[i], sps[i] + window, grid, stream=1, call_no=base + per_sp * i)
                    delays[i] = grid[sharded.argmin_cost_delay(curve, grid)]
            if world > 1:
                import torch.distributed as dist
                loc = np.array([delays[i] for i in mine])
                allv = sharded._gather_variable(loc, world, device, dist)
                order = np.concatenate([np.arange(r, n, world) for r in range(world)])
                delays[order] = allv
        fbs = np.asarray(sps, dtype=np.int64)
        for k in range(N_SYNC_PASSES):
            first = base + (1 if use_presync else 0) + k
            call_nos = first + per_sp * np.arange(n)
            costs, delays = _sync_pass(problem, sharded, delays, fbs, fbs + window, initial, radius, call_nos,
                                       rank, world, device)
        if hasattr(problem, "set_rng") and hasattr(problem, "seed"):
            problem.set_rng(problem.seed, base + per_sp * n)

    delay_ms = 1000.0 * delays
    if rank == 0 and out.get("csv_path"):
        with open(out["csv_path"], "w") as f:
            for pos, d in zip(sps, delay_ms):
                f.write(f"{pos},{_fmt(d)}\n")
    return {"syncpoints": sps, "delay_ms": delay_ms, "cost": costs, "rmse_ms": rmse_vs_linear_fit(sps, delay_ms)}


def _sync_pass(problem, sharded, initial, fb, fe, center, radius, call_nos, rank, world, device):
    n = len(initial)
    mine = np.arange(rank, n, world)
    if mine.size:
        c, d = problem.sync_batch(np.asarray(initial)[mine], fb[mine], fe[mine], center, radius,
                                  call_nos=call_nos[mine].astype(np.uint64))
    else:
        c, d = np.empty(0), np.empty(0)
    if world == 1:
        return c, d
    import torch.distributed as dist
    allc = sharded._gather_variable(c, world, device, dist)
    alld = sharded._gather_variable(d, world, device, dist)
    order = np.concatenate([np.arange(r, n, world) for r in range(world)])
    outc, outd = np.empty(n), np.empty(n)
    outc[order], outd[order] = allc, alld
    return outc, outd


def default_config(w, csv_path=None, use_presync=True):
    """config for a synthetic workload `w` (rs-sync_b200.synth.Workload), reference keys"""
    f0, f1 = w.meta.get("span", (int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1))
    return {
        "input": {"frame_range": [f0, f1], "initial_guess": 0.0, "use_simple_presync": use_presync,
                  "simple_presync_radius": 1000.0 * w.presync_radius, "simple_presync_step": 1000.0 * w.presync_step},
        "params": {"sync_window": w.sync_window, "syncpoints_format": "auto",
                   "syncpoint_distance": w.syncpoint_distance},
        "output": {"csv_path": csv_path},
    }


def main(argv=None):
    """python -m rs-sync_b200.driver config.json [workload]: runs the loop on a synthetic workload"""
    import importlib
    argv = sys.argv[1:] if argv is None else argv
    pkg = importlib.import_module(__package__)
    synth = importlib.import_module(__package__ + ".synth")
    with open(argv[0]) as f:
        config = json.load(f)
    syn = config["input"].get("synthetic", {})
    w = synth.make_workload(argv[1] if len(argv) > 1 else syn.get("workload", "C1"))
    config["input"].setdefault("frame_range", list(w.meta["span"]))
    # several GPUs in this one process: `"devices": [0, 1, ...]` under input.synthetic, or the C++
    # drop-in's RSSYNC_DEVICES=0,1,... (rssync_create_multi; results are the single-GPU results)
    import os
    devices = syn.get("devices")
    if devices is None and os.environ.get("RSSYNC_DEVICES"):
        devices = [int(d) for d in os.environ["RSSYNC_DEVICES"].split(",") if d.strip()]
    prob = pkg.SyncProblem(seed=int(syn.get("seed", 100)), devices=devices or None).load(w, bulk=True)
    res = run(prob, config, presync_delays=pkg.presync_delays)
    print(json.dumps({"syncpoints": len(res["syncpoints"]), "rmse_ms": res["rmse_ms"],
                      "first_delay_ms": float(res["delay_ms"][0]), "last_delay_ms": float(res["delay_ms"][-1])}))


if __name__ == "__main__":
    main()
