"""The syncpoint driver (rs-sync_b200/driver.py), the no-video restatement of core_testcode's main
(core_testcode.cpp:235-318): config keys, syncpoint lists, CSV format, and that the batched
(lock-step, explicit call numbers) execution reproduces the reference's sequential call sequence.
The compute backend on CPU is the oracle; the GPU engine is checked in test_gpu_parity.py."""
import importlib
import os

import numpy as np
import pytest

from conftest import workload


@pytest.fixture(scope="module")
def driver():
    return importlib.import_module("rs-sync_b200.driver")


def test_syncpoint_lists(driver):
    cfg = {"input": {"frame_range": [3900, 7200]},
           "params": {"sync_window": 60, "syncpoints_format": "auto", "syncpoint_distance": 120}}
    sps = driver.syncpoint_list(cfg)
    assert sps[0] == 3900 and sps[-1] + 60 < 7200 and len(sps) == 27 and sps[1] - sps[0] == 120
    cfg["params"].update(syncpoints_format="array", syncpoints_array=[10, 500, 20])
    assert driver.syncpoint_list(cfg) == [10, 500, 20]
    cfg["params"]["syncpoints_format"] = "bogus"
    with pytest.raises(ValueError):
        driver.syncpoint_list(cfg)


def test_rmse_is_plot_sync_metric(driver):
    import scipy.stats as st
    rng = np.random.default_rng(0)
    pos = np.arange(0, 3000, 120.0)
    d = -45 + 0.001 * pos + rng.normal(0, 0.2, pos.size)
    r = st.linregress(pos, d)  # plot_sync.py:19
    want = np.std(r.intercept + r.slope * pos - d)  # plot_sync.py:44
    assert abs(driver.rmse_vs_linear_fit(pos, d) - want) < 1e-12


def test_batched_equals_sequential_and_csv(driver, oracle_loader, tmp_path):
    w = workload("small")
    cfg = driver.default_config(w, csv_path=str(tmp_path / "out.csv"))
    cfg["params"].update(sync_window=12, syncpoint_distance=20)
    cfg["input"].update(simple_presync_radius=50.0, simple_presync_step=5.0)
    a = oracle_loader.OracleProblem(threads=4, seed=100).load(w)
    b = oracle_loader.OracleProblem(threads=4, seed=100).load(w)
    ra = driver.run(a, cfg, mode="sequential", debug_csv=str(tmp_path / "debug_a.csv"))
    rb = driver.run(b, cfg, mode="batched", debug_csv=str(tmp_path / "debug_b.csv"),
                    presync_delays=oracle_loader.presync_delays)
    assert len(ra["syncpoints"]) == 3
    assert np.array_equal(ra["delay_ms"], rb["delay_ms"]) and np.array_equal(ra["cost"], rb["cost"])
    assert a.call_counter() == b.call_counter() == 1 + 3 * 5
    assert np.max(np.abs(ra["delay_ms"] - 37.0)) < 2.5
    rows = open(tmp_path / "out.csv").read().strip().split("\n")
    assert rows[0] == f"{ra['syncpoints'][0]},{'%g' % ra['delay_ms'][0]}" and len(rows) == 3
    dbg = np.loadtxt(tmp_path / "debug_a.csv", delimiter=",")
    assert dbg.shape == (200, 2) and abs(dbg[0, 0] + 0.05) < 1e-6 and abs(dbg[-1, 0] - 0.05) < 1e-6
    assert open(tmp_path / "debug_a.csv").read() == open(tmp_path / "debug_b.csv").read()


def test_no_presync_uses_infinite_radius(driver, oracle_loader, tmp_path):
    """use_simple_presync false: Sync starts from initial_guess with radius = inf (core_testcode.cpp:307)"""
    w = workload("tiny")
    cfg = driver.default_config(w, use_presync=False)
    cfg["params"].update(sync_window=8, syncpoint_distance=100)
    cfg["input"]["initial_guess"] = 35.0
    o = oracle_loader.OracleProblem(threads=2, seed=3).load(w)
    r = driver.run(o, cfg, mode="sequential", debug_csv=None)
    o2 = oracle_loader.OracleProblem(threads=2, seed=3).load(w)
    d = 0.035
    for _ in range(4):
        d = o2.Sync(d, int(w.frame_ids[0]), int(w.frame_ids[0]) + 8, 0.035, np.inf)[1]
    assert r["delay_ms"][0] == 1000.0 * d


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line on
    stdout with the GPU arm's metric / unit / config keys, impl = reference, a cpu_baseline describing
    the run and an e2e block that moves no bytes.  Shortened sample, smallest workload."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    env = dict(os.environ, RSSYNC_REF_BUDGET="0.5")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "presync_loss_evals_per_s" and d["unit"] == "cells/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "offsets" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
