"""Regenerates the committed golden fixtures from the oracle (run from the repo root):
    python tests/golden/make_golden.py
rng_draws.json   pinned draws of the counter-based RNG
tiny_curve.json  oracle PreSync loss curve + Sync result on the `tiny` synthetic workload
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    L = loader.lib()
    draws = [L.orc_rng_index(100, 1, 0, d, 3900, it, k, 200) for d in (0, 7) for it in (0, 19) for k in (0, 1)]
    json.dump({"comment": "rng_index(seed=100, stream=1, call_no=0, offset_idx=d, frame=3900, iter=it, k=k, n=200) "
                          "for d in (0,7), it in (0,19), k in (0,1)", "draws": draws},
              open(os.path.join(HERE, "rng_draws.json"), "w"))
    synth = importlib.import_module("rs-sync_b200.synth")
    w = synth.make_workload("tiny")
    o = loader.OracleProblem(threads=2, seed=100).load(w)
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    delays = loader.presync_delays(0.0, w.presync_step, w.presync_radius)
    costs = o.presync_grid(fb, fe, delays, call_no=0)
    o.set_rng(100, 1)
    sc, sd = o.Sync(0.035, fb, fe - 1, 0.0, 0.2)
    json.dump({"comment": "oracle results on synth.make_workload('tiny'), seed 100: presync_grid(call_no=0) over "
                          "presync_delays(0, step, radius); Sync(0.035, fb, fe-1, 0, 0.2) with call_no=1",
               "delays_hex": [float(d).hex() for d in delays], "costs_hex": [float(c).hex() for c in costs],
               "sync_cost_hex": float(sc).hex(), "sync_delay_hex": float(sd).hex()},
              open(os.path.join(HERE, "tiny_curve.json"), "w"))
    print("wrote golden fixtures; argmin delay", delays[int(np.argmin(costs))], "sync delay", sd)


if __name__ == "__main__":
    main()
