"""world_size-2 gloo test of the multi-GPU host logic (rs-sync_b200/sharded.py) on CPU.  The
compute backend is the oracle (tests may use it); what is under test is the sharding, RNG keying
by global offset index / call number, the gather and the argmin."""
import importlib
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT, workload


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = importlib.import_module("rs-sync_b200.sharded")
    synth = importlib.import_module("rs-sync_b200.synth")
    from oracle import loader
    w = synth.make_workload("tiny")
    o = loader.OracleProblem(threads=1, seed=100).load(w)
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    delays = loader.presync_delays(0.0, 0.005, 0.05)
    curve = sharded.presync_grid_sharded(o, fb, fe, delays, stream=1, call_no=4, rank=rank, world=world)
    best = sharded.presync_sharded(o, 0.0, fb, fe, 0.005, 0.05, delays, call_no=4, rank=rank, world=world)
    ini = np.array([0.030, 0.034, 0.036])
    fbs = np.array([fb, fb + 2, fb + 4])
    c, d = sharded.sync_sharded(o, ini, fbs, fbs + 6, 0.0, 0.2, call_no_base=20, rank=rank, world=world)
    # orientation search, variant k on rank k % world
    # (a later first frame: the variable-rate ingest needs non-negative microsecond timestamps)
    w2 = synth.make_workload("tiny", first_frame=200)
    ts = w2.gyro_t0 + np.arange(w2.quats.shape[0]) / w2.gyro_rate
    o2 = loader.OracleProblem(threads=1, seed=100).load(w2)
    oc, od = sharded.orientation_search_sharded(
        o2, lambda pr, ors: loader.orientation_search(pr, ts, w2.omega, ors, 0.0, 200, 212, 0.01, 0.05),
        ["XYZ", "yXz", "ZXY"], seed=100, call_no_base=7, rank=rank, world=world)
    # the same through the batch form (one call per rank with explicit call numbers)
    def batch(pr, ors, call_nos):
        cs, ds = [], []
        for o_, cn in zip(ors, call_nos):
            pr.set_rng(100, int(cn))
            c_, d_ = loader.orientation_search(pr, ts, w2.omega, [o_], 0.0, 200, 212, 0.01, 0.05)
            cs.append(c_[0]); ds.append(d_[0])
        return np.array(cs), np.array(ds)
    bc, bd = sharded.orientation_search_sharded(o2, None, ["XYZ", "yXz", "ZXY"], seed=100, call_no_base=7,
                                                rank=rank, world=world, batch_fn=batch)
    assert np.array_equal(bc, oc) and np.array_equal(bd, od)
    if rank == 0:
        q.put((curve, best, c, d, oc, od))
    dist.barrier()
    dist.destroy_process_group()


def test_offset_and_syncpoint_sharding_world2(oracle_loader):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    curve, best, c, d, oc, od = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process answers
    w = workload("tiny")
    o = oracle_loader.OracleProblem(threads=2, seed=100).load(w)
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    delays = oracle_loader.presync_delays(0.0, 0.005, 0.05)
    whole = o.presync_grid(fb, fe, delays, call_no=4)
    assert np.array_equal(curve, whole)
    o.set_rng(100, 4)
    assert best == o.PreSync(0.0, fb, fe, 0.005, 0.05)
    o.set_rng(100, 20)
    seq = [o.Sync(x, fb + 2 * i, fb + 2 * i + 6, 0.0, 0.2) for i, x in enumerate((0.030, 0.034, 0.036))]
    assert np.array_equal(c, [s[0] for s in seq]) and np.array_equal(d, [s[1] for s in seq])
    # orientation search: equals the single-process loop (call numbers 7, 8, 9)
    w2 = workload("tiny", first_frame=200)
    ts = w2.gyro_t0 + np.arange(w2.quats.shape[0]) / w2.gyro_rate
    o = oracle_loader.OracleProblem(threads=2, seed=100).load(w2)
    o.set_rng(100, 7)
    wc, wd = oracle_loader.orientation_search(o, ts, w2.omega, ["XYZ", "yXz", "ZXY"], 0.0, 200, 212, 0.01, 0.05)
    assert np.array_equal(oc, wc) and np.array_equal(od, wd)
    assert int(np.argmin(oc)) == 0  # the true orientation has the lowest cost


def test_shard_range_partitions():
    sharded = importlib.import_module("rs-sync_b200.sharded")
    for n in (0, 1, 7, 200, 201, 2001):
        for world in (1, 2, 3, 8):
            parts = [sharded.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))


def test_replication_pieces_cover_the_arena():
    """sharded._merge_chunks: the pieces of a pipelined replication cover every ray of the arena once
    the chunks still in flight are grouped; what no chunk covers goes first (already on the device)"""
    import importlib
    sh = importlib.import_module("rs-sync_b200.sharded")
    assert sh._merge_chunks([], 4, 100) == [(-1, 0, 100)]
    assert sh._merge_chunks([], 4, 0) == []
    assert sh._merge_chunks([(0, 25), (25, 50), (50, 75), (75, 100)], 4, 100) == [(0, 0, 25), (1, 25, 50), (2, 50, 75), (3, 75, 100)]
    assert sh._merge_chunks([(60, 70), (70, 80), (80, 100)], 2, 100) == [(-1, 0, 60), (0, 60, 70), (2, 70, 100)]
    assert sh._merge_chunks([(10, 20), (40, 50)], 4, 100) == [(-1, 0, 10), (-1, 20, 40), (-1, 50, 100), (0, 10, 20), (1, 40, 50)]
    assert sh._merge_chunks([(i, i + 1) for i in range(0, 60, 2)], 4, 100) == [(29, 0, 100)]  # scattered: one piece, last
    for chunks in ([(0, 8), (8, 16), (16, 24), (24, 32), (32, 40)], [(32, 40), (0, 8)], [(5, 9)]):
        covered = [0] * 40
        for k, lo, hi in sh._merge_chunks(chunks, 3, 40):
            assert -1 <= k < len(chunks)
            for i in range(lo, hi):
                covered[i] += 1
        assert min(covered) >= 1


def test_library_and_python_replication_plans_agree(rsb):
    """the planner inside the library (rssync_create_multi problems, plan_replication in capi.cpp) and
    sharded._merge_chunks state the same rule: same pieces for the same chunks in flight (host-only)"""
    import importlib
    import random
    sh = importlib.import_module("rs-sync_b200.sharded")
    rnd = random.Random(5)
    cases = [([], 0), ([], 1000), ([(0, 250), (250, 500), (500, 750), (750, 1000)], 1000),
             ([(600, 700), (700, 800), (800, 1000)], 1000), ([(100, 200), (400, 500)], 1000),
             ([(i, i + 1) for i in range(0, 60, 2)], 100)]
    for _ in range(200):
        arena = rnd.randrange(1, 5000)
        n = rnd.randrange(0, 12)
        if rnd.random() < 0.5:  # contiguous chunks of one batch appended somewhere in the arena
            cuts = sorted(rnd.sample(range(arena + 1), min(n + 1, arena + 1)))
            chunks = [(a, b) for a, b in zip(cuts, cuts[1:])]
        else:                   # scattered, possibly overlapping ranges in any order
            chunks = []
            for _ in range(n):
                a = rnd.randrange(0, arena)
                chunks.append((a, rnd.randrange(a + 1, arena + 1)))
        cases.append((chunks, arena))
    for chunks, arena in cases:
        for groups in (1, 2, 4):
            assert rsb.probe_replication_plan(chunks, arena, groups) == sh._merge_chunks(chunks, groups, arena), (chunks, arena, groups)
