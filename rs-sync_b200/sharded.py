"""Multi-GPU decomposition of the hot path: one process per GPU, inputs replicated, independent
units sharded, one small gather per call.

* PreSync / DebugPreSync: contiguous offset ranges per rank.  Every frame's reduction stays on
  one GPU and the RNG is keyed by the global offset index, so the gathered curve equals the
  single-GPU curve bit for bit.
* Sync: syncpoint i goes to rank i % world; each rank advances its syncpoints side by side on its
  own device (in independent lanes); (cost, delay) pairs are gathered at the end.

The functions only need a `torch.distributed` process group (NCCL on GPUs, gloo in the CPU tests)
and a problem object with the SyncProblem methods, so the logic is testable without a GPU.
"""
from __future__ import annotations

import numpy as np


def shard_range(n, rank, world):
    """contiguous [lo, hi) slice of n units for `rank`"""
    return n * rank // world, n * (rank + 1) // world


def _gather_variable(local, group_size, device, dist):
    import torch
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(group_size)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    buf = torch.zeros(max(m, 1), dtype=torch.float64, device=device)
    buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(device)
    out = [torch.zeros(max(m, 1), dtype=torch.float64, device=device) for _ in range(group_size)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)])


def presync_grid_sharded(problem, frame_begin, frame_end, delays, *, stream=1, call_no=0, rank=0, world=1,
                         device="cpu"):
    """Loss curve of `delays` with the offsets sharded over `world` ranks; every rank returns the
    full curve."""
    delays = np.ascontiguousarray(delays, dtype=np.float64)
    lo, hi = shard_range(delays.shape[0], rank, world)
    local = problem.presync_grid(frame_begin, frame_end, delays[lo:hi], stream=stream, call_no=call_no,
                                 offset_index_base=lo)
    if world == 1:
        return np.asarray(local)
    import torch.distributed as dist
    return _gather_variable(np.asarray(local), world, device, dist)


def argmin_cost_delay(costs, delays):
    """std::min_element over (cost, delay) pairs, core_private.cpp:89"""
    best = 0
    for i in range(1, len(costs)):
        if costs[i] < costs[best] or (costs[i] == costs[best] and delays[i] < delays[best]):
            best = i
    return best


def presync_sharded(problem, initial_delay, frame_begin, frame_end, search_step, search_radius, delays, *,
                    call_no=0, rank=0, world=1, device="cpu"):
    """PreSync (rssync.h:19-21) over `world` ranks; `delays` is pre_sync's grid (presync_delays)."""
    curve = presync_grid_sharded(problem, frame_begin, frame_end, delays, stream=1, call_no=call_no, rank=rank,
                                 world=world, device=device)
    b = argmin_cost_delay(curve, delays)
    return float(curve[b]), float(delays[b])


def sync_sharded(problem, initial_delay, frame_begin, frame_end, search_center, search_radius, *, call_no_base=0,
                 rank=0, world=1, device="cpu"):
    """n independent Sync calls (one per syncpoint), syncpoint i on rank i % world.  Result i equals the
    i-th of n consecutive Sync calls of a single problem whose call counter starts at call_no_base."""
    ini = np.ascontiguousarray(initial_delay, dtype=np.float64)
    n = ini.shape[0]
    fb = np.ascontiguousarray(frame_begin, dtype=np.int64)
    fe = np.ascontiguousarray(frame_end, dtype=np.int64)
    cen = np.ascontiguousarray(np.broadcast_to(search_center, (n,)), dtype=np.float64)
    rad = np.ascontiguousarray(np.broadcast_to(search_radius, (n,)), dtype=np.float64)
    mine = np.arange(rank, n, world)
    if mine.size:
        cost, delay = problem.sync_batch(ini[mine], fb[mine], fe[mine], cen[mine], rad[mine],
                                         call_nos=call_no_base + mine.astype(np.uint64))
    else:
        cost, delay = np.empty(0), np.empty(0)
    if world == 1:
        return cost, delay
    import torch.distributed as dist
    allc = _gather_variable(cost, world, device, dist)
    alld = _gather_variable(delay, world, device, dist)
    # rank r contributed syncpoints r, r+world, ...: undo the round-robin
    order = np.concatenate([np.arange(r, n, world) for r in range(world)])
    outc, outd = np.empty(n), np.empty(n)
    outc[order] = allc
    outd[order] = alld
    return outc, outd


def orientation_search_sharded(problem, search_fn, orientations, *, seed, call_no_base=0, rank=0, world=1,
                               device="cpu", batch_fn=None):
    """The 48-variant orientation search (core_testcode.cpp:184-233) with variant k on rank k % world.
    `search_fn(problem, [orientation])` runs one variant (integrate, ingest, PreSync) and returns
    ([cost], [delay]); the RNG call number of variant k is call_no_base + k on every rank, so the
    gathered result equals the single-process loop.  `batch_fn(problem, orientations, call_nos)`, when
    given, runs all of the rank's variants in one call with those explicit call numbers (the engine then
    prepares the variants on host threads while the device evaluates the ones that are ready).
    Every rank returns all (cost, delay) pairs."""
    n = len(orientations)
    mine = list(range(rank, n, world))
    cost, delay = np.empty(len(mine)), np.empty(len(mine))
    if batch_fn is not None and mine:
        problem.set_rng(seed, call_no_base)
        c, d = batch_fn(problem, [orientations[k] for k in mine],
                        np.array([call_no_base + k for k in mine], dtype=np.uint64))
        cost[:], delay[:] = c, d
        mine_loop = []
    else:
        mine_loop = list(enumerate(mine))
    for j, k in mine_loop:
        problem.set_rng(seed, call_no_base + k)
        c, d = search_fn(problem, [orientations[k]])
        cost[j], delay[j] = c[0], d[0]
    if world == 1:
        return cost, delay
    import torch.distributed as dist
    allc = _gather_variable(cost, world, device, dist)
    alld = _gather_variable(delay, world, device, dist)
    order = np.concatenate([np.arange(r, n, world) for r in range(world)])
    outc, outd = np.empty(n), np.empty(n)
    outc[order] = allc
    outd[order] = alld
    return outc, outd
