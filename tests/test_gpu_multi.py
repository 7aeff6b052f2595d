"""One problem spread over several GPUs of ONE process (rssync_create_multi; the C++ drop-in picks it
up through RSSYNC_DEVICES): every result must be bit-identical to the single-device result — the RNG
is keyed by global offset / call numbers, each frame's reduction stays on one device.

Needs >= 2 GPUs: skipped on the single-GPU test box; run with `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu`.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, workload

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs in one process")


@pytest.fixture(scope="module")
def pair(rsb):
    w = workload("small")
    n = min(_n_gpus(), 4)
    one = rsb.SyncProblem(seed=100).load(w, bulk=True)
    many = rsb.SyncProblem(seed=100, devices=list(range(n))).load(w, bulk=True)
    assert many.device_count() == n and one.device_count() == 1
    return one, many, w


@needs2
def test_grid_sharded_by_offset_equals_one_device(pair):
    one, many, w = pair
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    for p in (one, many):
        p.set_rng(100, 5)
    d1, c1 = one.DebugPreSync(0.0, fb, fe, 0.1, 101)          # 101 offsets: uneven shards
    d2, c2 = many.DebugPreSync(0.0, fb, fe, 0.1, 101)
    assert np.array_equal(d1, d2) and np.array_equal(c1, c2)
    assert one.PreSync(0.0, fb, fe, w.presync_step, w.presync_radius) == many.PreSync(0.0, fb, fe, w.presync_step, w.presync_radius)
    # fewer delays than devices, a single delay, an empty frame range
    for delays in (np.array([0.03]), np.array([0.01, 0.02, 0.03])[:many.device_count() - 1 or 1]):
        assert np.array_equal(one.presync_grid(fb, fe, delays, call_no=9, offset_index_base=7),
                              many.presync_grid(fb, fe, delays, call_no=9, offset_index_base=7))
    assert np.array_equal(many.presync_grid(10 ** 6, 10 ** 6 + 5, np.array([0.0, 0.1])), np.zeros(2))
    assert many.stats()["last_grid_tasks"] == one.stats()["last_grid_tasks"]


@needs2
def test_sync_and_windows_sharded_by_syncpoint(pair):
    one, many, w = pair
    f0 = int(w.frame_ids[0])
    fbs = np.array([f0 + 7 * i for i in range(5)])
    fes = fbs + 25
    for p in (one, many):
        p.set_rng(100, 40)
    c1, d1 = one.presync_windows(0.0, fbs, fes, 0.004, 0.08)
    c2, d2 = many.presync_windows(0.0, fbs, fes, 0.004, 0.08)
    assert np.array_equal(c1, c2) and np.array_equal(d1, d2)
    s1 = one.sync_batch(d1, fbs, fes, 0.0, 0.2)
    s2 = many.sync_batch(d2, fbs, fes, 0.0, 0.2)
    assert np.array_equal(s1[0], s2[0]) and np.array_equal(s1[1], s2[1])
    assert one.call_counter() == many.call_counter() == 50
    assert one.stats()["sync_lbfgs_evals"] == many.stats()["sync_lbfgs_evals"]
    # a single Sync call runs on the primary
    assert one.Sync(0.038, f0, f0 + 30, 0.0, 0.2) == many.Sync(0.038, f0, f0 + 30, 0.0, 0.2)


@needs2
def test_inputs_changed_after_a_call_are_replicated_again(pair, synth_mod):
    one, many, w = pair
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    delays = np.linspace(-0.02, 0.02, 9)
    before = many.presync_grid(fb, fe, delays, call_no=3)
    w2 = synth_mod.make_workload("small", seed=9)
    for p in (one, many):  # replace one frame, then the gyro
        p.SetTrackResult(int(w.frame_ids[5]), w2.ts_a[5], w2.ts_b[5], w2.rays_a[5], w2.rays_b[5], w.n_rays)
    a, b = one.presync_grid(fb, fe, delays, call_no=3), many.presync_grid(fb, fe, delays, call_no=3)
    assert np.array_equal(a, b) and not np.array_equal(b, before)
    for p in (one, many):
        p.SetGyroQuaternions(w2.quats, w2.quats.shape[0], w2.gyro_rate, w2.gyro_t0)
    assert np.array_equal(one.presync_grid(fb, fe, delays, call_no=3), many.presync_grid(fb, fe, delays, call_no=3))
    for p in (one, many):  # back to the fixture's state
        p.load(w, bulk=True)
    assert np.array_equal(many.presync_grid(fb, fe, delays, call_no=3), before)


@needs2
def test_orientation_search_sharded_by_variant(pair, synth_mod):
    one, many, w = pair
    ts = w.gyro_t0 + np.arange(w.quats.shape[0]) / w.gyro_rate
    fb, fe = int(w.frame_ids[0]), int(w.frame_ids[-1]) + 1
    orients = synth_mod.ORIENTATIONS[:7]
    for p in (one, many):
        p.set_rng(100, 3)
    c1, d1 = one.orientation_search(ts, w.omega, orients, 0.0, fb, fe, 0.005, 0.05)
    c2, d2 = many.orientation_search(ts, w.omega, orients, 0.0, fb, fe, 0.005, 0.05)
    assert np.array_equal(c1, c2) and np.array_equal(d1, d2)
    assert one.call_counter() == many.call_counter()
    # both are left holding the last variant's gyro
    delays = np.linspace(-0.02, 0.02, 5)
    assert np.array_equal(one.presync_grid(fb, fe, delays, call_no=1), many.presync_grid(fb, fe, delays, call_no=1))
    for p in (one, many):
        p.load(w, bulk=True)


@needs2
def test_cxx_caller_uses_all_devices_through_the_environment(rsb, tmp_path):
    """an unmodified ISyncProblem caller (core_testcode-shaped): RSSYNC_DEVICES spreads it over two
    GPUs, and it prints what it prints on one"""
    w = workload("small")
    blob = tmp_path / "workload.bin"
    with open(blob, "wb") as f:
        f.write(struct.pack("<qqqdd", w.quats.shape[0], w.n_frames, w.n_rays, w.gyro_rate, w.gyro_t0))
        f.write(np.ascontiguousarray(w.quats).tobytes())
        f.write(np.ascontiguousarray(w.frame_ids).tobytes())
        for a in (w.ts_a, w.ts_b, w.rays_a, w.rays_b):
            f.write(np.ascontiguousarray(a).tobytes())
    src = tmp_path / "caller.cpp"
    src.write_text(r'''
#include <rssync.h>
#include <cstdio>
#include <cstdint>
#include <memory>
#include <vector>
int main(int, char** argv) {
    FILE* f = std::fopen(argv[1], "rb");
    int64_t nq, nf, nr; double rate, t0;
    if (!f || std::fread(&nq, 8, 1, f) != 1 || std::fread(&nf, 8, 1, f) != 1 || std::fread(&nr, 8, 1, f) != 1 ||
        std::fread(&rate, 8, 1, f) != 1 || std::fread(&t0, 8, 1, f) != 1) return 2;
    std::vector<double> q(4 * nq), tsa(nf * nr), tsb(nf * nr), ra(3 * nf * nr), rb(3 * nf * nr);
    std::vector<int64_t> ids(nf);
    if (std::fread(q.data(), 8, q.size(), f) != q.size() || std::fread(ids.data(), 8, nf, f) != (size_t)nf ||
        std::fread(tsa.data(), 8, tsa.size(), f) != tsa.size() || std::fread(tsb.data(), 8, tsb.size(), f) != tsb.size() ||
        std::fread(ra.data(), 8, ra.size(), f) != ra.size() || std::fread(rb.data(), 8, rb.size(), f) != rb.size()) return 3;
    std::unique_ptr<ISyncProblem> sp{CreateSyncProblem()};
    sp->SetGyroQuaternions(q.data(), (size_t)nq, rate, t0);
    for (int64_t i = 0; i < nf; ++i)
        sp->SetTrackResult(ids[i], &tsa[i * nr], &tsb[i * nr], &ra[3 * i * nr], &rb[3 * i * nr], (size_t)nr);
    std::vector<double> dd(41), cc(41);
    sp->DebugPreSync(0.0, ids[0], ids[0] + nf, 0.08, dd.data(), cc.data(), 41);
    auto p = sp->PreSync(0.0, ids[0], ids[0] + nf, 0.002, 0.1);
    auto s = sp->Sync(p.second, ids[0], ids[0] + 30, 0.0, 0.2);
    std::printf("%a %a %a %a", p.first, p.second, s.first, s.second);
    for (double c : cc) std::printf(" %a", c);
    std::printf("\n");
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(rsb.LIB_PATH)
    r = subprocess.run(["g++", "-std=c++17", f"-I{ROOT}/include", str(src), "-o", str(exe), f"-L{libdir}",
                        "-lrssync_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    outs = []
    for devs in (None, "0,1"):
        env = dict(os.environ)
        env.pop("RSSYNC_DEVICES", None)
        if devs:
            env["RSSYNC_DEVICES"] = devs
        r = subprocess.run([str(exe), str(blob)], capture_output=True, text=True, cwd=tmp_path, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1].split())  # (NCCL prints its version banner to stdout first)
    assert outs[0] == outs[1] and len(outs[0]) == 45
