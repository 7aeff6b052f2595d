"""The C-ABI library loads and exports every symbol include/rssync_b200.h declares, plus the
Itanium symbols a caller compiled against the reference's rssync.h needs.  No compute calls."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT, pkg


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rssync_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rssync_[a-z0-9_]+)\s*\(", text)))


def test_header_and_python_symbol_lists_agree(rsb):
    assert declared_symbols() == sorted(rsb.C_ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(rsb):
    lib = rsb.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_cxx_dropin_symbols_exported(rsb):
    out = subprocess.run(["nm", "-D", "--defined-only", rsb.LIB_PATH], capture_output=True, text=True).stdout
    have = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for sym in rsb.CXX_ABI_SYMBOLS:
        assert sym in have, sym


def test_reference_shaped_caller_links(tmp_path, rsb):
    """a translation unit written against the reference's interface (same declarations as
    src/core/public/rssync.h) compiles against include/rssync.h and links to the library"""
    src = tmp_path / "caller.cpp"
    src.write_text('''
#include <rssync.h>
#include <memory>
#include <vector>
int main(int argc, char**) {
    if (argc > 100) {  // never executed here (no GPU): link check only
        std::unique_ptr<ISyncProblem> sp{CreateSyncProblem()};
        std::vector<double> q(8, 0.0), d(4), c(4);
        std::vector<int64_t> t(2, 0);
        sp->SetGyroQuaternions(q.data(), 2, 1000.0, 0.0);
        sp->SetGyroQuaternions(t.data(), q.data(), 2);
        sp->SetTrackResult(1, q.data(), q.data(), q.data(), q.data(), 2);
        auto a = sp->PreSync(0.0, 0, 10, 0.002, 0.2);
        auto b = sp->Sync(a.second, 0, 10, 0.0, 0.2);
        sp->DebugPreSync(b.second, 0, 10, 0.2, d.data(), c.data(), 4);
    }
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(rsb.LIB_PATH)
    r = subprocess.run(["g++", "-std=c++17", f"-I{ROOT}/include", str(src), "-o", str(exe), f"-L{libdir}",
                        "-lrssync_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_vtable_layout_matches_reference_order(tmp_path):
    """virtual-member order decides the vtable slots; compare declaration order with the
    reference header's (rssync.h:13-28) without reading /root/reference at run time"""
    text = open(os.path.join(ROOT, "include", "rssync.h")).read()
    order = re.findall(r"virtual\s+[\w:<>, ]+?\s+(\w+)\s*\(", text)
    assert order == ["SetGyroQuaternions", "SetGyroQuaternions", "SetTrackResult", "PreSync", "Sync", "DebugPreSync"]
    assert text.index("virtual ~ISyncProblem") < text.index("virtual void SetGyroQuaternions")


def test_no_cpu_fallback_without_gpu(rsb):
    """without a CUDA device the product refuses to run instead of falling back"""
    from conftest import has_gpu
    if has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(rsb.RsSyncError) as e:
        rsb.SyncProblem()
    assert e.value.code == rsb.E_CUDA


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under the product package may reference it"""
    pk = os.path.join(ROOT, "rs-sync_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                body = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "oracle/", "oracle.loader", "liboracle", "orc_",
                               "rssync_oracle", "oracle_math"):
                    assert needle not in body, (os.path.join(dirpath, f), needle)


def test_staging_copy_forms_are_interchangeable(rsb):
    """the bulk SetTrackResult's checked staging copy (host code, no device): scalar, AVX2 and AVX2 with
    non-temporal stores give the same bytes, the same finite / non-finite verdict and the same bounds,
    for every length and destination alignment, with NaN / inf / signed zeros anywhere"""
    import struct
    import numpy as np
    rng = np.random.default_rng(3)
    try:
        rsb.probe_stage_copy(np.zeros(4), 1)
    except rsb.RsSyncError:
        pytest.skip("no AVX2 on this host")
    for n in (0, 1, 3, 4, 5, 31, 200, 600, 1031):
        for mis in (0, 1, 3, 5):
            base = rng.standard_normal(n) * 10.0 ** rng.integers(-300, 300, size=n)
            cases = [base]
            if n:
                for bad in (np.nan, np.inf, -np.inf):
                    c = base.copy()
                    c[rng.integers(0, n)] = bad
                    cases.append(c)
                z = base.copy()
                z[rng.integers(0, n)] = -0.0
                z[rng.integers(0, n)] = 0.0
                cases.append(z)
                cases.append(np.full(n, -0.0))
            for x in cases:
                b0 = (float(x[0]), float(x[0])) if n and np.isfinite(x[0]) else (0.0, 0.0)
                ref = rsb.probe_stage_copy(x, 0, bounds=b0, misalign=mis)
                assert ref[1] == bool(np.all(np.isfinite(x)))
                for mode in (1, 2):
                    got = rsb.probe_stage_copy(x, mode, bounds=b0, misalign=mis)
                    assert got[0].tobytes() == ref[0].tobytes() == x.tobytes()
                    assert got[1] == ref[1]
                    if ref[1]:  # bounds are only used when the data passed the check
                        assert struct.pack("dd", got[2], got[3]) == struct.pack("dd", ref[2], ref[3])
                        assert got[2] == min(b0[0], x.min()) and got[3] == max(b0[1], x.max()) if n else True
                    assert rsb.probe_stage_copy(x, mode, misalign=mis)[1] == ref[1]
