import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def pkg():
    return importlib.import_module("rs-sync_b200")


def synth():
    return importlib.import_module("rs-sync_b200.synth")


@pytest.fixture(scope="session")
def rsb():
    return pkg()


@pytest.fixture(scope="session")
def synth_mod():
    return synth()


@pytest.fixture(scope="session")
def oracle_loader():
    from oracle import loader
    loader.lib()
    return loader


_WORKLOADS = {}


def workload(name, **kw):
    key = (name, tuple(sorted(kw.items())))
    if key not in _WORKLOADS:
        _WORKLOADS[key] = synth().make_workload(name, **kw)
    return _WORKLOADS[key]


@pytest.fixture(scope="session")
def w_tiny():
    return workload("tiny")


@pytest.fixture(scope="session")
def w_small():
    return workload("small")


def has_gpu():
    try:
        import ctypes
        lib = pkg().load_library()
        h = ctypes.c_void_p()
        rc = lib.rssync_create(ctypes.byref(h))
        if h:
            lib.rssync_destroy(h)
        return rc == 0
    except Exception:
        return False


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
