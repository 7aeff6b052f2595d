"""ctypes loader for oracle/_ref/librssync_ref.so — the UNMODIFIED reference sources compiled
against oracle/shim (TEST INFRASTRUCTURE ONLY).  Built by `make -C oracle ref` where
/root/reference exists; the prebuilt .so travels to the GPU box, nothing here reads
/root/reference at run time."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "librssync_ref.so")
_LIB = None
c_double_p = C.POINTER(C.c_double)
c_i64_p = C.POINTER(C.c_int64)


def available():
    return os.path.exists(PATH)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def lib(path=None):
    """path: an alternative build of the same reference sources (tools/sync_vs_reference.py)"""
    global _LIB
    if path is not None:
        return _bind(C.CDLL(path))
    if _LIB is None:
        _LIB = _bind(C.CDLL(PATH))
    return _LIB


def _bind(L):
    if True:
        L.ref_create.restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_set_gyro_fixed.argtypes = [C.c_void_p, c_double_p, C.c_size_t, C.c_double, C.c_double]
        L.ref_set_gyro_var.argtypes = [C.c_void_p, c_i64_p, c_double_p, C.c_size_t]
        L.ref_set_track.argtypes = [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, C.c_size_t]
        L.ref_frame_order.argtypes = [C.c_void_p, c_i64_p, C.c_int]
        L.ref_gyro_rate.argtypes = [C.c_void_p]
        L.ref_gyro_rate.restype = C.c_double
        L.ref_gyro_start.argtypes = [C.c_void_p]
        L.ref_gyro_start.restype = C.c_double
        L.ref_presync.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_double, C.c_int64, C.c_int64, C.c_double,
                                  C.c_double, c_double_p, c_double_p]
        L.ref_debug_presync.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_int64,
                                        C.c_int64, C.c_double, c_double_p, c_double_p, C.c_int]
        L.ref_sync.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_double, C.c_int64, C.c_int64, C.c_double,
                               C.c_double, c_double_p, c_double_p]
        L.ref_spline_eval.argtypes = [C.c_void_p, c_double_p, C.c_int, c_double_p]
        L.ref_problem_matrix.argtypes = [C.c_void_p, C.c_int64, C.c_double, c_double_p]
        L.ref_guess_motion.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64,
                                       C.c_double, C.c_int, c_double_p]
        L.ref_loss.argtypes = [C.c_void_p, C.c_int64, C.c_double, c_double_p, C.c_double, c_double_p, c_double_p,
                               c_double_p, c_double_p]
        L.ref_slerp.argtypes = [c_double_p, c_double_p, C.c_double, c_double_p]
    return L


class RefProblem:
    """ISyncProblem of the reference itself (SyncProblemPrivate, core_private.hpp:44-61)."""

    def __init__(self, threads=1, seed=100, lib_path=None):
        self.L = lib(lib_path)
        self.h = C.c_void_p(self.L.ref_create())
        self.L.ref_set_threads(threads)
        self.seed = seed
        self.call_no = 0

    def __del__(self):
        try:
            self.L.ref_destroy(self.h)
        except Exception:
            pass

    def set_rng(self, seed, call_no=0):
        self.seed, self.call_no = seed, call_no

    def SetGyroQuaternions(self, *args):
        if len(args) == 4:
            data, count, rate, first = args
            data = np.ascontiguousarray(data, dtype=np.float64)
            self.L.ref_set_gyro_fixed(self.h, _dp(data), count, rate, first)
        else:
            ts, quats, count = args
            ts = np.ascontiguousarray(ts, dtype=np.int64)
            quats = np.ascontiguousarray(quats, dtype=np.float64)
            self.L.ref_set_gyro_var(self.h, ts.ctypes.data_as(c_i64_p), _dp(quats), count)

    def SetTrackResult(self, frame, ts_a, ts_b, rays_a, rays_b, count):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (ts_a, ts_b, rays_a, rays_b)]
        self.L.ref_set_track(self.h, int(frame), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), count)

    def load(self, w):
        self.SetGyroQuaternions(w.quats, w.quats.shape[0], w.gyro_rate, w.gyro_t0)
        for i, fid in enumerate(w.frame_ids):
            self.SetTrackResult(int(fid), w.ts_a[i], w.ts_b[i], w.rays_a[i], w.rays_b[i], w.ts_a.shape[1])
        return self

    def PreSync(self, initial, fb, fe, step, radius):
        c, d = C.c_double(), C.c_double()
        self.L.ref_presync(self.h, self.seed, self.call_no, initial, fb, fe, step, radius, C.byref(c), C.byref(d))
        self.call_no += 1
        return c.value, d.value

    def DebugPreSync(self, initial, fb, fe, radius, point_count, stream=2):
        delays, costs = np.empty(point_count), np.empty(point_count)
        self.L.ref_debug_presync(self.h, self.seed, stream, self.call_no, initial, fb, fe, radius, _dp(delays),
                                 _dp(costs), point_count)
        self.call_no += 1
        return delays, costs

    def Sync(self, initial, fb, fe, center, radius):
        c, d = C.c_double(), C.c_double()
        self.L.ref_sync(self.h, self.seed, self.call_no, initial, fb, fe, center, radius, C.byref(c), C.byref(d))
        self.call_no += 1
        return c.value, d.value

    def presync_grid(self, fb, fe, delays, stream=2, call_no=0, offset_index_base=0):
        """DebugPreSync evaluates a linspace; an arbitrary delay list is evaluated one point at a
        time (point_count = 1 divides by zero in the reference, core_private.cpp:345, so two-point
        calls are used and the first point kept).  The RNG key's offset index is the position in
        the call, so only whole linspace grids are comparable with the oracle at a given index."""
        delays = np.asarray(delays, dtype=np.float64)
        n = len(delays)
        if n >= 2:
            lin = np.array([delays[0] + (delays[-1] - delays[0]) * i / (n - 1) for i in range(n)])
            if np.allclose(lin, delays, rtol=0, atol=1e-12):
                mid, rad = (delays[0] + delays[-1]) / 2, (delays[-1] - delays[0]) / 2
                d, c = np.empty(n), np.empty(n)
                self.L.ref_debug_presync(self.h, self.seed, stream, call_no, mid, fb, fe, rad, _dp(d), _dp(c), n)
                return c
        out = np.empty(n)
        for i, x in enumerate(delays):
            d, c = np.empty(2), np.empty(2)
            self.L.ref_debug_presync(self.h, self.seed, stream, call_no, x + 1e-3, fb, fe, 1e-3, _dp(d), _dp(c), 2)
            out[i] = c[0]
        return out

    def spline_eval(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty((x.shape[0], 4))
        self.L.ref_spline_eval(self.h, _dp(x), x.shape[0], _dp(out))
        return out

    def problem_matrix(self, frame, delay, n):
        P = np.empty((n, 3))
        self.L.ref_problem_matrix(self.h, frame, delay, _dp(P))
        return P

    def guess_motion(self, frame, delay, iters, stream, call_no, offset_idx):
        m = np.empty(3)
        self.L.ref_guess_motion(self.h, self.seed, stream, call_no, offset_idx, frame, delay, iters, _dp(m))
        return m

    def loss(self, frame, delay, m, k):
        m = np.ascontiguousarray(m, dtype=np.float64)
        l3, l5, dd = C.c_double(), C.c_double(), C.c_double()
        g = np.empty(3)
        self.L.ref_loss(self.h, frame, delay, _dp(m), k, C.byref(l3), C.byref(l5), C.byref(dd), _dp(g))
        return l3.value, l5.value, dd.value, g

    def frame_order(self):
        n = self.L.ref_frame_order(self.h, None, 0)
        out = np.empty(n, dtype=np.int64)
        self.L.ref_frame_order(self.h, out.ctypes.data_as(c_i64_p), n)
        return out

    def gyro(self):
        return self.L.ref_gyro_rate(self.h), self.L.ref_gyro_start(self.h)
