"""Synthetic GoPro-shaped workloads for the rs-sync loss engine.

The reference's demo driver needs a real MP4 (core_testcode.cpp:97-162); there is none here, so
the inputs the engine would receive from it are synthesised: a fixed-rate gyro quaternion track
(the recurrence of core_testcode.cpp:41-46 on a band-limited angular velocity) and, per frame,
rolling-shutter-consistent ray pairs with per-ray timestamps (core_testcode.cpp:134-158) for a
static scene seen from a smoothly translating camera, with pixel noise and outliers and a known
true gyro delay.  Deterministic for a given (config, seed): numpy's PCG64 streams only.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# Hero6 2.7k 4:3 lens profile, README.md:59 of the reference; resolution thesis pdf-p.30
READOUT = 0.01111
FX = FY = 1186.0
CX, CY = 1355.389, 1020.317
K1, K2, K3, K4 = 0.04440465777694087, 0.01946789951179939, -0.004476697539343917, -0.002042912877740792
WIDTH, HEIGHT = 2704, 2028


# ---- quaternion helpers (w first, Hamilton product: quat.cpp:33-47) --------------------------
def quat_prod(p, q):
    pw, px, py, pz = p[..., 0], p[..., 1], p[..., 2], p[..., 3]
    qw, qx, qy, qz = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([
        pw * qw - px * qx - py * qy - pz * qz,
        pw * qx + px * qw + py * qz - pz * qy,
        pw * qy - px * qz + py * qw + pz * qx,
        pw * qz + px * qy - py * qx + pz * qw,
    ], axis=-1)


def quat_conj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def quat_rotate(q, p):
    """q (x) (0,p) (x) conj(q), quat.cpp:45-47."""
    pq = np.concatenate([np.zeros(p.shape[:-1] + (1,)), p], axis=-1)
    return quat_prod(q, quat_prod(pq, quat_conj(q)))[..., 1:]


def quat_from_aa(aa):
    """quat.cpp:5-17."""
    th2 = np.sum(aa * aa, axis=-1)
    th = np.sqrt(th2)
    safe = np.where(th2 > 0, th, 1.0)
    k = np.where(th2 > 0, np.sin(0.5 * safe) / safe, 0.5)
    w = np.where(th2 > 0, np.cos(0.5 * safe), 1.0)
    return np.concatenate([w[..., None], aa * k[..., None]], axis=-1)


def integrate_gyro(omega, dt):
    """q_i = normalise(exp(omega_i dt_i) (x) q_{i-1}), q_0 = identity (core_testcode.cpp:41-46),
    evaluated as a blocked scan so 30-minute traces do not need a Python loop per sample."""
    n = omega.shape[0]
    dq = quat_from_aa(omega * np.asarray(dt)[..., None])
    dq[0] = np.array([1.0, 0.0, 0.0, 0.0])
    B = 1024
    nb = (n + B - 1) // B
    pad = nb * B - n
    if pad:
        ident = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (pad, 1))
        dq = np.concatenate([dq, ident], axis=0)
    dq = dq.reshape(nb, B, 4)
    local = np.empty_like(dq)
    acc = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (nb, 1))
    for j in range(B):
        acc = quat_prod(dq[:, j], acc)
        acc /= np.linalg.norm(acc, axis=-1, keepdims=True)
        local[:, j] = acc
    out = np.empty_like(local)
    prefix = np.array([1.0, 0.0, 0.0, 0.0])
    for b in range(nb):
        out[b] = quat_prod(local[b], prefix[None, :])
        prefix = out[b, -1] / np.linalg.norm(out[b, -1])
    out = out.reshape(nb * B, 4)[:n]
    return out / np.linalg.norm(out, axis=-1, keepdims=True)


# ---- natural cubic spline on unit knots (minispline.cpp) for the generator's ground truth ----
class QuatTrack:
    def __init__(self, quats, sample_rate, first_timestamp):
        from scipy.linalg import solve_banded
        y = np.asarray(quats, dtype=np.float64)
        n = y.shape[0]
        ab = np.zeros((3, n))
        ab[0, 2:] = 1.0 / 3.0
        ab[1, 1:-1] = 4.0 / 3.0
        ab[2, :-2] = 1.0 / 3.0
        ab[1, 0] = ab[1, -1] = 2.0
        rhs = np.zeros_like(y)
        rhs[1:-1] = y[2:] - 2 * y[1:-1] + y[:-2]
        c = solve_banded((1, 1), ab, rhs)
        d = np.zeros_like(y)
        b = np.zeros_like(y)
        d[:-1] = (c[1:] - c[:-1]) / 3.0
        b[:-1] = (y[1:] - y[:-1]) - (2.0 * c[:-1] + c[1:]) / 3.0
        b[-1] = 3.0 * d[-2] + 2.0 * c[-2] + b[-2]
        self.y, self.b, self.c, self.d = y, b, c, d
        self.n = n
        self.sr = float(sample_rate)
        self.t0 = float(first_timestamp)

    def __call__(self, t):
        """Unit quaternion at gyro time t (array)."""
        x = (np.asarray(t) - self.t0) * self.sr
        idx = np.clip(np.floor(x), 0, self.n - 1).astype(np.int64)
        h = (x - idx)[..., None]
        q = ((self.d[idx] * h + self.c[idx]) * h + self.b[idx]) * h + self.y[idx]
        return q / np.linalg.norm(q, axis=-1, keepdims=True)


# ---- lens model ------------------------------------------------------------------------------
def undistort(px, py):
    """Pixel -> undistorted normalised image point; 9 Newton steps incl. the reference's `8*k4`
    derivative coefficient (core_testcode.cpp:63-95)."""
    x_ = (px - CX) / FX
    y_ = (py - CY) / FY
    theta_d = np.sqrt(x_ * x_ + y_ * y_)
    theta = np.full_like(theta_d, np.pi / 4.0)
    for _ in range(9):
        t2 = theta * theta
        t3 = t2 * theta
        t4 = t2 * t2
        t5 = t2 * t3
        t6 = t3 * t3
        t7 = t3 * t4
        t8 = t4 * t4
        t9 = t4 * t5
        cur = theta + K1 * t3 + K2 * t5 + K3 * t7 + K4 * t9
        dcur = 1 + 3 * K1 * t2 + 5 * K2 * t4 + 7 * K3 * t6 + 8 * K4 * t8
        new = theta - (cur - theta_d) / dcur
        bad = (new >= np.pi / 2) | (new <= 0)
        while np.any(bad):
            new = np.where(bad, (new + theta) / 2.0, new)
            bad = (new >= np.pi / 2) | (new <= 0)
        theta = new
    r = np.tan(theta)
    s = np.where(theta_d < 1e-9, 1.0 / np.cos(theta), r / np.where(theta_d < 1e-9, 1.0, theta_d))
    return x_ * s, y_ * s


LENS = (READOUT, FX, FY, CX, CY, K1, K2, K3, K4)  # rssync_lens field order


def pixel_to_ray(px, py):
    ux, uy = undistort(px, py)
    v = np.stack([ux, uy, np.ones_like(ux)], axis=-1)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def ray_to_pixel(ray):
    """Forward fisheye projection (inverse of pixel_to_ray)."""
    x, y, z = ray[..., 0], ray[..., 1], ray[..., 2]
    rxy = np.sqrt(x * x + y * y)
    theta = np.arctan2(rxy, z)
    t2 = theta * theta
    theta_d = theta * (1 + t2 * (K1 + t2 * (K2 + t2 * (K3 + t2 * K4))))
    s = np.where(rxy > 1e-12, theta_d / np.where(rxy > 1e-12, rxy, 1.0), 1.0)
    return FX * x * s + CX, FY * y * s + CY


# ---- workloads -------------------------------------------------------------------------------
@dataclass
class Workload:
    name: str
    fps: float
    gyro_rate: float
    gyro_t0: float
    quats: np.ndarray          # (n, 4) w,x,y,z
    omega: np.ndarray          # (n, 3) rad/s, the signal the quats were integrated from
    frame_ids: np.ndarray      # (F,) int64
    ts_a: np.ndarray           # (F, N)
    ts_b: np.ndarray           # (F, N)
    rays_a: np.ndarray         # (F, N, 3)
    rays_b: np.ndarray         # (F, N, 3)
    true_delay: np.ndarray     # (F,) seconds
    presync_radius: float
    presync_step: float
    sync_window: int = 60
    syncpoint_distance: int = 120
    meta: dict = field(default_factory=dict)
    px_a: np.ndarray = None    # (F, N, 2) tracked pixel positions in frame f (x, y)
    px_b: np.ndarray = None    # (F, N, 2) ... and in frame f + 1

    @property
    def n_frames(self):
        return int(self.frame_ids.shape[0])

    @property
    def n_rays(self):
        return int(self.ts_a.shape[1])

    def syncpoints(self):
        """`auto` syncpoint list, core_testcode.cpp:270-273."""
        f0, f1 = self.meta.get("span", (int(self.frame_ids[0]), int(self.frame_ids[-1]) + 1))
        return list(range(f0, f1 - self.sync_window, self.syncpoint_distance))

    def true_delay_at(self, frame):
        return float(self.true_delay[int(np.searchsorted(self.frame_ids, frame))])

    def gyro_timestamps_us(self):
        return np.round((self.gyro_t0 + np.arange(self.quats.shape[0]) / self.gyro_rate) * 1e6).astype(np.int64)


def angular_velocity(t, seed=1):
    rng = np.random.default_rng([seed, 7])
    w = np.zeros(t.shape + (3,))
    for ax in range(3):
        nterm = 3 + (ax % 2)
        f = rng.uniform(0.3, 8.0, nterm)
        a = rng.uniform(0.2, 1.0, nterm) / nterm
        ph = rng.uniform(0, 2 * np.pi, nterm)
        for k in range(nterm):
            w[..., ax] += a[k] * np.sin(2 * np.pi * f[k] * t + ph[k])
    return w


def camera_centre(t):
    return np.stack([0.10 * np.sin(0.7 * t), 0.05 * np.sin(1.1 * t + 1.0), 1.5 * t], axis=-1)


def make_workload(name="C1", *, frames=None, rays=None, first_frame=None, seed=1, fps=60.0,
                  gyro_rate=1000.0, radius=None, step=None, true_delay=0.037, drift=None,
                  noise_px=0.3, outlier_frac=0.10, sync_window=60, syncpoint_distance=120,
                  windows_only=None, procs=1, gyro_pad_radius=None):
    """procs > 1: the per-frame part (projection, rolling-shutter re-projection, undistortion) is cut
    into frame chunks and generated by that many forked worker processes -- a frame's data depends only
    on (seed, frame id) and the shared gyro track, so the result is identical to procs = 1."""
    presets = {
        "C1": dict(frames=300, rays=100, first_frame=0, radius=0.2, step=0.002),
        "C2": dict(frames=3300, rays=200, first_frame=3900, radius=0.2, step=0.002),
        "C3": dict(frames=10000, rays=500, first_frame=0, radius=1.0, step=0.001),
        "C4": dict(frames=108000, rays=200, first_frame=0, radius=0.2, step=0.002),
        # C4 as the syncpoint loop sees it: a 30-minute 60 fps trace, a syncpoint every 1000 frames
        # (107 of them), only the frames inside the sync windows tracked (README.md:66: frames may
        # be skipped), delay drifting -45 -> -40 ms
        "C4s": dict(frames=108000, rays=200, first_frame=0, radius=0.2, step=0.002),
        # the same trace with a syncpoint every 120 frames (the reference's default distance): 900
        # syncpoints; only the frames inside the sync windows are tracked
        "C4d": dict(frames=108000, rays=200, first_frame=0, radius=0.2, step=0.002),
        "tiny": dict(frames=12, rays=40, first_frame=5, radius=0.05, step=0.005),
        "small": dict(frames=64, rays=100, first_frame=100, radius=0.1, step=0.002),
    }
    p = dict(presets.get(name, presets["C1"]))
    if frames is not None: p["frames"] = frames
    if rays is not None: p["rays"] = rays
    if first_frame is not None: p["first_frame"] = first_frame
    if radius is not None: p["radius"] = radius
    if step is not None: p["step"] = step
    if name in ("C4", "C4s", "C4d") and drift is None:
        drift = (-0.045, -0.040)  # linear drift, thesis Fig. 8
    if name == "C4s":
        syncpoint_distance = 1000 if syncpoint_distance == 120 else syncpoint_distance
        windows_only = True if windows_only is None else windows_only
    if name == "C4d":
        windows_only = True if windows_only is None else windows_only
    F, N, f0 = p["frames"], p["rays"], p["first_frame"]

    frame_ids = np.arange(f0, f0 + F, dtype=np.int64)
    t_first = f0 / fps
    t_last = (f0 + F) / fps + READOUT
    if drift is not None:
        dtrue = np.linspace(drift[0], drift[1], F)
    else:
        dtrue = np.full(F, float(true_delay))
    span_frames = F
    if windows_only:  # keep the frames pos .. pos + sync_window of every `auto` syncpoint
        keep = np.zeros(F, dtype=bool)
        for pos in range(0, F - sync_window, syncpoint_distance):
            keep[pos:pos + sync_window + 1] = True
        frame_ids, dtrue = frame_ids[keep], dtrue[keep]
        F = int(frame_ids.shape[0])
    # the gyro track covers the frames +- (search radius + 1 s + |delay|); gyro_pad_radius widens it so
    # that workloads that differ only in their search radius share one track (and one scene)
    pad = max(p["radius"], gyro_pad_radius or 0.0) + 1.0 + float(np.max(np.abs(dtrue)))
    g0 = np.floor((t_first - pad) * gyro_rate) / gyro_rate
    ng = int(np.ceil((t_last + 1.0 / fps + pad - g0) * gyro_rate)) + 1
    tg = g0 + np.arange(ng) / gyro_rate
    omega = angular_velocity(tg, seed)
    quats = integrate_gyro(omega, np.full(ng, 1.0 / gyro_rate))
    track = QuatTrack(quats, gyro_rate, g0)

    parts = _frames_parallel(track, frame_ids, dtrue, N, seed, fps, noise_px, outlier_frac, procs)
    ts_a, ts_b, rays_a, rays_b, pa_x, pa_y, pb_x, pb_y = parts

    return Workload(name=name, fps=fps, gyro_rate=gyro_rate, gyro_t0=float(g0), quats=quats,
                    omega=omega, frame_ids=frame_ids, ts_a=np.ascontiguousarray(ts_a),
                    ts_b=np.ascontiguousarray(ts_b), rays_a=np.ascontiguousarray(rays_a),
                    rays_b=np.ascontiguousarray(rays_b), true_delay=dtrue,
                    presync_radius=p["radius"], presync_step=p["step"], sync_window=sync_window,
                    syncpoint_distance=syncpoint_distance,
                    meta=dict(seed=seed, noise_px=noise_px, outlier_frac=outlier_frac,
                              span=(int(f0), int(f0 + span_frames))),
                    px_a=np.ascontiguousarray(np.stack([pa_x, pa_y], axis=-1)),
                    px_b=np.ascontiguousarray(np.stack([pb_x, pb_y], axis=-1)))


def _frames(track, frame_ids, dtrue, N, seed, fps, noise_px, outlier_frac):
    """rays, timestamps and pixels of the given frames (arrays of shape (F, N[, ...]))"""
    F = int(frame_ids.shape[0])
    # per-frame draws (seeded per frame so a frame's data does not depend on the range asked for)
    U = np.empty((F, N, 10))
    for i, fid in enumerate(frame_ids):
        U[i] = np.random.default_rng([seed, int(fid)]).random((N, 10))
    pa_x = 100.0 + U[..., 0] * (WIDTH - 200.0)
    pa_y = 100.0 + U[..., 1] * (HEIGHT - 200.0)
    depth = 2.0 + U[..., 2] * 28.0
    # Box-Muller pixel noise
    rad = np.sqrt(-2.0 * np.log(1.0 - U[..., 3]))
    nx = noise_px * rad * np.cos(2 * np.pi * U[..., 4])
    ny = noise_px * rad * np.sin(2 * np.pi * U[..., 4])
    outlier = U[..., 5] < outlier_frac
    ox = 100.0 + U[..., 6] * (WIDTH - 200.0)
    oy = 100.0 + U[..., 7] * (HEIGHT - 200.0)

    tf_a = (frame_ids / fps)[:, None]
    tf_b = ((frame_ids + 1) / fps)[:, None]
    dcol = dtrue[:, None]
    ts_a = tf_a + READOUT * (pa_y / HEIGHT)
    rays_a = pixel_to_ray(pa_x, pa_y)
    q_a = track(ts_a + dcol)
    world_dir = quat_rotate(quat_conj(q_a), rays_a)
    X = camera_centre(ts_a) + depth[..., None] * world_dir

    pb_y = pa_y.copy()
    pb_x = pa_x.copy()
    for _ in range(6):  # the row fixes the exposure time of the second observation
        ts_b = tf_b + READOUT * (pb_y / HEIGHT)
        q_b = track(ts_b + dcol)
        v = X - camera_centre(ts_b)
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        cam = quat_rotate(q_b, v)
        pb_x, pb_y = ray_to_pixel(cam)
    pb_x = np.where(outlier, ox, pb_x + nx)
    pb_y = np.where(outlier, oy, pb_y + ny)
    ts_b = tf_b + READOUT * (pb_y / HEIGHT)
    rays_b = pixel_to_ray(pb_x, pb_y)
    return ts_a, ts_b, rays_a, rays_b, pa_x, pa_y, pb_x, pb_y


_SHARED = {}


def _frames_chunk(bounds):
    lo, hi = bounds
    a = _SHARED["args"]
    return _frames(a[0], a[1][lo:hi], a[2][lo:hi], *a[3:])


def _frames_parallel(track, frame_ids, dtrue, N, seed, fps, noise_px, outlier_frac, procs):
    F = int(frame_ids.shape[0])
    if procs <= 1 or F < 4 * procs:
        return _frames(track, frame_ids, dtrue, N, seed, fps, noise_px, outlier_frac)
    import multiprocessing as mp
    _SHARED["args"] = (track, frame_ids, dtrue, N, seed, fps, noise_px, outlier_frac)
    n_chunks = procs * 4
    bounds = [(F * i // n_chunks, F * (i + 1) // n_chunks) for i in range(n_chunks)]
    with mp.get_context("fork").Pool(procs) as pool:  # fork: the children see _SHARED without pickling it
        parts = pool.map(_frames_chunk, bounds)
    _SHARED.clear()
    return tuple(np.concatenate([p[k] for p in parts], axis=0) for k in range(8))


# the 48 axis permutation / sign variants of core_testcode.cpp:186-190.  Our mapping (the
# reference's lives in the third-party telemetry_parser crate): character i of the string names
# the INPUT axis routed to output axis i; upper case keeps the sign, lower case flips it.
ORIENTATIONS = [
    "YxZ", "Xyz", "XZy", "Zxy", "zyX", "yxZ", "ZXY", "zYx", "ZYX", "yXz", "YZX", "XyZ",
    "Yzx", "zXy", "YXz", "xyz", "yZx", "XYZ", "zxy", "xYz", "XYz", "zxY", "zXY", "xZy",
    "zyx", "xyZ", "Yxz", "xzy", "yZX", "yzX", "ZYx", "xYZ", "zYX", "ZxY", "yzx", "xZY",
    "Xzy", "XzY", "YzX", "Zyx", "XZY", "yxz", "xzY", "ZyX", "YXZ", "yXZ", "YZx", "ZXy"]


def orient_omega(omega, orient):
    out = np.empty_like(omega)
    for i, ch in enumerate(orient):
        src = "xyz".index(ch.lower())
        out[:, i] = omega[:, src] * (1.0 if ch.isupper() else -1.0)
    return out
