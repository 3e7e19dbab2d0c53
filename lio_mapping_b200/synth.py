"""Seeded synthetic lidar sweeps + IMU for the benchmark / parity configurations of BASELINE.json.

Nothing here is part of the product's compute path: it only fabricates inputs:
  * sensor presets follow src/processor_node.cc:64-80 of the reference (VLP-16: 16 rings -15..15 deg,
    HDL-64: 64 rings -24.9..2 deg) plus the 128x4096 stress sensor of BASELINE.json configs[4];
  * the trajectory/IMU model follows the analytic generator behind the reference fixture
    test/data/imu_pose_vel.txt (ellipse + sinusoidal z, euler-angle attitude, 200 Hz);
  * points are emitted in firing order (column-major: all rings of one azimuth, then the next
    azimuth, clockwise so that the reference's azimuth 2*pi - atan2(y, x) increases with time).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

G_NORM = 9.805


@dataclasses.dataclass
class Sensor:
    name: str
    rings: int
    lower_deg: float
    upper_deg: float
    cols: int
    max_range: float = 120.0
    scan_period: float = 0.1


SENSORS = {
    "vlp16": Sensor("vlp16", 16, -15.0, 15.0, 1800, 100.0),
    "hdl64": Sensor("hdl64", 64, -24.9, 2.0, 2032, 120.0),
    "stress128": Sensor("stress128", 128, -25.0, 15.0, 4096, 120.0),
}


# ----------------------------------------------------------------------------------------------
# scenes: axis-aligned boxes.  "room" is hit from the inside, "boxes" from the outside, plus an
# optional ground plane z = 0.
@dataclasses.dataclass
class Scene:
    boxes: np.ndarray            # (B, 6) [xmin ymin zmin xmax ymax zmax], hit from outside
    room: np.ndarray | None      # (6,) hit from inside, or None
    ground: bool


def indoor_scene(seed: int = 1) -> Scene:
    rng = np.random.default_rng(seed)
    room = np.array([-10.0, -7.5, -1.0, 10.0, 7.5, 2.0])
    boxes = []
    for _ in range(6):
        c = rng.uniform([-8, -6], [8, 6])
        if np.hypot(*c) < 3.5:          # keep the trajectory corridor free
            c = c / max(np.hypot(*c), 1e-3) * 4.5
        s = rng.uniform([0.4, 0.4, 0.8], [1.5, 1.5, 2.5])
        boxes.append([c[0] - s[0] / 2, c[1] - s[1] / 2, -1.0, c[0] + s[0] / 2, c[1] + s[1] / 2, -1.0 + s[2]])
    return Scene(np.array(boxes), room, False)


def outdoor_scene(seed: int = 3, n_buildings: int = 40, n_poles: int = 30) -> Scene:
    rng = np.random.default_rng(seed)
    boxes = []
    while len(boxes) < n_buildings:
        c = rng.uniform(-80, 80, size=2)
        r = np.hypot(*c)
        if 38.0 < r < 52.0 or r < 12.0:   # keep the ring road (trajectory) and its centre free
            continue
        s = rng.uniform([6, 6, 4], [20, 20, 15])
        b = [c[0] - s[0] / 2, c[1] - s[1] / 2, 0.0, c[0] + s[0] / 2, c[1] + s[1] / 2, s[2]]
        boxes.append(b)
    npole = 0
    while npole < n_poles:
        c = rng.uniform(-70, 70, size=2)
        r = np.hypot(*c)
        if 41.0 < r < 49.0:
            continue
        boxes.append([c[0] - 0.15, c[1] - 0.15, 0.0, c[0] + 0.15, c[1] + 0.15, rng.uniform(3, 8)])
        npole += 1
    return Scene(np.array(boxes), None, True)


def _raycast_block(origins: np.ndarray, dirs: np.ndarray, scene: Scene, max_range: float) -> np.ndarray:
    n = origins.shape[0]
    best = np.full(n, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
        if scene.ground:
            t = -origins[:, 2] * inv[:, 2]
            ok = (t > 0.05) & np.isfinite(t)
            best = np.where(ok & (t < best), t, best)
        if scene.room is not None:
            lo = (scene.room[None, :3] - origins) * inv
            hi = (scene.room[None, 3:] - origins) * inv
            tfar = np.nanmin(np.maximum(lo, hi), axis=1)
            ok = tfar > 0.05
            best = np.where(ok & (tfar < best), tfar, best)
        B = scene.boxes.shape[0]
        chunk = 8
        for b0 in range(0, B, chunk):
            bx = scene.boxes[b0:b0 + chunk]
            lo = (bx[None, :, :3] - origins[:, None, :]) * inv[:, None, :]
            hi = (bx[None, :, 3:] - origins[:, None, :]) * inv[:, None, :]
            tnear = np.nanmax(np.minimum(lo, hi), axis=2)
            tfar = np.nanmin(np.maximum(lo, hi), axis=2)
            hit = (tnear <= tfar) & (tnear > 0.05)
            t = np.where(hit, tnear, np.inf).min(axis=1)
            best = np.minimum(best, t)
    best[best > max_range] = np.inf
    return best


def raycast(origins: np.ndarray, dirs: np.ndarray, scene: Scene, max_range: float, block: int = 8192) -> np.ndarray:
    """Distance along each ray to the first hit (inf if none). float64, vectorised slab test.  Rays are independent, so
    the work is cut into cache-sized blocks spread over a thread pool (numpy releases the GIL); results do not depend on
    the blocking."""
    n = origins.shape[0]
    if n <= block:
        return _raycast_block(origins, dirs, scene, max_range)
    import concurrent.futures
    import os
    spans = [(a, min(a + block, n)) for a in range(0, n, block)]
    workers = max(1, min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    out = np.empty(n)
    with concurrent.futures.ThreadPoolExecutor(max_workers=workers) as ex:
        for (a, b), r in zip(spans, ex.map(lambda ab: _raycast_block(origins[ab[0]:ab[1]], dirs[ab[0]:ab[1]], scene, max_range), spans)):
            out[a:b] = r
    return out


# ----------------------------------------------------------------------------------------------
# trajectory + IMU (body = IMU frame; world z up; gravity vector (0,0,-g))
def _euler2rot(e):
    roll, pitch, yaw = e[..., 0], e[..., 1], e[..., 2]
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    R = np.empty(e.shape[:-1] + (3, 3))
    R[..., 0, 0] = cy * cp; R[..., 0, 1] = cy * sp * sr - sy * cr; R[..., 0, 2] = sy * sr + cy * cr * sp
    R[..., 1, 0] = sy * cp; R[..., 1, 1] = cy * cr + sy * sr * sp; R[..., 1, 2] = sp * sy * cr - cy * sr
    R[..., 2, 0] = -sp;     R[..., 2, 1] = cp * sr;                R[..., 2, 2] = cp * cr
    return R


def _euler_rates_to_body(e):
    roll, pitch = e[..., 0], e[..., 1]
    cr, sr, cp, sp = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch)
    M = np.zeros(e.shape[:-1] + (3, 3))
    M[..., 0, 0] = 1; M[..., 0, 2] = -sp
    M[..., 1, 1] = cr; M[..., 1, 2] = sr * cp
    M[..., 2, 1] = -sr; M[..., 2, 2] = cr * cp
    return M


@dataclasses.dataclass
class Trajectory:
    ax: float = 3.0      # ellipse semi-axes [m]
    ay: float = 2.0
    az: float = 0.1
    period: float = 40.0
    z0: float = 0.3
    k_roll: float = 0.03
    k_pitch: float = 0.04
    g_norm: float = G_NORM

    def state(self, t):
        t = np.asarray(t, dtype=np.float64)
        K = 2 * math.pi / self.period
        K1 = 5.0
        p = np.stack([self.ax * np.cos(K * t), self.ay * np.sin(K * t), self.az * np.sin(K1 * K * t) + self.z0], -1)
        dp = np.stack([-K * self.ax * np.sin(K * t), K * self.ay * np.cos(K * t), self.az * K1 * K * np.cos(K1 * K * t)], -1)
        ddp = np.stack([-K * K * self.ax * np.cos(K * t), -K * K * self.ay * np.sin(K * t),
                        -self.az * (K1 * K) ** 2 * np.sin(K1 * K * t)], -1)
        # heading follows the velocity direction (yaw = K t + pi/2), small roll/pitch oscillation
        e = np.stack([self.k_roll * np.cos(t), self.k_pitch * np.sin(t), K * t + math.pi / 2], -1)
        de = np.stack([-self.k_roll * np.sin(t), self.k_pitch * np.cos(t), np.full_like(t, K)], -1)
        R = _euler2rot(e)
        gyro = np.einsum("...ij,...j->...i", _euler_rates_to_body(e), de)
        gvec = np.array([0.0, 0.0, -self.g_norm])
        acc = np.einsum("...ji,...j->...i", R, ddp - gvec)
        return p, R, dp, gyro, acc


def rot_to_quat(R):
    """(..., 3, 3) -> (..., 4) as (x, y, z, w), w >= 0."""
    R = np.asarray(R)
    out = np.empty(R.shape[:-2] + (4,))
    flat = R.reshape(-1, 3, 3)
    o = out.reshape(-1, 4)
    for i, m in enumerate(flat):
        tr = m[0, 0] + m[1, 1] + m[2, 2]
        if tr > 0:
            s = math.sqrt(tr + 1.0) * 2
            q = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
        else:
            k = int(np.argmax([m[0, 0], m[1, 1], m[2, 2]]))
            j, l = (k + 1) % 3, (k + 2) % 3
            s = math.sqrt(m[k, k] - m[j, j] - m[l, l] + 1.0) * 2
            q = [0.0] * 4
            q[k] = 0.25 * s
            q[3] = (m[l, j] - m[j, l]) / s
            q[j] = (m[j, k] + m[k, j]) / s
            q[l] = (m[l, k] + m[k, l]) / s
        q = np.array(q)
        if q[3] < 0:
            q = -q
        o[i] = q / np.linalg.norm(q)
    return out


def quat_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


# ----------------------------------------------------------------------------------------------
def make_sweep(sensor: Sensor, scene: Scene, traj: Trajectory | None, t_end: float, seed: int,
               R_lb: np.ndarray | None = None, t_lb: np.ndarray | None = None, range_noise: float = 0.01,
               distort: bool = True, static_pose=None) -> np.ndarray:
    """One sweep ending at time t_end, expressed in the (moving) lidar frame at each firing
    instant.  Returns (N, 4) float32 [x, y, z, intensity]; returns with no hit are dropped."""
    rng = np.random.default_rng(seed)
    Rn, C = sensor.rings, sensor.cols
    elev = np.deg2rad(np.linspace(sensor.lower_deg, sensor.upper_deg, Rn))
    col = np.arange(C)
    az = -2 * math.pi * (col + 0.25) / C          # clockwise; +0.25 keeps atan2 off the +-pi seam
    tcol = t_end - sensor.scan_period + sensor.scan_period * col / C
    d_local = np.stack([np.cos(elev)[None, :] * np.cos(az)[:, None],
                        np.cos(elev)[None, :] * np.sin(az)[:, None],
                        np.broadcast_to(np.sin(elev)[None, :], (C, Rn))], -1)      # (C, R, 3)
    if R_lb is None:
        R_lb = np.eye(3)
    if t_lb is None:
        t_lb = np.zeros(3)
    if static_pose is not None:
        Rwb = np.broadcast_to(static_pose[0], (C, 3, 3))
        pwb = np.broadcast_to(static_pose[1], (C, 3))
    else:
        tq = tcol if distort else np.full(C, t_end)
        pwb, Rwb, _, _, _ = traj.state(tq)
    # T_wl = T_wb * T_lb^-1  (Estimator.cc:1387-1390: rot_l = R_b * R_lb^-1, pos_l = P_b - rot_l * t_lb)
    Rwl = Rwb @ R_lb.T
    pwl = pwb - np.einsum("cij,j->ci", Rwl, t_lb)
    d_world = np.einsum("cij,crj->cri", Rwl, d_local).reshape(-1, 3)
    o_world = np.repeat(pwl, Rn, axis=0)
    rng_t = raycast(o_world, d_world, scene, sensor.max_range)
    ok = np.isfinite(rng_t)
    r = rng_t + rng.normal(0.0, range_noise, size=rng_t.shape)
    pts = d_local.reshape(-1, 3) * r[:, None]
    inten = rng.uniform(0, 100, size=r.shape)
    out = np.concatenate([pts, inten[:, None]], 1)[ok]
    return np.ascontiguousarray(out, dtype=np.float32)


def make_imu(traj: Trajectory, t0: float, t1: float, rate: float = 200.0, seed: int = 0,
             acc_n: float = 0.0, gyr_n: float = 0.0):
    """IMU samples on (t0, t1] at `rate` Hz: returns t, acc (N,3), gyro (N,3)."""
    n0 = int(math.floor(t0 * rate + 1e-9)) + 1
    n1 = int(math.floor(t1 * rate + 1e-9))
    t = np.arange(n0, n1 + 1) / rate
    _, _, _, gyro, acc = traj.state(t)
    if acc_n > 0 or gyr_n > 0:
        rng = np.random.default_rng(seed)
        dt = 1.0 / rate
        acc = acc + rng.normal(0, acc_n / math.sqrt(dt), acc.shape)
        gyro = gyro + rng.normal(0, gyr_n / math.sqrt(dt), gyro.shape)
    return t, acc, gyro


def default_config(kind: str):
    """(sensor, scene, trajectory) triples for the BASELINE.json configs."""
    if kind == "vlp16":
        return SENSORS["vlp16"], indoor_scene(1), Trajectory(ax=3.0, ay=2.0, az=0.05, period=40.0, z0=0.3)
    if kind == "hdl64":
        return SENSORS["hdl64"], outdoor_scene(3), Trajectory(ax=45.0, ay=45.0, az=0.05, period=60.0, z0=1.8)
    if kind == "stress128":
        return SENSORS["stress128"], outdoor_scene(4, 60, 60), Trajectory(ax=45.0, ay=45.0, az=0.05, period=60.0, z0=1.8)
    raise ValueError(kind)
