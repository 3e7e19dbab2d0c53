"""Shared scenario builders for the parity tests (inputs only; all checking is oracle-side)."""
import numpy as np

from lio_mapping_b200 import synth


def frame_clouds(oracle, kind, n_frames, t0=1.0, seed0=10, which="less_flat", leaf=0.4):
    """One stage-A feature cloud (default surface_points_less_flat) of n_frames consecutive sweeps, voxel filtered,
    + their ground-truth lidar poses."""
    sensor, scene, traj = synth.default_config(kind)
    clouds, poses = [], []
    for f in range(n_frames):
        t_end = t0 + 0.1 * f
        sw = synth.make_sweep(sensor, scene, traj, t_end, seed=seed0 + f, distort=False)
        r = oracle.stage_a(sw, sensor.lower_deg, sensor.upper_deg, sensor.rings)
        clouds.append(oracle.voxel_grid(r[which], leaf))
        p, R, _, _, _ = traj.state(np.array(t_end))
        poses.append((R, p))
    return sensor, clouds, poses


def rel_transform(pose_pivot, pose_i):
    """T_pivot^-1 * T_i as (R float32 3x3, t float32 3) and as tf7 (qx,qy,qz,qw,px,py,pz)."""
    Rp, pp = pose_pivot
    Ri, pi = pose_i
    R = Rp.T @ Ri
    t = Rp.T @ (pi - pp)
    q = synth.rot_to_quat(R)
    tf7 = np.array([q[0], q[1], q[2], q[3], t[0], t[1], t[2]], np.float32)
    return R.astype(np.float32), t.astype(np.float32), tf7


def build_map(oracle, clouds, poses, pivot=0, leaf=0.4):
    parts = []
    for i in range(pivot, len(clouds) - 1):
        if i == pivot:
            parts.append(clouds[i])
        else:
            R, t, _ = rel_transform(poses[pivot], poses[i])
            c = oracle.transform_cloud(clouds[i], R, t)
            c[:, 3] = i
            parts.append(c)
    return oracle.voxel_grid(np.concatenate(parts, 0), leaf)


# ---------------------------------------------------------------------------------------------
# sliding-window scenario (steady state): sweeps + IMU + ground-truth states
class Sequence:
    """n_total consecutive scans of a default synthetic config.  Scan k ends at t0 + 0.1 k."""

    def __init__(self, oracle, kind="vlp16", n_total=14, t0=1.0, seed0=100, distort=False, imu_rate=200.0,
                 imu_noise=False, tlb=(0.0, 0.0, -0.1)):
        self.kind = kind
        self.sensor, self.scene, self.traj = synth.default_config(kind)
        self.t = t0 + 0.1 * np.arange(n_total)
        self.R_lb = np.eye(3)
        self.t_lb = np.array(tlb, dtype=np.float64)
        self.rate = imu_rate
        self.raw, self.less_flat = [], []
        for k in range(n_total):
            sw = synth.make_sweep(self.sensor, self.scene, self.traj, float(self.t[k]), seed=seed0 + k, R_lb=self.R_lb,
                                  t_lb=self.t_lb, distort=distort)
            self.raw.append(sw)
            r = oracle.stage_a(sw, self.sensor.lower_deg, self.sensor.upper_deg, self.sensor.rings)
            self.less_flat.append(r["less_flat"])
        p, R, v, gyro, acc = self.traj.state(self.t)
        self.gt_p, self.gt_R, self.gt_v = p, R, v
        self.gt_q = synth.rot_to_quat(R)
        self.imu_at_frame = (acc, gyro)
        an, gn = (0.2, 0.02) if imu_noise else (0.0, 0.0)
        self.imu = [None]
        for k in range(1, n_total):
            self.imu.append(synth.make_imu(self.traj, float(self.t[k - 1]), float(self.t[k]), imu_rate, seed=seed0 + 1000 + k,
                                           acc_n=an * 0.05, gyr_n=gn * 0.05))

    def state16(self, k, noise=None):
        s = np.zeros(16)
        s[0:3] = self.gt_p[k]
        s[3:7] = self.gt_q[k]
        s[7:10] = self.gt_v[k]
        if noise is not None:
            s[0:3] += noise[0:3]
            s[7:10] += noise[3:6]
        return s

    def tf_lb7(self):
        q = synth.rot_to_quat(self.R_lb)
        return np.array([q[0], q[1], q[2], q[3], *self.t_lb], np.float32)


def warm_start(est, seq, oracle_mod, W, pose_noise=0.0, seed=0, make_pim=None):
    """Initialise an estimator (oracle or product: same method names) with frames 0..W-1 of seq."""
    rng = np.random.default_rng(seed)
    est.set_extrinsic(seq.tf_lb7())
    pims = []
    for k in range(W):
        pim = None
        if k > 0:
            acc0, gyr0 = seq.imu_at_frame[0][k - 1], seq.imu_at_frame[1][k - 1]
            pim = make_pim(acc0, gyr0)
            tt, acc, gyr = seq.imu[k]
            last = seq.t[k - 1]
            for j in range(len(tt)):
                pim.push_back(tt[j] - last, acc[j], gyr[j])
                last = tt[j]
        pims.append(pim)
        noise = rng.normal(0, pose_noise, 6) if (pose_noise > 0 and k > 0) else None
        surf_ds = oracle_mod.voxel_grid(seq.less_flat[k], est.cfg["surf_filter_size"])
        est.init_frame(k, seq.state16(k, noise), surf_ds, pim)
    est.finish_init(seq.imu_at_frame[0][W - 1], seq.imu_at_frame[1][W - 1])
    return pims


def feed_scan(est, seq, k):
    tt, acc, gyr = seq.imu[k]
    last = seq.t[k - 1]
    for j in range(len(tt)):
        est.process_imu(tt[j] - last, acc[j], gyr[j], tt[j])
        last = tt[j]
    est.process_scan(seq.less_flat[k])
