"""Synthetic steady-state LIO scenario (sweeps + IMU + ground truth) shared by bench.py, smoke() and
the tests.  Pure input fabrication + glue: no compute of the hot path happens here.

BASELINE.json configs -> scenario kinds:
  "vlp16"     VLP-16 indoor, window 10/10 (configs[1])
  "hdl64"     HDL-64 outdoor_test_config_64, window 10/10 (configs[2], the metric's configuration)
  "stress128" 128 x 4096 sweep, window 15/15 (configs[4])
"""
from __future__ import annotations

import numpy as np

from . import synth

WINDOWS = {"vlp16": 10, "hdl64": 10, "stress128": 15}

# outdoor_test_config_64.yaml of the reference (config/outdoor_test_config_64.yaml:1-71) with the
# synthetic window size; the indoor file differs in prior_factor / cutoff_deskew / keep_features.
EST_CFG = {
    "hdl64": dict(min_match_sq_dis=1.0, min_plane_dis=0.2, surf_filter_size=0.4, keep_features=0, estimate_extrinsic=1,
                  opt_extrinsic=1, imu_factor=1, point_distance_factor=1, prior_factor=1, marginalization_factor=1,
                  enable_deskew=1, cutoff_deskew=1, acc_n=0.2, gyr_n=0.02, acc_w=2e-4, gyr_w=2e-5, g_norm=9.805,
                  max_num_iterations=10, odom_max_iterations=10),
}
EST_CFG["stress128"] = dict(EST_CFG["hdl64"])
EST_CFG["vlp16"] = dict(EST_CFG["hdl64"])


class Scenario:
    """n_total consecutive scans; scan k ends at t0 + 0.1 k.  Raw sweeps only (stage A is the caller's job)."""

    def __init__(self, kind: str = "hdl64", n_total: int = 14, t0: float = 1.0, seed0: int = 100, imu_rate: float = 200.0,
                 tlb=(0.0, 0.0, -0.1), distort: bool = False):
        self.kind = kind
        self.sensor, self.scene, self.traj = synth.default_config(kind)
        self.t = t0 + 0.1 * np.arange(n_total)
        self.R_lb = np.eye(3)
        self.t_lb = np.array(tlb, dtype=np.float64)
        self.raw = [synth.make_sweep(self.sensor, self.scene, self.traj, float(self.t[k]), seed=seed0 + k, R_lb=self.R_lb,
                                     t_lb=self.t_lb, distort=distort) for k in range(n_total)]
        p, R, v, gyro, acc = self.traj.state(self.t)
        self.gt_p, self.gt_R, self.gt_v = p, R, v
        self.gt_q = synth.rot_to_quat(R)
        self.imu_at_frame = (acc, gyro)
        self.imu = [None] + [synth.make_imu(self.traj, float(self.t[k - 1]), float(self.t[k]), imu_rate) for k in range(1, n_total)]

    def state16(self, k, noise=None):
        s = np.zeros(16)
        s[0:3] = self.gt_p[k]; s[3:7] = self.gt_q[k]; s[7:10] = self.gt_v[k]
        if noise is not None:
            s[0:3] += noise[0:3]; s[7:10] += noise[3:6]
        return s

    def tf_lb7(self):
        q = synth.rot_to_quat(self.R_lb)
        return np.array([q[0], q[1], q[2], q[3], *self.t_lb], np.float32)


def warm_start(est, scn: Scenario, W: int, surf_ds_of, make_pim, pose_noise: float = 0.01, seed: int = 1):
    """Initialise an estimator with frames 0..W-1.  surf_ds_of(k) -> down-sampled surf cloud of frame k."""
    rng = np.random.default_rng(seed)
    est.set_extrinsic(scn.tf_lb7())
    for k in range(W):
        pim = None
        if k > 0:
            pim = make_pim(scn.imu_at_frame[0][k - 1], scn.imu_at_frame[1][k - 1])
            tt, acc, gyr = scn.imu[k]
            last = scn.t[k - 1]
            for j in range(len(tt)):
                pim.push_back(tt[j] - last, acc[j], gyr[j])
                last = tt[j]
        noise = rng.normal(0, pose_noise, 6) if (pose_noise > 0 and k > 0) else None
        est.init_frame(k, scn.state16(k, noise), surf_ds_of(k), pim)
    est.finish_init(scn.imu_at_frame[0][W - 1], scn.imu_at_frame[1][W - 1])


def feed_imu(est, scn: Scenario, k: int):
    tt, acc, gyr = scn.imu[k]
    last = scn.t[k - 1]
    if hasattr(est, "process_imu_batch"):   # product: one C-ABI crossing for the scan's IMU messages
        cache = scn.__dict__.setdefault("_imu_batches", {})
        if k not in cache:
            tt64 = np.ascontiguousarray(tt, np.float64)
            cache[k] = (np.ascontiguousarray(np.diff(np.concatenate([[last], tt64]))), np.ascontiguousarray(acc, np.float64),
                        np.ascontiguousarray(gyr, np.float64), tt64)
        est.process_imu_batch(*cache[k])
        return
    for j in range(len(tt)):
        est.process_imu(tt[j] - last, acc[j], gyr[j], tt[j])
        last = tt[j]
