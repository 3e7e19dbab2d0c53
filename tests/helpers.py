"""Shared scenario builders for the parity tests (inputs only; all checking is oracle-side)."""
import numpy as np

from lio_mapping_b200 import synth


def frame_clouds(oracle, kind, n_frames, t0=1.0, seed0=10):
    """surface_points_less_flat of n_frames consecutive sweeps + their ground-truth lidar poses."""
    sensor, scene, traj = synth.default_config(kind)
    clouds, poses = [], []
    for f in range(n_frames):
        t_end = t0 + 0.1 * f
        sw = synth.make_sweep(sensor, scene, traj, t_end, seed=seed0 + f, distort=False)
        r = oracle.stage_a(sw, sensor.lower_deg, sensor.upper_deg, sensor.rings)
        clouds.append(oracle.voxel_grid(r["less_flat"], 0.4))
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
