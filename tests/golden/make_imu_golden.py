"""Generates tests/golden/imu_pose_vel_10s.npz from the reference fixture test/data/imu_pose_vel.txt
(layout per include/utils/LoadVirtual.h:84-106: t, qw qx qy qz, tx ty tz, vx vy vz, gx gy gz,
ax ay az, ba x3, bg x3).  Run in the build container, where /root/reference is mounted:
    python tests/golden/make_imu_golden.py
The first 10 s (2001 rows) are kept; the clean file's bias columns are uninitialised denormals and
are stored as zeros."""
import os

import numpy as np

SRC = "/root/reference/test/data/imu_pose_vel.txt"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "imu_pose_vel_10s.npz")

if __name__ == "__main__":
    rows = []
    with open(SRC) as f:
        for line in f:
            v = line.split()
            if len(v) >= 17:
                rows.append([float(x) for x in v[:17]])
    a = np.array(rows[:2001])
    np.savez_compressed(DST, t=a[:, 0], q_wxyz=a[:, 1:5], p=a[:, 5:8], v=a[:, 8:11], gyro=a[:, 11:14], acc=a[:, 14:17])
    print("wrote", DST, a.shape)
