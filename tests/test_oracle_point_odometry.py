"""lio::PointOdometry restated (oracle/o_podom.cc): the scan-to-scan odometry recovers the sensor motion of a synthetic
drive from motion-distorted sweeps, and the pass-through mode (enable_odom off) only forwards clouds."""
import numpy as np

from lio_mapping_b200 import synth
from tests import helpers


def sweeps(oracle, kind, n, seed0=70, t0=1.0):
    sensor, scene, traj = synth.default_config(kind)
    out = []
    for f in range(n):
        t_end = t0 + 0.1 * f
        sw = synth.make_sweep(sensor, scene, traj, t_end, seed=seed0 + f, distort=True)
        r = oracle.stage_a(sw, sensor.lower_deg, sensor.upper_deg, sensor.rings)
        p, R, _, _, _ = traj.state(np.array(t_end))
        out.append(dict(sharp=r["sharp"], less_sharp=r["less_sharp"], flat=r["flat"], less_flat=r["less_flat"], full=r["laser_scans"], pose=(R, p)))
    return out


def test_odometry_tracks_the_motion(oracle):
    fr = sweeps(oracle, "vlp16", 6)
    po = oracle.PointOdometryOracle(0.1, 1, 25)
    ts = None
    for f, s in enumerate(fr):
        ts, te, info = po.process(s["sharp"], s["less_sharp"], s["flat"], s["less_flat"], s["full"])
        if f == 0:
            assert info["iterations"] == 0 and info["published"] == 0 and np.array_equal(ts, [0, 0, 0, 1, 0, 0, 0])
            assert np.array_equal(po.cloud("last_corner"), s["less_sharp"])
        else:
            assert 1 <= info["iterations"] <= 25 and info["published"] == 1 and info["matches"] > 100
            _, _, tf_true = helpers.rel_transform(fr[0]["pose"], s["pose"])
            assert np.linalg.norm(ts[4:] - tf_true[4:]) < 0.05 * f + 0.02, (f, ts, tf_true)
            assert min(np.abs(ts[:4] - tf_true[:4]).max(), np.abs(ts[:4] + tf_true[:4]).max()) < 0.01
            # de-skewed clouds carry the ring id only (TransformToEnd: intensity <- int(intensity))
            lc = po.cloud("last_corner")
            assert lc.shape == s["less_sharp"].shape and np.array_equal(lc[:, 3], np.floor(s["less_sharp"][:, 3]))
            # the /compact_data payload decodes to the state it was built from
            tf7, c, sf, full = oracle.compact_decode(po.cloud("compact"))
            assert np.array_equal(tf7, ts) and np.array_equal(c, lc) and np.array_equal(sf, po.cloud("last_surf")) and np.array_equal(full, po.cloud("full"))


def test_pass_through_when_odometry_is_disabled(oracle):
    fr = sweeps(oracle, "vlp16", 3)
    po = oracle.PointOdometryOracle(0.1, 2, 25)
    po.process(*[fr[0][k] for k in ("sharp", "less_sharp", "flat", "less_flat", "full")])
    ts1, _, info1 = po.process(*[fr[1][k] for k in ("sharp", "less_sharp", "flat", "less_flat", "full")])
    assert info1["published"] == 1                       # frame_count 1: 1 % 2 == 1
    po.set_enable_odom(False)                            # Estimator -> /enable_odom false after IMU initialisation
    ts2, _, info2 = po.process(*[fr[2][k] for k in ("sharp", "less_sharp", "flat", "less_flat", "full")])
    assert info2["iterations"] == 0 and info2["published"] == 0 and np.array_equal(ts2, ts1)   # frame_count 2: io_ratio gate
    assert np.array_equal(po.cloud("last_corner"), fr[2]["less_sharp"]) and np.array_equal(po.cloud("last_surf"), fr[2]["less_flat"])
