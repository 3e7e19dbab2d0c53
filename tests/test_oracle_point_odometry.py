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


def test_first_search_matches_a_numpy_restatement(oracle):
    """Independent check of the neighbour rules (PointOdometry.cc:345-381, :452-506): on the second sweep transform_es_ is still the
    identity, so TransformToStart leaves the queries where they are and the search can be restated in numpy - brute-force nearest
    point (ties by index), then the sequential ring scans with their strict 'smaller wins' updates."""
    fr = sweeps(oracle, "vlp16", 2)
    po = oracle.PointOdometryOracle(0.1, 1, 1)          # one iteration: the indices of the only search stay readable
    for s in fr:
        po.process(s["sharp"], s["less_sharp"], s["flat"], s["less_flat"], s["full"])
    last_c, last_s = fr[0]["less_sharp"], fr[0]["less_flat"]

    def sqd(a, b):
        d = (a[:3] - b[:3]).astype(np.float32)
        return np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])

    def search(q, last, surf):
        d = (last[:, :3] - q[:3]).astype(np.float32)
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        c = int(np.lexsort((np.arange(len(last)), d2))[0])
        if not d2[c] < 25:
            return (-1, -1, -1) if surf else (-1, -1)
        scan = int(last[c, 3])
        i2 = i3 = -1
        b2 = b3 = np.float32(25)
        for j in range(c + 1, len(last)):
            r = int(last[j, 3])
            if r > scan + 2.5:
                break
            v = sqd(last[j], q)
            if surf:
                if r <= scan:
                    if v < b2: b2, i2 = v, j
                elif v < b3: b3, i3 = v, j
            elif r > scan and v < b2: b2, i2 = v, j
        for j in range(c - 1, -1, -1):
            r = int(last[j, 3])
            if r < scan - 2.5:
                break
            v = sqd(last[j], q)
            if surf:
                if r >= scan:
                    if v < b2: b2, i2 = v, j
                elif v < b3: b3, i3 = v, j
            elif r < scan and v < b2: b2, i2 = v, j
        return (c, i2, i3) if surf else (c, i2)

    # sharp / flat points carry ring + rel_time in the intensity; with transform_es_ = identity TransformToStart is the identity map
    mc = po.matches("corner", len(fr[1]["sharp"]))
    ms = po.matches("surf", len(fr[1]["flat"]))
    for i in range(0, len(mc), 3):
        assert tuple(mc[i]) == search(fr[1]["sharp"][i], last_c, False), i
    for i in range(0, len(ms), 7):
        assert tuple(ms[i]) == search(fr[1]["flat"][i], last_s, True), i
    assert (mc[:, 1] >= 0).mean() > 0.5 and ((ms[:, 1] >= 0) & (ms[:, 2] >= 0)).mean() > 0.5
