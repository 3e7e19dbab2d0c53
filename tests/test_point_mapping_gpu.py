"""lio::PointMapping with the cube map in HBM (csrc/cubemap.cu) against the oracle's restatement
(oracle/o_cubemap.cc) on a drifting odometry: cube directory, cube contents and the mapped pose."""
import numpy as np
import pytest

from lio_mapping_b200 import synth
from tests import helpers

pytestmark = pytest.mark.gpu


def _frames(oracle, kind, n, seed0=40):
    sensor, scene, traj = synth.default_config(kind)
    out = []
    p0 = R0 = None
    for f in range(n):
        t_end = 1.0 + 0.1 * f
        sw = synth.make_sweep(sensor, scene, traj, t_end, seed=seed0 + f, distort=False)
        r = oracle.stage_a(sw, sensor.lower_deg, sensor.upper_deg, sensor.rings)
        p, R, _, _, _ = traj.state(np.array(t_end))
        if f == 0:
            p0, R0 = p, R
        _, _, tf7 = helpers.rel_transform((R0, p0), (R, p))
        tf_odom = tf7.copy(); tf_odom[4:] += np.array([0.03, -0.02, 0.01], np.float32) * f      # accumulated odometry error
        out.append((r["less_sharp"], r["less_flat"], tf_odom, tf7))
    return out


@pytest.mark.parametrize("kind", ["vlp16", "hdl64"])
def test_point_mapping_process_parity(oracle, kind):
    from lio_mapping_b200.point_mapping import PointMapping
    frames = _frames(oracle, kind, 6)
    po = oracle.PointMappingOracle()
    pg = PointMapping(max_points=1 << 17)
    for f, (corner, surf, tf_odom, tf_true) in enumerate(frames):
        to, io = po.process(corner, surf, tf_odom)
        tg, ig = pg.Process(corner, surf, tf_odom)
        assert pg.centre() == po.centre()
        if f == 0:
            # empty map: no optimisation on either side, the whole insert + per-cube VoxelGrid is bit-exact
            assert ig == io and ig["iterations"] == 0
            assert np.array_equal(tg, to)
            for which in ("corner", "surf"):
                so, sg = po.cube_sizes(which), pg.cube_sizes(which)
                assert np.array_equal(so, sg)
                for idx in np.nonzero(so)[0]:
                    assert np.array_equal(pg.cube(idx, which), po.cube(idx, which)), (which, idx)
        else:
            # the 6 x 6 float Gauss-Newton reduces in a different order: pose to 1e-4 (like the scan-to-map operator test),
            # the map pulled from the cubes and the cube populations to a fraction of a percent
            assert ig["iterations"] >= 1 and abs(ig["iterations"] - io["iterations"]) <= 1
            for key in ("corner_from_map", "surf_from_map"):
                assert abs(ig[key] - io[key]) <= 2 + 0.005 * io[key], (f, key, ig[key], io[key])
            assert np.abs(tg[4:] - to[4:]).max() <= 2e-4 and np.abs(tg[:4] - to[:4]).max() <= 2e-5, (f, tg, to)
            for which in ("corner", "surf"):
                so, sg = po.cube_sizes(which), pg.cube_sizes(which)
                assert np.array_equal(so > 0, sg > 0)
                assert np.abs(so - sg).sum() <= 2 + 0.005 * so.sum()
        # the mapped pose stays near the truth while the odometry drifts
        if f >= 1:
            assert np.linalg.norm(tg[4:] - tf_true[4:]) < 0.08


def test_point_mapping_recentres_when_the_sensor_leaves_the_centre_cubes(oracle):
    """A pose far from the origin forces the cube array to shift (PointMapping.cc:821-931): directory centre and cube
    placement follow the oracle."""
    from lio_mapping_b200.point_mapping import PointMapping
    frames = _frames(oracle, "vlp16", 2)
    po = oracle.PointMappingOracle()
    pg = PointMapping(max_points=1 << 17)
    for f, (corner, surf, tf_odom, _) in enumerate(frames):
        tf = tf_odom.copy()
        tf[4:] += np.array([430.0, -260.0, 120.0], np.float32) * (f + 1)     # > 8 cubes away along x: several shifts
        to, io = po.process(corner, surf, tf)
        tg, ig = pg.Process(corner, surf, tf)
        assert pg.centre() == po.centre() and pg.centre() != (10, 10, 5)
        for which in ("corner", "surf"):
            assert np.array_equal(po.cube_sizes(which) > 0, pg.cube_sizes(which) > 0)
