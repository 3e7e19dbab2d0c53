"""lio::PointOdometry on the device (csrc/podom.cu) against the oracle's restatement (oracle/o_podom.cc):
match indices of the first search, the odometry over a drive, the de-skewed clouds, the /compact_data payload and the
pass-through mode."""
import numpy as np
import pytest

from tests.test_oracle_point_odometry import sweeps

pytestmark = pytest.mark.gpu

KEYS = ("sharp", "less_sharp", "flat", "less_flat", "full")


def _quat_diff(a, b):
    return min(np.abs(a - b).max(), np.abs(a + b).max())


@pytest.mark.parametrize("kind", ["vlp16", "hdl64"])
def test_first_search_is_exact_and_one_round_matches(oracle, kind):
    """With one iteration the whole sweep is one search + one solve: the neighbour indices (nearest point, ring neighbours with
    the sequential first-wins rule) equal the kd-tree path of the oracle, and the step differs only by reduction order."""
    from lio_mapping_b200.point_odometry import PointOdometry
    fr = sweeps(oracle, kind, 2)
    po = oracle.PointOdometryOracle(0.1, 1, 1)
    pg = PointOdometry(0.1, 1, 1)
    for s in fr:
        to, eo, io = po.process(*[s[k] for k in KEYS])
        tg, eg, ig = pg.Process(*[s[k] for k in KEYS])
    assert ig["iterations"] == io["iterations"] == 1 and ig["matches"] == io["matches"] > 100
    s = fr[1]
    for name, key in (("corner", "sharp"), ("surf", "flat")):
        mo, mg = po.matches(name, len(s[key])), pg.matches(name, len(s[key]))
        assert mo.shape == mg.shape and np.array_equal(mo, mg), name
        assert (mo[:, 1] >= 0).sum() > 0.5 * len(mo)
    assert np.abs(eg[4:] - eo[4:]).max() <= 2e-6 + 1e-4 * np.abs(eo[4:]).max() and _quat_diff(eg[:4], eo[:4]) <= 2e-6
    assert np.abs(tg[4:] - to[4:]).max() <= 2e-6 + 1e-4 * np.abs(to[4:]).max() and _quat_diff(tg[:4], to[:4]) <= 2e-6


@pytest.mark.parametrize("kind", ["vlp16", "hdl64"])
def test_odometry_over_a_drive(oracle, kind):
    from lio_mapping_b200.point_odometry import PointOdometry
    fr = sweeps(oracle, kind, 5)
    po = oracle.PointOdometryOracle(0.1, 2, 25)
    pg = PointOdometry(0.1, 2, 25)
    for f, s in enumerate(fr):
        to, eo, io = po.process(*[s[k] for k in KEYS])
        tg, eg, ig = pg.Process(*[s[k] for k in KEYS])
        assert ig["published"] == io["published"] and ig["frame_count"] == io["frame_count"]
        if f == 0:
            assert ig == io and np.array_equal(tg, to) and np.array_equal(eg, eo)
            assert np.array_equal(pg.cloud("last_corner"), s["less_sharp"]) and np.array_equal(pg.cloud("last_surf"), s["less_flat"])
            continue
        assert abs(ig["iterations"] - io["iterations"]) <= 1 and abs(ig["matches"] - io["matches"]) <= 2 + 0.01 * io["matches"]
        scale = max(1.0, float(np.abs(to[4:]).max()))
        assert np.abs(eg[4:] - eo[4:]).max() <= 2e-4 and _quat_diff(eg[:4], eo[:4]) <= 2e-5, (f, eg, eo)
        assert np.abs(tg[4:] - to[4:]).max() <= 3e-4 * scale and _quat_diff(tg[:4], to[:4]) <= 5e-5, (f, tg, to)
        for which in ("last_corner", "last_surf"):
            co, cg = po.cloud(which), pg.cloud(which)
            assert co.shape == cg.shape and np.array_equal(co[:, 3], cg[:, 3])
            assert np.abs(co[:, :3] - cg[:, :3]).max() <= 1e-3
        if io["published"]:
            cd_o, cd_g = po.cloud("compact"), pg.compact_data()
            assert cd_o.shape == cd_g.shape and np.array_equal(cd_o[2, :3], cd_g[2, :3])
            assert np.array_equal(cd_g[0, :3], tg[4:]) and np.array_equal(cd_g[1], tg[:4])
            assert np.abs(cd_o[3:, :3] - cd_g[3:, :3]).max() <= 1e-3
            # receiving side: the payload decodes through the library's own wire decoder
            from lio_mapping_b200 import wire
            tf7, c, sf, full = wire.compact_decode(cd_g)
            assert np.array_equal(tf7, tg) and np.array_equal(c, pg.cloud("last_corner")) and np.array_equal(full, pg.cloud("full"))
        else:
            with pytest.raises(Exception):
                pg.compact_data()
    assert pg.last_launches() <= 25 + 10 + 3


def test_pass_through_is_bit_exact(oracle):
    """After /enable_odom false the node only forwards: clouds untouched, transform_sum_ frozen, payload bit-identical."""
    from lio_mapping_b200.point_odometry import PointOdometry
    fr = sweeps(oracle, "vlp16", 4)
    po = oracle.PointOdometryOracle(0.1, 1, 25)
    pg = PointOdometry(0.1, 1, 25)
    for s in fr[:2]:
        to, _, _ = po.process(*[s[k] for k in KEYS])
        tg, _, _ = pg.Process(*[s[k] for k in KEYS])
    po.set_enable_odom(False); pg.EnableOdom(False)
    for s in fr[2:]:
        to2, _, io = po.process(*[s[k] for k in KEYS])
        tg2, _, ig = pg.Process(*[s[k] for k in KEYS])
        assert ig == io and ig["iterations"] == 0 and ig["published"] == 1
        assert np.array_equal(to2, to) and np.array_equal(tg2, tg)
        assert np.array_equal(pg.cloud("last_corner"), s["less_sharp"]) and np.array_equal(pg.cloud("full"), s["full"])
        cd_o, cd_g = po.cloud("compact"), pg.compact_data()
        assert np.array_equal(cd_o[3:], cd_g[3:]) and np.array_equal(cd_o[2, :3], cd_g[2, :3]) and np.array_equal(cd_g[0, :3], tg[4:])


def test_capacity_and_argument_errors():
    from lio_mapping_b200 import _lib
    from lio_mapping_b200.point_odometry import PointOdometry
    pg = PointOdometry(0.1, 1, 25, max_feature_points=64, max_full_points=64)
    big = np.zeros((65, 4), np.float32)
    small = np.zeros((8, 4), np.float32)
    with pytest.raises(_lib.LioError):
        pg.Process(small, big, small, small, small)
    with pytest.raises(_lib.LioError):
        pg.Process(small, small, small, small, big)
    ts, te, info = pg.Process(small, small, small, small, small)      # first sweep: only stored
    assert info["iterations"] == 0 and np.array_equal(ts, [0, 0, 0, 1, 0, 0, 0])
    ts, te, info = pg.Process(small, small, small, small, small)      # too few last points (:324): no matching, identity motion
    assert info["iterations"] == 0 and np.allclose(ts, [0, 0, 0, 1, 0, 0, 0])
