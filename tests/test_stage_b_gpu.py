"""Stage B parity: device VoxelGrid and voxel-hash kNN + plane fit vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,leaf,seed", [(5000, 0.4, 0), (200000, 0.4, 1), (30000, 0.2, 2), (1, 0.4, 3), (700, 5.0, 4)])
def test_voxel_grid_parity(oracle, n, leaf, seed):
    from lio_mapping_b200 import ops
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.uniform(-40, 40, size=(n, 2)), rng.uniform(-2, 6, size=(n, 1)), rng.uniform(0, 9, size=(n, 1))], 1)
    pts = pts.astype(np.float32)
    o = oracle.voxel_grid(pts, leaf)
    g = ops.voxel_grid(pts, leaf)
    assert g.shape == o.shape
    assert np.array_equal(g, o)


def test_voxel_grid_idempotent_and_empty(oracle):
    from lio_mapping_b200 import ops
    assert ops.voxel_grid(np.zeros((0, 4), np.float32), 0.4).shape == (0, 4)
    rng = np.random.default_rng(9)
    pts = rng.uniform(-10, 10, size=(20000, 4)).astype(np.float32)
    g1 = ops.voxel_grid(pts, 0.4)
    # one centroid per voxel -> filtering again with the same grid origin keeps the count if the bbox min voxel is unchanged
    assert g1.shape[0] <= pts.shape[0]
    assert np.all(np.isfinite(g1))


@pytest.mark.parametrize("kind", ["vlp16", "hdl64"])
def test_calculate_features_parity(oracle, kind):
    from lio_mapping_b200 import ops
    sensor, clouds, poses = helpers.frame_clouds(oracle, kind, 4)
    m = helpers.build_map(oracle, clouds, poses)
    for fi, jitter in [(1, 0.0), (3, 0.02)]:
        _, _, tf7 = helpers.rel_transform(poses[0], poses[fi])
        tf7 = tf7.copy()
        tf7[4:] += jitter
        po, co, so = oracle.calculate_features(m, clouds[fi], tf7)
        pg, cg, sg = ops.calculate_features(m, clouds[fi], tf7)
        assert so.shape[0] > 0.3 * clouds[fi].shape[0], "scenario too sparse to be meaningful"
        assert np.array_equal(sg, so)           # identical accepted index set, in order
        assert np.array_equal(pg, po)
        assert np.array_equal(cg, co)           # identical plane coefficients (same float op order)


def test_calculate_features_edge_cases(oracle):
    from lio_mapping_b200 import ops
    rng = np.random.default_rng(3)
    tf7 = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    # map with fewer than 5 points: nothing can match
    m = rng.uniform(-1, 1, size=(3, 4)).astype(np.float32)
    s = rng.uniform(-1, 1, size=(100, 4)).astype(np.float32)
    pg, cg, sg = ops.calculate_features(m, s, tf7)
    po, co, so = oracle.calculate_features(m, s, tf7)
    assert sg.shape[0] == so.shape[0] == 0
    # planar map + queries far away (no neighbours within 1 m) and near
    xy = rng.uniform(-5, 5, size=(4000, 2))
    m = np.concatenate([xy, np.full((4000, 1), -1.5) + rng.normal(0, 0.01, (4000, 1)), np.zeros((4000, 1))], 1).astype(np.float32)
    m = oracle.voxel_grid(m, 0.4)
    s = np.concatenate([rng.uniform(-6, 6, size=(3000, 2)), rng.uniform(-3.5, 0.5, size=(3000, 1)), np.zeros((3000, 1))], 1)
    s = s.astype(np.float32)
    po, co, so = oracle.calculate_features(m, s, tf7)
    pg, cg, sg = ops.calculate_features(m, s, tf7)
    assert np.array_equal(sg, so) and np.array_equal(cg, co)


@pytest.mark.parametrize("kind,keep", [("vlp16", False), ("hdl64", False), ("vlp16", True)])
def test_laser_odom_parity(oracle, kind, keep):
    """CalculateLaserOdom: same iteration count, same feature sets, pose within float round-off of the oracle.

    The 6x6 normal equations are accumulated in double from float rows on both sides but in a different
    order (tree vs sequential), so the solved float update may differ in the last ulp: tolerance 2e-5."""
    from lio_mapping_b200 import ops
    sensor, clouds, poses = helpers.frame_clouds(oracle, kind, 4)
    m = helpers.build_map(oracle, clouds, poses)
    _, _, tf7 = helpers.rel_transform(poses[0], poses[3])
    tf0 = tf7.copy()
    tf0[4:] += np.array([0.05, -0.04, 0.02], np.float32)   # perturbed initial guess: the GN loop has work to do
    to, po, co, so, ito = oracle.laser_odom(m, clouds[3], tf0, keep_features=int(keep))
    tg, pg, cg, sg, itg = ops.laser_odom(m, clouds[3], tf0, keep_features=keep)
    assert itg == ito and ito >= 2
    assert np.allclose(tg, to, atol=2e-5, rtol=0)
    # converged pose is closer to the unperturbed transform than the initial guess
    assert np.linalg.norm(tg[4:] - tf7[4:]) < 0.5 * np.linalg.norm(tf0[4:] - tf7[4:])
    assert abs(sg.shape[0] - so.shape[0]) <= max(2, so.shape[0] // 2000)
    if sg.shape[0] == so.shape[0] and np.array_equal(sg, so):
        assert np.allclose(pg, po, atol=1e-4, rtol=0)
        assert np.allclose(cg, co, atol=1e-4, rtol=0)


def test_laser_odom_first_round_is_exact(oracle):
    """One round only: features are computed from identical inputs -> bit-exact, pose within round-off."""
    from lio_mapping_b200 import ops
    sensor, clouds, poses = helpers.frame_clouds(oracle, "vlp16", 3)
    m = helpers.build_map(oracle, clouds, poses)
    _, _, tf7 = helpers.rel_transform(poses[0], poses[2])
    to, po, co, so, ito = oracle.laser_odom(m, clouds[2], tf7, max_iter=1)
    tg, pg, cg, sg, itg = ops.laser_odom(m, clouds[2], tf7, max_iter=1)
    assert itg == ito == 1
    assert np.array_equal(sg, so) and np.array_equal(pg, po) and np.array_equal(cg, co)
    assert np.allclose(tg, to, atol=2e-6, rtol=0)


def test_transform_to_end_parity(oracle):
    """TransformToEnd / k_deskew: float32 with device sinf/acosf vs libm -> 2e-6 relative to the point range."""
    from lio_mapping_b200 import ops
    rng = np.random.default_rng(5)
    n = 50000
    pts = rng.uniform(-60, 60, size=(n, 3)).astype(np.float32)
    ring = rng.integers(0, 64, size=n).astype(np.float32)
    rel = (rng.uniform(0, 0.1, size=n)).astype(np.float32)
    cloud = np.concatenate([pts, (ring + rel)[:, None]], 1).astype(np.float32)
    for q, p in [((0.01, -0.02, 0.05, 0.9985), (0.8, -0.1, 0.02)), ((0, 0, 0, 1), (0.5, 0, 0)), ((0, 0, 1e-5, 1), (0, 0, 0))]:
        q = np.asarray(q, np.float64); q /= np.linalg.norm(q)
        tf7 = np.concatenate([q, p]).astype(np.float32)
        o = oracle.transform_to_end(cloud, tf7, 10.0)
        g = ops.transform_to_end(cloud, tf7, 10.0)
        assert np.array_equal(g[:, 3], o[:, 3])            # relative time channel is exact
        assert np.max(np.abs(g[:, :3] - o[:, :3])) < 2e-6 * 100.0
    assert ops.transform_to_end(np.zeros((0, 4), np.float32), tf7).shape == (0, 4)


@pytest.mark.parametrize("kind", ["vlp16", "hdl64"])
def test_calculate_line_features_parity(oracle, kind):
    """Point-to-line branch (USE_CORNER / PointMapping corner matching): bit-exact against the oracle."""
    from lio_mapping_b200 import ops
    sensor, clouds, poses = helpers.frame_clouds(oracle, kind, 4, which="less_sharp", leaf=0.2)
    m = helpers.build_map(oracle, clouds, poses, leaf=0.2)
    for fi, jitter in [(1, 0.0), (3, 0.03)]:
        _, _, tf7 = helpers.rel_transform(poses[0], poses[fi])
        tf7 = tf7.copy()
        tf7[4:] += jitter
        po, co, so = oracle.calculate_line_features(m, clouds[fi], tf7)
        pg, cg, sg = ops.calculate_line_features(m, clouds[fi], tf7)
        assert so.shape[0] >= 0.2 * clouds[fi].shape[0], "scenario too sparse to be meaningful"
        assert so.shape[0] % 2 == 0 and np.array_equal(so[0::2], so[1::2])      # two features per accepted corner point
        assert np.array_equal(sg, so)
        assert np.array_equal(pg, po)
        assert np.array_equal(cg, co)


def test_calculate_line_features_edge_cases(oracle):
    from lio_mapping_b200 import ops
    rng = np.random.default_rng(4)
    tf7 = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    # isotropic blob: lambda_3 > 3 lambda_2 fails almost everywhere; exact collinear map: always a line
    blob = np.concatenate([rng.normal(0, 0.2, size=(400, 3)) + [5, 0, 0], np.zeros((400, 1))], 1).astype(np.float32)
    t = np.linspace(-3, 3, 300)
    line = np.stack([5 + 0 * t, 0.5 + 0 * t, t, 0 * t], 1).astype(np.float32)
    q = np.concatenate([rng.normal(0, 0.1, size=(200, 3)) + [5, 0.5, 0], np.zeros((200, 1))], 1).astype(np.float32)
    for m in (blob, line, np.concatenate([blob, line]), line[:4]):
        po, co, so = oracle.calculate_line_features(m, q, tf7)
        pg, cg, sg = ops.calculate_line_features(m, q, tf7)
        assert np.array_equal(sg, so) and np.array_equal(pg, po) and np.array_equal(cg, co)
    assert ops.calculate_line_features(line, np.zeros((0, 4), np.float32), tf7)[0].shape == (0, 4)


@pytest.mark.parametrize("kind,variant", [("vlp16", 0), ("hdl64", 0), ("vlp16", 1), ("hdl64", 1)])
def test_scan_to_map_parity(oracle, kind, variant):
    """PointMapping::OptimizeTransformTobeMapped (variant 0) and MapBuilder::OptimizeMap (variant 1, rotation information
    matrix + left-multiplicative update): corner + surf matching and the 6-DoF float GN, device vs oracle.
    Round 1 works on identical inputs (bit-exact matches); later rounds inherit the last-ulp difference of the 6x6
    reduction order (double tree vs sequential float), so the pose is compared at 2e-5."""
    from lio_mapping_b200 import ops
    sensor, sc, poses = helpers.frame_clouds(oracle, kind, 4)
    _, cc, _ = helpers.frame_clouds(oracle, kind, 4, which="less_sharp", leaf=0.2)
    smap = helpers.build_map(oracle, sc, poses)
    cmap = helpers.build_map(oracle, cc, poses, leaf=0.2)
    _, _, tf7 = helpers.rel_transform(poses[0], poses[3])
    tf0 = tf7.copy()
    tf0[4:] += np.array([0.05, -0.04, 0.02], np.float32)
    # one round: identical matches, in the reference's order (corner block then surf block)
    to, po, co, so, ito = oracle.scan_to_map(cmap, smap, cc[3], sc[3], tf0, max_iter=1, variant=variant)
    tg, pg, cg, sg, itg = ops.scan_to_map(cmap, smap, cc[3], sc[3], tf0, max_iter=1, variant=variant)
    assert itg == ito == 1
    assert np.array_equal(sg, so) and np.array_equal(pg[:, :3], po[:, :3]) and np.array_equal(cg, co)
    assert np.allclose(tg, to, atol=2e-6, rtol=0)
    # full loop
    to, po, co, so, ito = oracle.scan_to_map(cmap, smap, cc[3], sc[3], tf0, variant=variant)
    tg, pg, cg, sg, itg = ops.scan_to_map(cmap, smap, cc[3], sc[3], tf0, variant=variant)
    assert itg == ito and 2 <= ito <= 10
    assert np.allclose(tg, to, atol=2e-5, rtol=0)
    assert np.linalg.norm(tg[4:] - tf7[4:]) < 0.5 * np.linalg.norm(tf0[4:] - tf7[4:])
    assert abs(sg.shape[0] - so.shape[0]) <= max(2, so.shape[0] // 2000)


def test_scan_to_map_guards(oracle):
    """Map-size guard (PointMapping.cc:327-329) and the < 50 matches `continue` (:609-611): the pose is left untouched."""
    from lio_mapping_b200 import ops
    sensor, sc, poses = helpers.frame_clouds(oracle, "vlp16", 3)
    _, cc, _ = helpers.frame_clouds(oracle, "vlp16", 3, which="less_sharp", leaf=0.2)
    smap = helpers.build_map(oracle, sc, poses)
    cmap = helpers.build_map(oracle, cc, poses, leaf=0.2)
    tf0 = np.array([0, 0, 0, 1, 0.1, 0.2, 0.3], np.float32)
    for cm, sm_ in ((cmap[:10], smap), (cmap, smap[:100])):
        tg, pg, cg, sg, itg = ops.scan_to_map(cm, sm_, cc[2], sc[2], tf0)
        to, po, co, so, ito = oracle.scan_to_map(cm, sm_, cc[2], sc[2], tf0)
        assert np.array_equal(tg, tf0) and np.array_equal(to, tf0) and itg == ito == 0
    # queries far from the map: no matches in any round -> 10 skipped rounds, pose unchanged
    far = sc[2].copy(); far[:, :3] += 500.0
    farc = cc[2].copy(); farc[:, :3] += 500.0
    tg, pg, cg, sg, itg = ops.scan_to_map(cmap, smap, farc, far, tf0)
    to, po, co, so, ito = oracle.scan_to_map(cmap, smap, farc, far, tf0)
    assert np.array_equal(tg, tf0) and np.array_equal(to, tf0) and itg == ito == 10 and sg.shape[0] == so.shape[0] == 0
