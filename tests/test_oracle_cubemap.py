"""Rolling cube map of lio::PointMapping, oracle restatement (oracle/o_cubemap.cc; no device counterpart yet): invariants
of re-centring (PointMapping.cc:809-931), cube selection (:944-1003) and UpdateMapDatabase (:1112-1208)."""
import numpy as np


def _cube_of(v, cen):
    c = int((v + 25.0) / 50.0) + cen
    return c - 1 if v + 25.0 < 0 else c


def test_recentre_keeps_world_to_cube_relation(oracle):
    cm = oracle.CubeMap()
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    centre, cen = cm.recentre([0.0, 0.0, 0.0])
    assert centre == (10, 10, 5) and cen == (10, 10, 5)
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.uniform(-120, 120, (500, 2)), rng.uniform(-40, 40, (500, 1)), rng.uniform(0, 9, (500, 1))], 1).astype(np.float32)
    cm.update(pts[:200], pts[200:], [], ident, cen)
    # walk the sensor far enough that the arrays shift along +x, -y and +z
    for pos in ([380.0, 0.0, 0.0], [410.0, -370.0, 0.0], [410.0, -370.0, 160.0]):
        centre, cen = cm.recentre(pos)
        assert all(3 <= c < n - 3 for c, n in zip(centre, (21, 21, 11)))
        assert centre == tuple(_cube_of(p, c) for p, c in zip(pos, cen))
        # every stored point is still found in the cube its world position maps to under the new centre
        kept = 0
        for name, block in (("corner", pts[:200]), ("surf", pts[200:])):
            for p in block:
                ijk = [_cube_of(float(p[a]), cen[a]) for a in range(3)]
                if all(0 <= ijk[a] < n for a, n in zip(range(3), (21, 21, 11))):
                    cube = cm.cube(oracle.CubeMap.to_index(*ijk), name)
                    assert (cube == p).all(1).any()
                    kept += 1
        assert kept > 0


def test_select_matches_angle_criterion(oracle):
    cm = oracle.CubeMap()
    pos = np.array([3.0, -7.0, 1.0], np.float32)
    centre, cen = cm.recentre(pos)
    R = np.eye(3)                                         # sensor z axis = world z
    zaxis = (pos + R @ np.array([0, 0, 10.0])).astype(np.float32)
    valid, surround = cm.select(pos, zaxis, centre)
    assert len(surround) == 125 and set(valid) <= set(surround)
    assert oracle.CubeMap.to_index(*centre) in set(surround)
    # independent statement: a cube is valid iff one of its corners sees the z axis under an angle in (30, 150) degrees
    vset = set(valid.tolist())
    for i in range(centre[0] - 2, centre[0] + 3):
        for j in range(centre[1] - 2, centre[1] + 3):
            for k in range(centre[2] - 2, centre[2] + 3):
                c = 50.0 * (np.array([i, j, k]) - np.array(cen))
                ang = []
                for s in np.array(np.meshgrid([-1, 1], [-1, 1], [-1, 1])).T.reshape(-1, 3):
                    d = c + 25.0 * s - pos
                    ang.append(np.degrees(np.arccos(np.clip(d @ (zaxis - pos) / (np.linalg.norm(d) * 10.0), -1, 1))))
                ang = np.array(ang)
                inside = ((ang > 30.0) & (ang < 150.0)).any()
                near = (np.abs(ang - 30.0) < 1e-3).any() or (np.abs(ang - 150.0) < 1e-3).any()
                if not near:
                    assert (oracle.CubeMap.to_index(i, j, k) in vset) == inside


def test_update_inserts_and_filters_valid_cubes(oracle):
    cm = oracle.CubeMap()
    pos = np.zeros(3, np.float32)
    centre, cen = cm.recentre(pos)
    valid, _ = cm.select(pos, np.array([0, 0, 10.0], np.float32), centre)
    rng = np.random.default_rng(3)
    surf = np.concatenate([rng.uniform(-20, 20, (4000, 3)), rng.uniform(0, 5, (4000, 1))], 1).astype(np.float32)
    corner = np.concatenate([rng.uniform(-20, 20, (800, 3)), rng.uniform(0, 5, (800, 1))], 1).astype(np.float32)
    tf7 = np.array([0, 0, np.sin(0.1), np.cos(0.1), 1.0, 2.0, 0.5], np.float32)
    cm.update(corner, surf, valid, tf7, cen)
    idx = oracle.CubeMap.to_index(*centre)
    assert idx in set(valid.tolist())                      # the sensor's own cube sees the horizon band
    got_s, got_c = cm.cube(idx, "surf"), cm.cube(idx, "corner")
    # expected: points mapped by PointAssociateToMap, those falling into the centre cube, voxel-filtered at 0.4 / 0.2
    Rm = np.array([[np.cos(0.2), -np.sin(0.2), 0], [np.sin(0.2), np.cos(0.2), 0], [0, 0, 1]], np.float32)
    for got, src, leaf in ((got_s, surf, 0.4), (got_c, corner, 0.2)):
        w = oracle.transform_cloud(src, Rm, tf7[4:])       # same rotation up to float rounding of the quaternion product
        inside = np.all((w[:, :3] >= -25.0) & (w[:, :3] < 25.0), axis=1)
        exp = oracle.voxel_grid(w[inside], leaf)
        assert abs(got.shape[0] - exp.shape[0]) <= max(2, exp.shape[0] // 200)
        assert got.shape[0] <= inside.sum()                # filtered (sparse clouds may keep every point)
    # a second identical update grows the cube by at most the new points and filters again (idempotent size up to merging)
    n1 = cm.cube(idx, "surf").shape[0]
    cm.update(corner, surf, valid, tf7, cen)
    assert cm.cube(idx, "surf").shape[0] <= n1 + 5


def test_point_mapping_process_tracks_a_drifting_odometry(oracle):
    """PointMapping::Process in a loop (oracle): the odometry input drifts, the scan-to-map optimisation against the growing cube
    map pulls the mapped pose back towards the ground truth."""
    from lio_mapping_b200 import synth
    from tests import helpers
    sensor, scene, traj = synth.default_config("vlp16")
    pm = oracle.PointMappingOracle()
    p0 = R0 = None
    err_odom, err_map = [], []
    for f in range(7):
        t_end = 1.0 + 0.1 * f
        sw = synth.make_sweep(sensor, scene, traj, t_end, seed=40 + f, distort=False)
        r = oracle.stage_a(sw, sensor.lower_deg, sensor.upper_deg, sensor.rings)
        p, R, _, _, _ = traj.state(np.array(t_end))
        if f == 0:
            p0, R0 = p, R
        Rrel, trel, tf7 = helpers.rel_transform((R0, p0), (R, p))
        drift = np.array([0.03, -0.02, 0.01], np.float32) * f               # accumulated odometry error
        tf_odom = tf7.copy(); tf_odom[4:] += drift
        tobe, info = pm.process(r["less_sharp"], r["less_flat"], tf_odom)
        if f == 0:
            assert info["iterations"] == 0 and info["surf_from_map"] == 0     # empty map: the optimiser returns at its guard
        else:
            assert info["surf_from_map"] > 100 and info["corner_from_map"] > 10 and info["iterations"] >= 1
        err_odom.append(float(np.linalg.norm(tf_odom[4:] - tf7[4:])))
        err_map.append(float(np.linalg.norm(tobe[4:] - tf7[4:])))
    # the mapped pose stays within a few centimetres while the raw odometry has drifted by ~0.2 m
    assert err_odom[-1] > 0.15 and err_map[-1] < 0.05 and max(err_map[1:]) < 0.06, (err_odom, err_map)
