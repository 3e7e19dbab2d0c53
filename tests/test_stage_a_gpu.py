"""Stage A parity: CUDA path (through the C-ABI) vs the CPU oracle — bit-exact index sets."""
import numpy as np
import pytest

from lio_mapping_b200 import synth

pytestmark = pytest.mark.gpu


def _run_gpu(sw, sensor, **cfg):
    from lio_mapping_b200.point_processor import PointProcessor
    pp = PointProcessor(sensor.lower_deg, sensor.upper_deg, sensor.rings, max_points=max(sw.shape[0], 1), **cfg)
    pp.SetInputCloud(sw)
    pp.Process()
    out = {k: pp.cloud(k) for k in ["laser_scans", "cloud_in_rings", "corner_points_sharp", "corner_points_less_sharp",
                                     "surface_points_flat", "surface_points_less_flat"]}
    out["idx_sharp"] = pp.index("sharp")
    out["idx_less_sharp"] = pp.index("less_sharp")
    out["idx_flat"] = pp.index("flat")
    out["idx_orig"] = pp.index("orig")
    out["scan_ranges"] = pp.scan_ranges()
    out["mask"], out["labels"] = pp.mask_labels()
    out["start_ori"] = pp.start_ori()
    pp.close()
    return out


def _compare(g, o, sensor):
    assert np.array_equal(g["idx_orig"], o["idx_orig_index"])
    assert np.array_equal(g["scan_ranges"], o["scan_ranges"])
    assert np.array_equal(g["laser_scans"][:, :3], o["laser_scans"][:, :3])
    # intensity = ring + rel_time: azimuth uses atan2f -> tolerance
    assert np.allclose(g["laser_scans"][:, 3], o["laser_scans"][:, 3], atol=2e-6 * sensor.rings + 1e-6)
    assert np.allclose(g["cloud_in_rings"][:, 3], o["cloud_in_rings"][:, 3], atol=1e-4)
    assert abs(g["start_ori"] - o["start_ori"]) < 1e-6
    # bit-exact sets
    assert np.array_equal(g["mask"], o["mask"])
    assert np.array_equal(g["labels"], o["labels"])
    assert np.array_equal(g["idx_sharp"], o["idx_sharp"])
    assert np.array_equal(g["idx_less_sharp"], o["idx_less_sharp"])
    assert np.array_equal(g["idx_flat"], o["idx_flat"])
    assert np.array_equal(g["corner_points_sharp"][:, :3], o["sharp"][:, :3])
    assert np.array_equal(g["corner_points_less_sharp"][:, :3], o["less_sharp"][:, :3])
    assert np.array_equal(g["surface_points_flat"][:, :3], o["flat"][:, :3])
    # per-ring voxel grid: same voxels in the same order, centroids bit-equal (same summation order)
    assert g["surface_points_less_flat"].shape == o["less_flat"].shape
    assert np.array_equal(g["surface_points_less_flat"][:, :3], o["less_flat"][:, :3])
    assert np.allclose(g["surface_points_less_flat"][:, 3], o["less_flat"][:, 3], atol=1e-4)


@pytest.mark.parametrize("kind,seed", [("vlp16", 1), ("vlp16", 7), ("hdl64", 3), ("stress128", 4)])
def test_stage_a_parity(oracle, kind, seed):
    sensor, scene, traj = synth.default_config(kind)
    sw = synth.make_sweep(sensor, scene, traj, 1.0 + 0.1 * seed, seed=seed)
    o = oracle.stage_a(sw, sensor.lower_deg, sensor.upper_deg, sensor.rings)
    g = _run_gpu(sw, sensor)
    _compare(g, o, sensor)


def test_stage_a_nan_and_ragged(oracle):
    sensor, scene, traj = synth.default_config("vlp16")
    sw = synth.make_sweep(sensor, scene, traj, 2.0, seed=11)
    rng = np.random.default_rng(5)
    sw = sw.copy()
    sw[rng.integers(0, sw.shape[0], 200), 0] = np.nan          # NaN skip path (PointProcessor.cc:240-244)
    sw[rng.integers(0, sw.shape[0], 50), 2] = np.inf
    keep = np.ones(sw.shape[0], bool)
    ring_of = np.arange(sw.shape[0]) % sensor.rings
    keep[(ring_of == 3)] = False                                  # an empty ring
    keep[(ring_of == 5) & (np.arange(sw.shape[0]) > 16 * 8)] = False   # a ring with <= 11 points (skipped)
    keep[(ring_of == 9) & (rng.uniform(size=sw.shape[0]) < 0.7)] = False  # a ragged ring
    sw = np.ascontiguousarray(sw[keep])
    o = oracle.stage_a(sw, sensor.lower_deg, sensor.upper_deg, sensor.rings)
    g = _run_gpu(sw, sensor)
    _compare(g, o, sensor)


def test_stage_a_empty_and_tiny(oracle):
    sensor, _, _ = synth.default_config("vlp16")
    tiny = np.array([[1, 0, 0, 5], [0, 1, 0.1, 6], [2, 2, -0.2, 7]], np.float32)
    o = oracle.stage_a(tiny, sensor.lower_deg, sensor.upper_deg, sensor.rings)
    g = _run_gpu(tiny, sensor)
    _compare(g, o, sensor)
    assert g["corner_points_sharp"].shape[0] == 0 and g["surface_points_less_flat"].shape[0] == 0
