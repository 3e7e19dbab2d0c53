"""Ring-field variant of stage A (PointToRing for lio::PointXYZIR input, PointProcessor.cc:428-536) on the device vs the
oracle (first run on hardware: round-1 driver GPU tests, both cases passed)."""
import numpy as np
import pytest

from lio_mapping_b200 import synth

pytestmark = pytest.mark.gpu


def _rings_of(sw, sensor):
    ele = np.degrees(np.arctan2(sw[:, 2], np.hypot(sw[:, 0], sw[:, 1])))
    factor = (sensor.rings - 1) / (sensor.upper_deg - sensor.lower_deg)
    return np.clip(((ele - sensor.lower_deg) * factor + 0.5).astype(np.int64), 0, 65535).astype(np.uint16)


@pytest.mark.parametrize("kind,seed", [("vlp16", 2), ("hdl64", 5)])
def test_stage_a_ring_field_parity(oracle, kind, seed):
    from lio_mapping_b200.point_processor import PointProcessor
    sensor, scene, traj = synth.default_config(kind)
    sw = synth.make_sweep(sensor, scene, traj, 1.0 + 0.1 * seed, seed=seed)
    rings = _rings_of(sw, sensor)
    rings[::97] = 300                      # out-of-range ring ids are dropped
    sw = sw.copy(); sw[11::131, 1] = np.nan
    o = oracle.stage_a(sw, sensor.lower_deg, sensor.upper_deg, sensor.rings, ring_field=rings)
    pp = PointProcessor(sensor.lower_deg, sensor.upper_deg, sensor.rings, max_points=sw.shape[0])
    pp.SetInputCloud(sw)
    pp.ProcessWithRingField(rings)
    g_laser, g_full = pp.cloud("laser_scans"), pp.cloud("cloud_in_rings")
    assert np.array_equal(pp.index("orig"), o["idx_orig_index"])
    assert np.array_equal(pp.scan_ranges(), o["scan_ranges"])
    assert np.array_equal(g_laser[:, :3], o["laser_scans"][:, :3])
    assert np.allclose(g_laser[:, 3], o["laser_scans"][:, 3], atol=2e-6 * sensor.rings + 1e-6)   # ring + rel_time (atan2f)
    assert np.allclose(g_full[:, 3], o["cloud_in_rings"][:, 3], atol=1e-4)
    assert abs(pp.start_ori() - o["start_ori"]) < 1e-6
    m, lab = pp.mask_labels()
    assert np.array_equal(m, o["mask"]) and np.array_equal(lab, o["labels"])
    assert np.array_equal(pp.index("sharp"), o["idx_sharp"])
    assert np.array_equal(pp.index("less_sharp"), o["idx_less_sharp"])
    assert np.array_equal(pp.index("flat"), o["idx_flat"])
    assert np.array_equal(pp.cloud("surface_points_less_flat")[:, :3], o["less_flat"][:, :3])
    pp.close()
