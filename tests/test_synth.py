"""Synthetic inputs: seeded, deterministic, of the documented shapes."""
import numpy as np

from lio_mapping_b200 import synth


def test_sweeps_are_seeded_and_shaped():
    for kind, rings, pts in (("vlp16", 16, 16 * 1800), ("hdl64", 64, None)):
        sensor, scene, traj = synth.default_config(kind)
        assert sensor.rings == rings
        a = synth.make_sweep(sensor, scene, traj, 1.0, seed=5, distort=False)
        b = synth.make_sweep(sensor, scene, traj, 1.0, seed=5, distort=False)
        c = synth.make_sweep(sensor, scene, traj, 1.0, seed=6, distort=False)
        assert a.dtype == np.float32 and a.ndim == 2 and a.shape[1] == 4
        assert np.array_equal(a, b) and not np.array_equal(a, c)
        if pts is not None:
            assert a.shape[0] <= pts and a.shape[0] > 0.9 * pts       # closed room: nearly every ray returns
        assert np.all(np.isfinite(a)) and np.abs(a[:, :3]).max() < 200.0


def test_imu_stream_matches_trajectory():
    sensor, scene, traj = synth.default_config("vlp16")
    tt, acc, gyr = synth.make_imu(traj, 1.0, 1.1, 200.0, seed=1, acc_n=0.0, gyr_n=0.0)
    assert len(tt) == acc.shape[0] == gyr.shape[0] and 19 <= len(tt) <= 21
    assert np.all(np.diff(tt) > 0) and tt[0] > 1.0 and tt[-1] <= 1.1 + 1e-9
    # noise-free specific force has the magnitude of gravity plus the (small) trajectory acceleration
    assert np.all(np.abs(np.linalg.norm(acc, axis=1) - 9.805) < 3.0)
