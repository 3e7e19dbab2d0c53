"""The phase-by-phase C-ABI of one scan (lio_est_open_scan / get_parameters / assemble / solve / close_scan, SURVEY 8b:
`assemble(...) -> H, g, cost` and `solve(...)`) against the monolithic call, the oracle, and a compiled C++ client."""
import os
import struct
import subprocess

import numpy as np
import pytest

from lio_mapping_b200 import scenario
from tests import helpers

pytestmark = pytest.mark.gpu


def _est(W, **cfg):
    from lio_mapping_b200 import estimator
    return estimator.Estimator(window_size=W, opt_window_size=W, max_frame_points=1 << 15, max_scan_points=1 << 17, **cfg)


@pytest.mark.parametrize("device_solver", [1, 0])
def test_stepwise_equals_process_scan_and_oracle(oracle, device_solver):
    from lio_mapping_b200 import estimator
    W = 5
    seq = helpers.Sequence(oracle, "vlp16", n_total=9, distort=False)
    cfg = dict(odom_max_iterations=1, opt_extrinsic=0, device_solver=device_solver)
    mk = lambda a, g: estimator.Pim(a, g, np.zeros(3), np.zeros(3), acc_n=0.2, gyr_n=0.02)
    a, b = _est(W, **cfg), _est(W, **cfg)
    eo = oracle.Estimator(window_size=W, opt_window_size=W, odom_max_iterations=1, opt_extrinsic=0)
    helpers.warm_start(a, seq, oracle, W, pose_noise=0.01, seed=1, make_pim=mk)
    helpers.warm_start(b, seq, oracle, W, pose_noise=0.01, seed=1, make_pim=mk)
    helpers.warm_start(eo, seq, oracle, W, pose_noise=0.01, seed=1,
                       make_pim=lambda a_, g_: oracle.Pim(a_, g_, np.zeros(3), np.zeros(3), acc_n=0.2, gyr_n=0.02))
    for k in range(W, 9):
        helpers.feed_scan(a, seq, k)
        helpers.feed_scan(eo, seq, k)
        tt, acc, gyr = seq.imu[k]
        last = seq.t[k - 1]
        for j in range(len(tt)):
            b.process_imu(tt[j] - last, acc[j], gyr[j], tt[j]); last = tt[j]
        b.open_scan(seq.less_flat[k])
        pose, sb, ex = b.parameters()
        H, g, cost = b.assemble(pose, sb, ex)
        Ha, ga, ca = a.normal_equations()            # first linearisation of the monolithic solve: same point, same structure
        if Ha.shape == H.shape:                      # (the gates may shrink the solver's own system; assemble applies none)
            assert np.abs(H - Ha).max() <= 1e-12 * np.abs(Ha).max() and np.abs(g - ga).max() <= 1e-12 * max(1.0, np.abs(ga).max())
            assert abs(cost - ca) <= 1e-12 * ca
        if k == W:                                   # identical inputs on both sides: H, g of the ceres problem vs the oracle's
            Ho, go = eo.normal_equations()
            if Ho.shape == H.shape:
                assert np.abs(H - Ho).max() <= 1e-9 * np.abs(Ho).max()
                assert np.abs(g - go).max() <= 1e-9 * max(1.0, np.abs(go).max())
        # assembling somewhere else leaves the estimator untouched
        H2, g2, c2 = b.assemble(pose + 1e-3, sb, ex)
        assert c2 != cost
        p2, s2, e2 = b.parameters()
        assert np.array_equal(p2, pose) and np.array_equal(s2, sb) and np.array_equal(e2, ex)
        pose, sb, ex, summ = b.solve(pose, sb, ex)
        b.close_scan(pose, sb, ex)
        sa = a.summary()
        assert summ["iterations"] == sa["iterations"] and summ["final_cost"] == sa["final_cost"]
        assert np.array_equal(b.states(), a.states()), k
    with pytest.raises(Exception):
        b.close_scan()                               # nothing open


def _write_scenario(path, seq, oracle, W, n_scans):
    with open(path, "wb") as f:
        f.write(struct.pack("<iii", W, W, n_scans))
        f.write(np.asarray(seq.tf_lb7(), np.float32).tobytes())
        rng = np.random.default_rng(1)
        for k in range(W):
            noise = rng.normal(0, 0.01, 6) if k > 0 else None
            f.write(np.asarray(seq.state16(k, noise), np.float64).tobytes())
            if k > 0:
                tt, acc, gyr = seq.imu[k]
                dts = np.diff(np.concatenate([[seq.t[k - 1]], tt]))
                f.write(struct.pack("<i", len(tt)))
                f.write(np.concatenate([dts[:, None], acc, gyr], 1).astype(np.float64).tobytes())
                a0, g0 = seq.imu_at_frame[0][k - 1], seq.imu_at_frame[1][k - 1]
            else:
                f.write(struct.pack("<i", 0))
                a0 = g0 = np.zeros(3)
            f.write(np.asarray(a0, np.float64).tobytes()); f.write(np.asarray(g0, np.float64).tobytes())
            pts = np.ascontiguousarray(oracle.voxel_grid(seq.less_flat[k], float(np.float32(0.4))), np.float32)
            f.write(struct.pack("<i", pts.shape[0])); f.write(pts.tobytes())
        f.write(np.asarray(seq.imu_at_frame[0][W - 1], np.float64).tobytes())
        f.write(np.asarray(seq.imu_at_frame[1][W - 1], np.float64).tobytes())
        for k in range(W, W + n_scans):
            tt, acc, gyr = seq.imu[k]
            dts = np.diff(np.concatenate([[seq.t[k - 1]], tt]))
            f.write(struct.pack("<i", len(tt)))
            f.write(np.concatenate([dts[:, None], acc, gyr, tt[:, None]], 1).astype(np.float64).tobytes())
            pts = np.ascontiguousarray(seq.less_flat[k], np.float32)
            f.write(struct.pack("<i", pts.shape[0])); f.write(pts.tobytes())


@pytest.mark.parametrize("mode", ["", "stepwise"])
def test_compiled_cxx_shim_matches_python_driver(oracle, tmp_path, mode):
    """examples/estimator_shim.cc (the reference's Estimator control flow over the C ABI, compiled with g++) driven by a
    scenario file reproduces the window states of the Python-driven estimator bit for bit."""
    from lio_mapping_b200 import estimator
    from tests.test_host_operators import _build_shim
    W, n_scans = 5, 3
    seq = helpers.Sequence(oracle, "vlp16", n_total=W + n_scans, distort=False)
    exe = _build_shim(tmp_path)
    scen, out = tmp_path / "scenario.bin", tmp_path / "states.bin"
    _write_scenario(scen, seq, oracle, W, n_scans)
    run = subprocess.run([str(exe), str(scen), str(out)] + ([mode] if mode else []), capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, (run.stdout, run.stderr)
    got = np.fromfile(out, np.float64).reshape(n_scans, W + 1, 16)
    ref = estimator.Estimator(window_size=W, opt_window_size=W, max_frame_points=1 << 15, max_scan_points=1 << 17,
                              **scenario.EST_CFG["vlp16"])
    helpers.warm_start(ref, seq, oracle, W, pose_noise=0.01, seed=1,
                       make_pim=lambda a, g: estimator.Pim(a, g, np.zeros(3), np.zeros(3), acc_n=0.2, gyr_n=0.02))
    for s in range(n_scans):
        helpers.feed_scan(ref, seq, W + s)
        assert np.array_equal(got[s], ref.states()), s
