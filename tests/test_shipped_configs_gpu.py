"""Full-window parity (product estimator through the C-ABI vs the CPU oracle) on the configurations that are actually
benchmarked and shipped, not only on the small VLP-16 windows of test_estimator_gpu.py:

  * HDL-64 outdoor_test_config_64, window 10/10      - BASELINE.json configs[2], the bench's headline workload;
  * stress 128 x 4096, window 15/15                  - BASELINE.json configs[4];
  * window 7 / opt 5 (outdoor yaml:12-13) and window 12 / opt 7 + keep_features (indoor yaml:12-13, :68): pivot_idx > 0,
    i.e. the frame merge of BuildLocalMap (Estimator.cc:1384, 1409-1441) and the pivot slide of SlideWindow (:2570-2635);
  * keep_features = 1 inside the estimator (Estimator.cc:978-980: LaserOdom rounds append to the newest frame's features).

Tolerances: pose <= 1e-4 relative (north_star), quaternion 1e-4, feature count 0.5 % (10 float LaserOdom rounds reduce in
a different order), local map identical on the first scan (identical inputs by construction).  Velocities carry no
north_star bound; they are checked at 5e-3 relative because the window problem amplifies round-off: on the 7/5 window the
ORACLE run against a copy of itself whose start differs by 1e-11 m agrees to 1e-15 after two scans and only to 1e-9 (pose) /
1e-7 (velocity) after the third - a 1e6 amplification within one scan (the eps = 1e-8 pseudo-inverse threshold of the prior
sits below the round-off of its spectrum) - and the product differs from the oracle by 3e-10 at that point
(measured, round 2: pose 5e-5, velocity 1e-3 relative on the scan after)."""
import numpy as np
import pytest

from lio_mapping_b200 import scenario
from tests import helpers

pytestmark = pytest.mark.gpu


def _pair(oracle, seq, W, O, max_frame_points, **cfg):
    from lio_mapping_b200 import estimator
    eo = oracle.Estimator(window_size=W, opt_window_size=O, **cfg)
    eg = estimator.Estimator(window_size=W, opt_window_size=O, max_frame_points=max_frame_points,
                             max_scan_points=max(len(c) for c in seq.less_flat) + 16, **cfg)
    n5 = dict(acc_n=cfg.get("acc_n", 0.2), gyr_n=cfg.get("gyr_n", 0.02))
    helpers.warm_start(eo, seq, oracle, W, pose_noise=0.01, seed=1, make_pim=lambda a, g: oracle.Pim(a, g, np.zeros(3), np.zeros(3), **n5))
    helpers.warm_start(eg, seq, oracle, W, pose_noise=0.01, seed=1, make_pim=lambda a, g: estimator.Pim(a, g, np.zeros(3), np.zeros(3), **n5))
    return eo, eg


def _perturbed_oracle(oracle, seq, W, O, eps, **cfg):
    """A second oracle whose newest warm-start frame is displaced by eps metres: its divergence from the unperturbed
    oracle measures how strongly the window problem amplifies a difference of the size of fp64 round-off."""
    eo2 = oracle.Estimator(window_size=W, opt_window_size=O, **cfg)
    orig = seq.state16

    def shifted(k, noise=None):
        s16 = orig(k, noise)
        if k == W - 1:
            s16[0] += eps
        return s16
    seq.state16 = shifted
    try:
        n5 = dict(acc_n=cfg.get("acc_n", 0.2), gyr_n=cfg.get("gyr_n", 0.02))
        helpers.warm_start(eo2, seq, oracle, W, pose_noise=0.01, seed=1, make_pim=lambda a, g: oracle.Pim(a, g, np.zeros(3), np.zeros(3), **n5))
    finally:
        seq.state16 = orig
    return eo2


def _run(oracle, seq, W, O, n_scans, max_frame_points, cost_tol=1e-2, calibrate=False, **cfg):
    """calibrate: the bounds become max(nominal, 30 x the oracle's own divergence under a 1e-10 m perturbation of its start)
    - on the short shipped windows (O = 5, 7) the problem amplifies round-off by up to 1e6 within ONE scan (measured: the
    oracle against a copy displaced by 1e-11 m agrees to 1e-15 after two scans and to 1e-9 after the third), so a fixed
    1e-4 cannot be promised by any implementation, the reference's own included."""
    eo, eg = _pair(oracle, seq, W, O, max_frame_points, **cfg)
    eo2 = _perturbed_oracle(oracle, seq, W, O, 1e-10, **cfg) if calibrate else None
    worst = 0.0
    for k in range(W, W + n_scans):
        helpers.feed_scan(eo, seq, k)
        helpers.feed_scan(eg, seq, k)
        so, sg = eo.summary(), eg.summary()
        xo, xg = eo.states(), eg.states()
        scale = max(1.0, np.abs(xo[:, :3]).max())
        vscale = max(1.0, np.abs(xo[:, 7:10]).max())
        amp_p = amp_q = amp_v = amp_c = amp_e = 0.0
        if eo2 is not None:
            helpers.feed_scan(eo2, seq, k)
            x2, s2 = eo2.states(), eo2.summary()
            amp_p = 30 * np.abs(x2[:, :3] - xo[:, :3]).max() / scale
            amp_q = 30 * np.abs(x2[:, 3:7] - xo[:, 3:7]).max()
            amp_v = 30 * np.abs(x2[:, 7:10] - xo[:, 7:10]).max() / vscale
            amp_c = 30 * abs(s2["final_cost"] - so["final_cost"]) / so["final_cost"]
            amp_e = 30 * np.abs(eo2.extrinsic() - eo.extrinsic()).max()
        if k == W:
            assert sg["map_size"] == so["map_size"]
            assert np.array_equal(eg.local_map(), eo.local_map())
        assert abs(sg["map_size"] - so["map_size"]) <= 2 + 1e-4 * so["map_size"], (k, sg["map_size"], so["map_size"])
        assert abs(sg["num_features"] - so["num_features"]) <= 0.005 * so["num_features"], (k, sg["num_features"], so["num_features"])
        assert sg["has_prior"] == so["has_prior"], k
        err = np.abs(xg[:, :3] - xo[:, :3]).max() / scale
        worst = max(worst, err)
        assert err <= max(1e-4, amp_p), (k, err, amp_p)                      # north_star: pose error <= 1e-4 relative
        assert np.abs(xg[:, 3:7] - xo[:, 3:7]).max() <= max(1e-4, amp_q), (k, amp_q)
        assert np.abs(xg[:, 7:10] - xo[:, 7:10]).max() / vscale <= max(5e-3, amp_v), (k, amp_v)
        assert abs(sg["final_cost"] - so["final_cost"]) / so["final_cost"] <= max(cost_tol, amp_c), (k, sg["final_cost"], so["final_cost"], amp_c)
        # the lidar-IMU rotation is barely observable over a short window: same round-off amplification as the velocities
        assert np.abs(eg.extrinsic() - eo.extrinsic()).max() <= max(2e-3, amp_e), (k, amp_e)
        if k <= W + 1:   # before the round-off amplification sets in both sides solve the same well-posed problem
            assert err <= 1e-8 and np.abs(xg[:, 3:7] - xo[:, 3:7]).max() <= 1e-8, (k, err)
    return worst


def test_hdl64_window10_parity(oracle):
    """The bench's headline workload (HDL-64, outdoor_test_config_64, W = O = 10), three consecutive scans."""
    seq = helpers.Sequence(oracle, "hdl64", n_total=13, distort=False)
    _run(oracle, seq, 10, 10, 3, 1 << 16, **scenario.EST_CFG["hdl64"])


def test_stress128_window15_parity(oracle):
    """BASELINE.json configs[4]: 128 x 4096 sweep, W = O = 15, two scans (the second one carries the first prior)."""
    seq = helpers.Sequence(oracle, "stress128", n_total=17, distort=False)
    _run(oracle, seq, 15, 15, 2, 1 << 18, **scenario.EST_CFG["stress128"])


def test_outdoor_shipped_window_7_5(oracle):
    """config/outdoor_test_config_64.yaml:12-13: window 7, opt window 5 (pivot_idx = 2) on HDL-64 sweeps."""
    seq = helpers.Sequence(oracle, "hdl64", n_total=12, distort=False)
    _run(oracle, seq, 7, 5, 5, 1 << 16, calibrate=True, **scenario.EST_CFG["hdl64"])


def test_indoor_shipped_window_12_7_keep_features(oracle):
    """config/indoor_test_config.yaml:12-13, :50, :68: window 12, opt window 7 (pivot_idx = 5), prior_factor 0,
    keep_features 1 on VLP-16 sweeps."""
    seq = helpers.Sequence(oracle, "vlp16", n_total=17, distort=False)
    cfg = dict(scenario.EST_CFG["vlp16"], prior_factor=0, keep_features=1)
    _run(oracle, seq, 12, 7, 5, 1 << 15, calibrate=True, **cfg)


def test_keep_features_in_estimator(oracle):
    """keep_features = 1 with W = O: the newest frame's feature list grows over the LaserOdom rounds (Estimator.cc:978-980)
    and all of them enter the window problem."""
    seq = helpers.Sequence(oracle, "vlp16", n_total=9, distort=False)
    cfg = dict(scenario.EST_CFG["vlp16"], keep_features=1, opt_extrinsic=0)
    eo, eg = _pair(oracle, seq, 5, 5, 1 << 15, **cfg)
    for k in range(5, 9):
        helpers.feed_scan(eo, seq, k)
        helpers.feed_scan(eg, seq, k)
        so, sg = eo.summary(), eg.summary()
        no, ng = eo.features(5)[0].shape[0], eg.features(5)[0].shape[0]
        assert no > 1.5 * eo.features(4)[0].shape[0]          # really accumulated over several rounds
        assert abs(ng - no) <= 0.01 * no, (k, ng, no)
        xo, xg = eo.states(), eg.states()
        assert np.abs(xg[:, :3] - xo[:, :3]).max() <= 1e-4 * max(1.0, np.abs(xo[:, :3]).max()), k
        assert np.abs(xg[:, 3:7] - xo[:, 3:7]).max() <= 1e-4, k
