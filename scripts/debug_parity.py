import sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle_py as O
from lio_mapping_b200 import estimator
from tests import helpers
W = 5
seq = helpers.Sequence(O, "vlp16", n_total=10, distort=False)
cfg = dict(odom_max_iterations=1, prior_factor=1)
eo = O.Estimator(window_size=W, opt_window_size=W, **cfg)
eg = estimator.Estimator(window_size=W, opt_window_size=W, max_frame_points=1 << 15, max_scan_points=1 << 17, **cfg)
helpers.warm_start(eo, seq, O, W, pose_noise=0.01, seed=1, make_pim=lambda a, g: O.Pim(a, g, np.zeros(3), np.zeros(3), acc_n=0.2, gyr_n=0.02))
helpers.warm_start(eg, seq, O, W, pose_noise=0.01, seed=1, make_pim=lambda a, g: estimator.Pim(a, g, np.zeros(3), np.zeros(3), acc_n=0.2, gyr_n=0.02))
for k in range(W, 10):
    helpers.feed_scan(eo, seq, k); helpers.feed_scan(eg, seq, k)
    so, sg = eo.summary(), eg.summary()
    Ho, go = eo.normal_equations(); Hg, gg, c0 = eg.normal_equations()
    print("scan", k)
    for key in ["iterations", "successful", "termination", "initial_cost", "final_cost", "cost_pim", "cost_ppp", "cost_marg", "turn_off", "convergence_flag", "map_size", "num_features", "has_prior"]:
        print("   %-18s oracle %-22r gpu %-22r" % (key, so[key], sg[key]))
    print("   H shapes", Ho.shape, Hg.shape)
    if Ho.shape == Hg.shape:
        d = np.abs(Hg - Ho); i = np.unravel_index(d.argmax(), d.shape)
        print("   max|dH| %.3e at %s (|H|max %.3e)  max|dg| %.3e (|g|max %.3e)" % (d.max(), i, np.abs(Ho).max(), np.abs(gg - go).max(), np.abs(go).max()))
        n = Ho.shape[0]
        blk = lambda a: [np.abs(a[15*j:15*j+15, 15*j:15*j+15]).max() for j in range(n // 15)]
        print("   diag-block dH", ["%.1e" % v for v in blk(Hg - Ho)])
    xo, xg = eo.states(), eg.states()
    print("   max state diff pos %.3e quat %.3e vel %.3e bias %.3e" % (np.abs(xg[:, :3] - xo[:, :3]).max(), np.abs(xg[:, 3:7] - xo[:, 3:7]).max(), np.abs(xg[:, 7:10] - xo[:, 7:10]).max(), np.abs(xg[:, 10:] - xo[:, 10:]).max()))
    print("   ex oracle", eo.extrinsic(), "gpu", eg.extrinsic())
