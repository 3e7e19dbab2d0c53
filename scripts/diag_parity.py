#!/usr/bin/env python
"""Per-scan parity table (product estimator vs oracle) for one window shape; diagnostic, GPU box only.
   python scripts/diag_parity.py KIND W O NSCANS ["dict(key=value, ...)"] [gpu-only overrides dict]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_py as O
from tests import helpers
from lio_mapping_b200 import scenario, estimator

O.build()
kind, W, Oo, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = dict(scenario.EST_CFG[kind])
if len(sys.argv) > 5:
    cfg.update(eval(sys.argv[5]))
gpu_extra = eval(sys.argv[6]) if len(sys.argv) > 6 else {}
seq = helpers.Sequence(O, kind, n_total=W + n, distort=False)
eo = O.Estimator(window_size=W, opt_window_size=Oo, **cfg)
eg = estimator.Estimator(window_size=W, opt_window_size=Oo, max_frame_points=1 << (18 if kind == "stress128" else 16),
                         max_scan_points=max(len(c) for c in seq.less_flat) + 16, **dict(cfg, **gpu_extra))
mk = dict(acc_n=cfg["acc_n"], gyr_n=cfg["gyr_n"])
helpers.warm_start(eo, seq, O, W, pose_noise=0.01, seed=1, make_pim=lambda a, g: O.Pim(a, g, np.zeros(3), np.zeros(3), **mk))
helpers.warm_start(eg, seq, O, W, pose_noise=0.01, seed=1, make_pim=lambda a, g: estimator.Pim(a, g, np.zeros(3), np.zeros(3), **mk))
print("kind %s W %d O %d cfg %s gpu %s" % (kind, W, Oo, sys.argv[5] if len(sys.argv) > 5 else "", gpu_extra))
for k in range(W, W + n):
    helpers.feed_scan(eo, seq, k)
    t0 = time.perf_counter(); helpers.feed_scan(eg, seq, k); tg = time.perf_counter() - t0
    so, sg = eo.summary(), eg.summary()
    xo, xg = eo.states(), eg.states()
    scale = max(1.0, np.abs(xo[:, :3]).max())
    print("k %2d  map %6d/%6d feat %7d/%7d it %2d/%2d succ %2d/%2d cost %.9g/%.9g  pos %.2e quat %.2e vel %.2e ba %.2e bg %.2e prior %d/%d  t_solve %.2f ms t_gpu %.2f ms" % (
        k, sg["map_size"], so["map_size"], sg["num_features"], so["num_features"], sg["iterations"], so["iterations"],
        sg["successful"], so.get("successful", -1), sg["final_cost"], so["final_cost"],
        np.abs(xg[:, :3] - xo[:, :3]).max() / scale, np.abs(xg[:, 3:7] - xo[:, 3:7]).max(), np.abs(xg[:, 7:10] - xo[:, 7:10]).max(),
        np.abs(xg[:, 10:13] - xo[:, 10:13]).max(), np.abs(xg[:, 13:16] - xo[:, 13:16]).max(), sg["has_prior"], so["has_prior"],
        1e3 * sg["t_solve"], 1e3 * tg))

tr = eg.solver_trace()
if tr.any():
    names = ["verdict", "lidar_blk", "gradient", "H_gather", "alpha", "tiles", "cholesky", "dogleg", "tail"]
    print("device solver trace of the last solve (us): gap = kernel entry - previous kernel exit (asm_ppp + k_factors + launch latency)")
    prev_end = None
    for ev in range(24):
        if tr[ev, 0] == 0:
            continue
        c = tr[ev, 1:11].astype(np.float64)
        ph = []
        last = c[0]
        for k in range(1, 10):
            if c[k] > 0:
                ph.append("%s %.1f" % (names[k - 1], (c[k] - last) / 1965.0)); last = c[k]
        gap = (tr[ev, 0] - prev_end) / 1e3 if prev_end else 0.0
        print("  ev %2d  kernel %.1f us  gap %.1f us | %s" % (ev, (tr[ev, 11] - tr[ev, 0]) / 1e3, gap, "  ".join(ph)))
        prev_end = tr[ev, 11]
    cp = eg.chol_profile
    NB = (15 * (Oo + 1) + 6 + 7) // 8
    pp = cp[:4 * NB].reshape(NB, 4)
    print("  Cholesky of evaluation 1 inside k_step, cycles per panel [solve, own, diag, update_total]:", pp[:3].tolist(), "...", pp[-3:].tolist())
    print("  sums", pp.sum(0).tolist(), "backsub", int(cp[4 * NB]), "total %.1f us" % ((pp[:, 0].sum() + pp[:, 3].sum() + cp[4 * NB]) / 1965.0))
