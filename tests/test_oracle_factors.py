"""fp64 oracle pinned by independent math: numeric-diff Jacobians (the reference's Check() convention,
src/factor/PivotPointPlaneFactor.cc:139-239), the reference IMU fixture, LAPACK cross-checks."""
import os

import numpy as np
import pytest

from lio_mapping_b200 import synth
from tests import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imu_pose_vel_10s.npz")


def rand_pose(rng, scale=5.0):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    return np.concatenate([rng.uniform(-scale, scale, 3), q])


def test_ppp_jacobian_numeric(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        x0, xi, xe = rand_pose(rng), rand_pose(rng), rand_pose(rng, 0.5)
        p = rng.uniform(-20, 20, 3); w = rng.normal(size=3); w /= np.linalg.norm(w); coeff = np.r_[0.8 * w, rng.normal()]
        r, J = oracle.ppp_evaluate(p, coeff, x0, xi, xe)
        eps = 1e-6
        for b, x in enumerate([x0, xi, xe]):
            for k in range(6):
                d = np.zeros(6); d[k] = eps
                xs = [x0, xi, xe]
                xs[b] = oracle.pose_plus(x, d)          # p + dp, q * DeltaQ(dtheta) normalised
                r2, _ = oracle.ppp_evaluate(p, coeff, *xs)
                assert abs((r2 - r) / eps - J[b][k]) < 5e-4 * max(1.0, abs(J[b][k])), (b, k)
            assert J[b][6] == 0.0


def test_ppp_rank6_structure(oracle):
    """J (1x18) factors through g = [a, p x a] with a = R_lpi^T w."""
    rng = np.random.default_rng(1)
    x0, xi, xe = rand_pose(rng), rand_pose(rng), rand_pose(rng, 0.5)
    R0, Ri, Re = (synth.quat_to_rot(x[3:]) for x in (x0, xi, xe))
    Rlp, Rli = R0 @ Re.T, Ri @ Re.T
    Plp, Pli = x0[:3] - Rlp @ xe[:3], xi[:3] - Rli @ xe[:3]
    Rlpi, Plpi = Rlp.T @ Rli, Rlp.T @ (Pli - Plp)
    G, JJ, res = [], [], []
    for _ in range(40):
        p = rng.uniform(-20, 20, 3); w = rng.normal(size=3); b = rng.normal()
        r, J = oracle.ppp_evaluate(p, np.r_[w, b], x0, xi, xe)
        a = Rlpi.T @ w
        G.append(np.r_[a, np.cross(p, a)])
        JJ.append(np.concatenate([j[:6] for j in J]))
        res.append(r - (a @ p + w @ Plpi + b))
    G, JJ = np.array(G), np.array(JJ)
    M, resid, rank, _ = np.linalg.lstsq(G, JJ, rcond=None)
    assert rank == 6
    assert np.abs(G @ M - JJ).max() < 1e-10 * np.abs(JJ).max()
    assert np.abs(res).max() < 1e-11


def test_imu_fixture_kat(oracle):
    """IntegrationBase on the reference fixture reproduces its ground-truth motion (what the reference's
    pimtest logs, test/test_imu_processor/test_imu_factor.cc:196-330); ImuFactor residual ~ 0 at ground truth."""
    g = np.load(GOLD)
    t, q, p, v, gyro, acc = g["t"], g["q_wxyz"], g["p"], g["v"], g["gyro"], g["acc"]
    gvec = np.array([0, 0, -9.81])
    INTERVAL = 20
    for start in [0, 400, 1200]:
        pim = oracle.Pim(acc[start], gyro[start], np.zeros(3), np.zeros(3), g_norm=9.81)
        for j in range(start + 1, start + INTERVAL + 1):
            pim.push_back(t[j] - t[j - 1], acc[j], gyro[j])
        s = pim.get()
        i, j = start, start + INTERVAL
        dt = t[j] - t[i]
        assert abs(s["sum_dt"] - dt) < 1e-12
        qi = np.r_[q[i, 1:], q[i, 0]]; qj = np.r_[q[j, 1:], q[j, 0]]
        Ri, Rj = synth.quat_to_rot(qi), synth.quat_to_rot(qj)
        dp_gt = Ri.T @ (p[j] - p[i] - v[i] * dt - 0.5 * gvec * dt * dt)
        dv_gt = Ri.T @ (v[j] - v[i] - gvec * dt)
        dR_gt = Ri.T @ Rj
        assert np.abs(s["delta_p"] - dp_gt).max() < 2e-4
        assert np.abs(s["delta_v"] - dv_gt).max() < 2e-3
        assert np.abs(synth.quat_to_rot(s["delta_q"]) - dR_gt).max() < 1e-4
        assert np.allclose(s["covariance"], s["covariance"].T, atol=1e-18)
        assert np.all(np.linalg.eigvalsh(s["covariance"]) > -1e-18)
        r, _ = pim.imu_factor(np.r_[p[i], qi], np.r_[v[i], np.zeros(6)], np.r_[p[j], qj], np.r_[v[j], np.zeros(6)])
        assert np.linalg.norm(r) < 5.0            # whitened residual: a few sigma (mid-point discretisation) at ground truth


def test_imu_factor_jacobian_numeric(oracle):
    rng = np.random.default_rng(2)
    traj = synth.Trajectory(ax=3.0, ay=2.0, az=0.1, period=20.0)
    tt, acc, gyr = synth.make_imu(traj, 1.0, 1.1)
    _, _, _, g0, a0 = traj.state(np.array(1.0))
    pim = oracle.Pim(a0, g0, np.array([0.01, -0.02, 0.03]), np.array([0.001, 0.002, -0.001]), acc_n=0.2, gyr_n=0.02)
    last = 1.0
    for j in range(len(tt)):
        pim.push_back(tt[j] - last, acc[j], gyr[j]); last = tt[j]
    pi, pj = rand_pose(rng, 2.0), rand_pose(rng, 2.0)
    sbi, sbj = rng.normal(0, 0.1, 9), rng.normal(0, 0.1, 9)
    r, J = pim.imu_factor(pi, sbi, pj, sbj)
    eps = 1e-6
    blocks = [pi, sbi, pj, sbj]
    for b in range(4):
        n = 6 if b in (0, 2) else 9
        for k in range(n):
            d = np.zeros(n); d[k] = eps
            xs = list(blocks)
            xs[b] = oracle.pose_plus(blocks[b], d) if b in (0, 2) else blocks[b] + d
            r2, _ = pim.imu_factor(*xs)
            num = (r2 - r) / eps
            scale = max(1.0, np.abs(J[b][:, k]).max())
            assert np.abs(num - J[b][:, k]).max() < 2e-3 * scale, (b, k)


def test_sym_eigen_vs_lapack(oracle):
    rng = np.random.default_rng(3)
    for n, rank in [(6, 6), (15, 15), (156, 156), (156, 150)]:
        B = rng.normal(size=(rank, n))
        A = B.T @ B * 10.0
        ev, V = oracle.sym_eigen(A)
        ref = np.linalg.eigvalsh(A)
        assert np.all(np.diff(ev) >= -1e-9)
        assert np.allclose(ev, ref, atol=1e-9 * max(1.0, ref.max()))
        assert np.abs(V @ np.diag(ev) @ V.T - A).max() < 1e-9 * ref.max()
        assert np.abs(V.T @ V - np.eye(n)).max() < 1e-10


@pytest.fixture(scope="module")
def vlp_seq(oracle):
    return helpers.Sequence(oracle, "vlp16", n_total=9, distort=False)


def test_estimator_steady_state(oracle, vlp_seq):
    """End-to-end oracle window solve on the synthetic VLP-16 sequence: converges, stays near ground truth."""
    W = 5
    est = oracle.Estimator(window_size=W, opt_window_size=W, opt_extrinsic=0)
    helpers.warm_start(est, vlp_seq, oracle, W, pose_noise=0.01, seed=1,
                       make_pim=lambda a, g: oracle.Pim(a, g, np.zeros(3), np.zeros(3), acc_n=0.2, gyr_n=0.02))
    for k in range(W, 9):
        helpers.feed_scan(est, vlp_seq, k)
        s = est.summary()
        assert s["num_features"] > 1000
        assert s["final_cost"] <= s["initial_cost"] * (1 + 1e-9)
        assert np.isfinite(s["final_cost"])
        st = est.states()
        # after SlideWindow the newest optimised frame sits at index W-1 (and is duplicated at W)
        err = np.linalg.norm(st[W - 1, 0:3] - vlp_seq.gt_p[k])
        assert err < 0.05, (k, err, s)
    assert est.summary()["has_prior"] == 1.0
