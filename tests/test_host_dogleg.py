"""The product's dogleg controller (solver_host.cc, stand-in for ceres::Solve as the reference configures it) against the
oracle's independent Ceres-1.14 restatement (oracle/o_solver.cc: per-residual-block Problem, corrector, trust-region
loop) on toy nonlinear least-squares problems — host only, same iteration-level decisions and the same minimiser."""
import numpy as np
import pytest

from lio_mapping_b200 import _lib


def _problem(seed, n, m, amp, noise):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(m, n))
    B = 0.3 * rng.normal(size=(m, n))
    xt = rng.normal(size=n)
    y = A @ xt + amp * np.sin(B @ xt) + noise * rng.normal(size=m)
    return A, B, y, xt


def _ours(A, B, y, x0, amp, cauchy, max_iter):
    m, n = A.shape
    x = np.ascontiguousarray(x0, np.float64).copy()
    s = np.zeros(8)
    _lib.check(_lib.lib().lio_host_dogleg_toy(n, m, np.ascontiguousarray(A), np.ascontiguousarray(B), np.ascontiguousarray(y), amp,
                                              int(cauchy), x, max_iter, s), "dogleg_toy")
    return x, dict(iterations=int(s[0]), successful=int(s[1]), termination=int(s[2]), initial_cost=s[3], final_cost=s[4])


@pytest.mark.parametrize("seed,n,m,amp,noise,cauchy,max_iter", [
    (0, 6, 30, 0.05, 0.0, False, 10),       # nearly linear, exact data: converges by tolerance
    (1, 21, 80, 0.8, 0.01, False, 10),      # strongly nonlinear: rejected steps, radius shrinks
    (2, 21, 80, 0.8, 0.5, True, 10),        # robust loss with outlier-sized noise: corrector path
    (3, 96, 200, 0.3, 0.05, True, 10),      # window-sized system
    (4, 171, 342, 0.05, 0.01, False, 10),   # n of the HDL-64 window (blocked Cholesky path)
    (5, 21, 80, 1.5, 0.0, False, 50),       # long run to a tolerance exit
])
def test_controller_matches_oracle_solver(oracle, seed, n, m, amp, noise, cauchy, max_iter):
    A, B, y, xt = _problem(seed, n, m, amp, noise)
    x0 = xt + 0.5 * np.random.default_rng(seed + 100).normal(size=n)
    xo, so = oracle.toy_solve(A, B, y, x0, amp=amp, use_cauchy=cauchy, max_iter=max_iter)
    xg, sg = _ours(A, B, y, x0, amp, cauchy, max_iter)
    assert abs(sg["initial_cost"] - so["initial_cost"]) <= 1e-12 * max(1.0, so["initial_cost"])
    assert sg["iterations"] == so["iterations"] and sg["successful"] == so["successful"], (sg, so)
    assert sg["termination"] == so["termination"]
    assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-9 * max(1e-12, so["final_cost"]) + 1e-18
    assert np.abs(xg - xo).max() <= 1e-8 * max(1.0, np.abs(xo).max())
    assert sg["final_cost"] <= sg["initial_cost"]


def test_controller_handles_a_singular_direction(oracle):
    """Rank-deficient J^T J (two identical columns): the regularised solve (mu D^2) must still take descent steps."""
    A, B, y, xt = _problem(9, 12, 40, 0.0, 0.0)
    A[:, 5] = A[:, 4]; B[:, 5] = B[:, 4]
    y = A @ xt
    x0 = np.zeros(12)
    xo, so = oracle.toy_solve(A, B, y, x0, amp=0.0, max_iter=10)
    xg, sg = _ours(A, B, y, x0, 0.0, False, 10)
    assert sg["iterations"] == so["iterations"] and sg["termination"] == so["termination"]
    assert sg["final_cost"] <= 1e-12 * max(1.0, sg["initial_cost"]) + 1e-9 and so["final_cost"] <= 1e-9
    assert np.abs(A @ xg - y).max() < 1e-5
