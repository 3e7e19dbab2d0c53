"""Pieces of the device-resident solver in isolation (GPU): the tiled shared-memory Cholesky vs LAPACK."""
import ctypes as C

import numpy as np
import pytest

from lio_mapping_b200 import _lib

pytestmark = pytest.mark.gpu


def _solve(A, b):
    n = A.shape[0]
    x = np.zeros(n)
    ok = C.c_int()
    _lib.check(_lib.lib().lio_dev_cholesky_solve_host(np.ascontiguousarray(A), np.ascontiguousarray(b), n, x, C.byref(ok), None, 0), "chol")
    return x, ok.value


@pytest.mark.parametrize("n", [1, 7, 8, 9, 36, 96, 171, 201, 216])
def test_tiled_cholesky_vs_lapack(n):
    rng = np.random.default_rng(n)
    J = rng.normal(size=(3 * n + 5, n))
    A = J.T @ J
    d = 1.0 / (1.0 + np.sqrt(np.diag(A)))        # Jacobi-scaled like the solver's systems
    A = A * d[:, None] * d[None, :]
    b = rng.normal(size=n)
    x, ok = _solve(A, b)
    assert ok == 1
    xr = np.linalg.solve(A, b)
    assert np.abs(x - xr).max() <= 1e-10 * max(1.0, np.abs(xr).max()) * np.linalg.cond(A) ** 0.5


def test_tiled_cholesky_ill_conditioned_window_like():
    """Entries spread over 13 decades like a window problem before scaling is applied to its diagonal only."""
    rng = np.random.default_rng(5)
    n = 171
    s = 10.0 ** rng.uniform(-3, 3.5, n)
    J = rng.normal(size=(4 * n, n)) * s[None, :]
    A = J.T @ J
    d = 1.0 / (1.0 + np.sqrt(np.diag(A)))
    A = A * d[:, None] * d[None, :] + 1e-8 * np.eye(n)
    b = rng.normal(size=n)
    x, ok = _solve(A, b)
    assert ok == 1
    r = A @ x - b
    assert np.abs(r).max() <= 1e-9 * max(1.0, np.abs(b).max(), np.abs(A).max() * np.abs(x).max())


def test_tiled_cholesky_reports_indefinite():
    n = 40
    A = np.eye(n)
    A[17, 17] = -1.0
    _, ok = _solve(A, np.ones(n))
    assert ok == 0
