"""Dense fp64 kernels of the host shell (blocked Cholesky, symmetric eigen-solver) against LAPACK (numpy) — the sizes the
window solver uses: n = 15 (O + 1) [+ 6] for the dogleg step, 15 O + 6 for the marginalisation.  Host only."""
import ctypes as C

import numpy as np
import pytest

from lio_mapping_b200 import _lib


def _spd(rng, n, cond=1e6):
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    w = np.exp(rng.uniform(0, np.log(cond), n))
    return (q * w) @ q.T


@pytest.mark.parametrize("n", [1, 6, 15, 47, 48, 49, 96, 165, 171, 261])
def test_cholesky_solve_vs_lapack(n):
    rng = np.random.default_rng(n)
    A = _spd(rng, n)
    A = 0.5 * (A + A.T)
    b = rng.normal(size=n)
    L = np.zeros((n, n)); x = np.zeros(n)
    _lib.check(_lib.lib().lio_host_cholesky_solve(n, np.ascontiguousarray(A), b, L, x), "chol")
    Lr = np.linalg.cholesky(A)
    assert np.abs(L - Lr).max() <= 1e-10 * np.abs(Lr).max()
    xr = np.linalg.solve(A, b)
    assert np.abs(x - xr).max() <= 1e-8 * max(1.0, np.abs(xr).max())
    assert np.abs(A @ x - b).max() <= 1e-9 * max(1.0, np.abs(b).max()) * np.linalg.cond(A) ** 0.5


def test_cholesky_rejects_indefinite():
    A = np.eye(80); A[40, 40] = -1.0
    rc = _lib.lib().lio_host_cholesky_solve(80, A, np.ones(80), np.zeros((80, 80)), np.zeros(80))
    assert rc == -5                                     # LIO_ERR_NUMERIC: the dogleg controller raises mu and retries
    A = np.eye(80); A[79, 79] = np.nan
    assert _lib.lib().lio_host_cholesky_solve(80, A, np.ones(80), np.zeros((80, 80)), np.zeros(80)) == -5


@pytest.mark.parametrize("n,threads", [(1, 1), (3, 1), (15, 1), (33, 2), (96, 1), (156, 1), (156, 4), (171, 3)])
def test_sym_eigen_vs_lapack(n, threads):
    rng = np.random.default_rng(100 + n)
    # marginalisation-like spectrum: a few near-null directions next to entries ~1e13
    w = np.concatenate([np.zeros(min(4, n - 1)), np.exp(rng.uniform(0, 30, n - min(4, n - 1)))]) if n > 1 else np.array([2.5])
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    A = (q * w) @ q.T
    A = 0.5 * (A + A.T)
    ev = np.zeros(n); V = np.zeros((n, n))
    _lib.check(_lib.lib().lio_host_sym_eigen(n, np.ascontiguousarray(A), ev, V, threads), "eig")
    wr = np.linalg.eigvalsh(A)
    scale = np.abs(wr).max()
    assert np.all(np.diff(ev) >= 0)                                   # ascending, like SelfAdjointEigenSolver
    assert np.abs(ev - wr).max() <= 1e-13 * scale * n
    assert np.abs(V.T @ V - np.eye(n)).max() <= 1e-12 * n             # orthonormal columns
    assert np.abs((V * ev) @ V.T - A).max() <= 1e-13 * scale * n      # reconstruction


def test_sym_eigen_threads_bit_identical():
    rng = np.random.default_rng(7)
    B = rng.normal(size=(156, 156)); A = B @ B.T
    out = []
    for th in (1, 2, 4):
        ev = np.zeros(156); V = np.zeros((156, 156))
        _lib.check(_lib.lib().lio_host_sym_eigen(156, A, ev, V, th), "eig")
        out.append((ev, V))
    for ev, V in out[1:]:
        assert np.array_equal(ev, out[0][0]) and np.array_equal(V, out[0][1])
