#!/usr/bin/env python
"""Per-panel cycle profile of the device solver's tiled Cholesky (GPU box only): python scripts/chol_profile.py [n]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lio_mapping_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 171
rng = np.random.default_rng(1)
J = rng.normal(size=(3 * n, n)); A = J.T @ J
d = 1 / (1 + np.sqrt(np.diag(A))); A = A * d[:, None] * d[None, :]
b = rng.normal(size=n); x = np.zeros(n); ok = C.c_int()
NB = (n + 7) // 8
prof = np.zeros(4 * NB + 1, np.int64)
L = _lib.lib()
for _ in range(3):
    _lib.check(L.lio_dev_cholesky_solve_host(A, b, n, x, C.byref(ok), prof.ctypes.data_as(C.c_void_p), 0), "chol")
p = prof[:4 * NB].reshape(NB, 4)
print("n", n, "ok", ok.value, "err", np.abs(x - np.linalg.solve(A, b)).max())
print("panel: solve own_update diag update_total(cycles)")
for k in range(NB):
    print(k, p[k].tolist())
print("sum", p.sum(0).tolist(), "backsub", int(prof[4 * NB]), "total cycles", int(p[:, 0].sum() + p[:, 3].sum() + prof[4 * NB]), "= %.1f us" % ((p[:, 0].sum() + p[:, 3].sum() + prof[4 * NB]) / 1965.0))
