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
# columns: [0] panel solve + barrier (chain warp's clock64, read behind the barrier), [1] %globaltimer (ns) when the chain warp starts the
# next diagonal tile, [2] diagonal tile factor + inverse (cycles), [3] update phase incl. the wait for the workers (cycles)
wall = np.diff(p[:NB - 1, 1])
print("panel: solve(cyc) diag(cyc) update_total(cyc) | wall to the next panel (ns)")
for k in range(NB):
    print(k, int(p[k, 0]), int(p[k, 2]), int(p[k, 3]), "|", int(wall[k]) if k < len(wall) else "-")
cyc = int(p[:, 0].sum() + p[:, 3].sum() + prof[4 * NB])
print("chain-warp cycles: loop", int(p[:, 0].sum() + p[:, 3].sum()), "backsub", int(prof[4 * NB]), "total", cyc, "= %.1f us at 1.965 GHz (wall by %%globaltimer: see the [k_chol_test] lines on stderr)" % (cyc / 1965.0))
