"""Independent float64 numpy re-derivations that pin parts of the oracle the reference holds no fixture for (SURVEY 8c:
"the oracle is the pin, so it needs independent cross-checks"):
  * the Ceres-1.14 trust-region / TRADITIONAL_DOGLEG loop written at the JACOBIAN level (QR least squares of the augmented
    system for the regularised Gauss-Newton step, ||J s|| forms for the Cauchy point and the model cost change, per-residual
    corrector) against oracle/o_solver.cc, which works on the normal equations;
  * MarginalizationInfo::Marginalize (Schur complement through an eigen pseudo-inverse + eigen square root,
    MarginalizationFactor.cc:206-311) against numpy's pinv / eigh on a random well-conditioned problem;
  * stability properties of the fp32 front end: k-NN ties and voxel boundaries."""
import numpy as np
import pytest


# ---- 1. trust region, re-derived ------------------------------------------------------------------------------------------
def _residuals(A, B, y, amp, x):
    t = B @ x
    r = A @ x + amp * np.sin(t) - y
    J = A + (amp * np.cos(t))[:, None] * B
    return r, J


def _corrected(r, J, cauchy):
    """Per-residual robustification: cost 1/2 sum rho(r^2); CauchyLoss(1): rho' = 1/(1+s), rho'' < 0 -> scale r and J by sqrt(rho')."""
    if not cauchy:
        return 0.5 * float(r @ r), r, J
    s = r * r
    w = 1.0 / np.sqrt(1.0 + s)
    return 0.5 * float(np.log1p(s).sum()), r * w, J * w[:, None]


def numpy_ceres_dogleg(A, B, y, x0, amp, cauchy, max_iter):
    x = np.array(x0, float)
    cost, r, J = _corrected(*_residuals(A, B, y, amp, x), cauchy)
    initial_cost = cost
    scale = 1.0 / (1.0 + np.linalg.norm(J, axis=0))                      # jacobi_scaling, fixed at iteration 0
    gmax = np.abs(J.T @ r).max()
    J = J * scale
    radius, mu = 1e4, 1e-8
    it = succ = 0
    term = 0
    if gmax <= 1e-10:
        return x, dict(iterations=0, successful=0, termination=1, initial_cost=initial_cost, final_cost=cost)
    reuse = False
    invalid = 0
    x_norm = np.linalg.norm(x)
    while True:
        if it >= max_iter:
            term = 0
            break
        if radius < 1e-32:
            term = 1
            break
        it += 1
        if not reuse:
            reuse = True
            g = J.T @ r
            D = np.sqrt(np.clip((J * J).sum(0), 1e-6, 1e32))
            grad = g / D
            alpha = (grad @ grad) / np.linalg.norm(J @ (grad / D)) ** 2   # Cauchy point along the scaled gradient
            ok = False
            while mu < 1.0:
                # min || J z + r ||^2 + mu || D z ||^2 by QR on the augmented system (no normal equations)
                Ja = np.vstack([J, np.sqrt(mu) * np.diag(D)])
                z, *_ = np.linalg.lstsq(Ja, np.concatenate([-r, np.zeros(len(D))]), rcond=None)
                if np.isfinite(z).all():
                    gn = D * z
                    ok = True
                    break
                mu *= 10.0
        valid = ok
        if ok:
            gnorm, nnorm = np.linalg.norm(grad), np.linalg.norm(gn)
            if nnorm <= radius:
                step, snorm = gn.copy(), nnorm
            elif gnorm * alpha >= radius:
                step, snorm = -(radius / gnorm) * grad, radius
            else:
                a = -alpha * grad
                d = gn - a
                # || a + beta d || = radius, beta in (0, 1)
                qa, qb, qc = d @ d, 2 * (a @ d), a @ a - radius * radius
                beta = (-qb + np.sqrt(qb * qb - 4 * qa * qc)) / (2 * qa)
                step = a + beta * d
                snorm = np.linalg.norm(step)
            step = step / D
            Js = J @ step
            model = -float(Js @ (r + 0.5 * Js))
            valid = model > 0
        if not valid:
            invalid += 1
            if invalid >= 5:
                term = 2
                break
            mu *= 10.0
            reuse = False
            continue
        invalid = 0
        cand = x + step * scale
        ccost, cr, cJ = _corrected(*_residuals(A, B, y, amp, cand), cauchy)
        if np.linalg.norm(x - cand) <= 1e-8 * (x_norm + 1e-8):
            term = 1
            break
        change = cost - ccost
        if abs(change) <= 1e-6 * cost:
            term = 1
            break
        rel = change / model
        if rel > 1e-3:
            x, cost, r, J = cand, ccost, cr, cJ * scale
            x_norm = np.linalg.norm(x)
            succ += 1
            gmax = np.abs((J / scale).T @ r).max()
            if rel < 0.25:
                radius *= 0.5
            if rel > 0.75:
                radius = min(1e16, max(radius, 3.0 * snorm))
            mu = max(1e-8, 2.0 * mu / 10.0)
            reuse = False
            if gmax <= 1e-10:
                term = 1
                break
        else:
            radius *= 0.5
            reuse = True
    return x, dict(iterations=it, successful=succ, termination=term, initial_cost=initial_cost, final_cost=cost)


def _problem(seed, n, m, amp, noise):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(m, n))
    B = 0.3 * rng.normal(size=(m, n))
    xt = rng.normal(size=n)
    y = A @ xt + amp * np.sin(B @ xt) + noise * rng.normal(size=m)
    return A, B, y, xt


@pytest.mark.parametrize("seed,n,m,amp,noise,cauchy,max_iter,x0_sigma", [
    (0, 6, 30, 0.05, 0.0, False, 10, 0.5),
    (1, 21, 80, 0.8, 0.01, False, 10, 0.5),      # rejected steps: radius halves, dogleg interpolation between Cauchy and GN
    (2, 21, 80, 0.8, 0.5, True, 10, 0.5),        # robust loss: corrector path
    (5, 21, 80, 1.5, 0.0, False, 50, 0.5),
    (7, 12, 60, 2.5, 0.0, False, 30, 2.0),       # far start, strongly nonlinear: steepest-descent-limited steps
    (8, 15, 60, 1.0, 0.3, True, 1, 0.5),         # exactly one iteration: scaling + corrector + first radius update
])
def test_trust_region_loop_vs_numpy_jacobian_level(oracle, seed, n, m, amp, noise, cauchy, max_iter, x0_sigma):
    A, B, y, xt = _problem(seed, n, m, amp, noise)
    x0 = xt + x0_sigma * np.random.default_rng(seed + 100).normal(size=n)
    xo, so = oracle.toy_solve(A, B, y, x0, amp=amp, use_cauchy=cauchy, max_iter=max_iter)
    xn, sn = numpy_ceres_dogleg(A, B, y, x0, amp, cauchy, max_iter)
    assert abs(sn["initial_cost"] - so["initial_cost"]) <= 1e-12 * max(1.0, so["initial_cost"])
    assert (sn["iterations"], sn["successful"], sn["termination"]) == (so["iterations"], so["successful"], so["termination"]), (sn, so)
    assert abs(sn["final_cost"] - so["final_cost"]) <= 1e-7 * max(1e-12, so["final_cost"]) + 1e-16
    assert np.abs(xn - xo).max() <= 1e-6 * max(1.0, np.abs(xo).max())


# ---- 2. marginalisation algebra ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cauchy", [False, True])
def test_marginalize_schur_and_square_root_vs_numpy(oracle, cauchy):
    rng = np.random.default_rng(11)
    nb, bs, nr = 5, 3, 4
    pairs = [(0, 1), (0, 2), (1, 0), (0, 3), (2, 0), (0, 4), (1, 2), (3, 4), (0, 1), (4, 0), (0, 2), (3, 0)]   # (1,2), (3,4) do not touch block 0
    bi = np.array([p[0] for p in pairs]); bj = np.array([p[1] for p in pairs])
    W = rng.normal(size=(len(pairs), 2, nr, bs))
    x = rng.normal(size=(nb, bs))
    y = rng.normal(size=(len(pairs), nr)) * 0.3
    out = oracle.toy_marginalize(nb, bs, bi, bj, W, y, x, use_cauchy=cauchy)
    m, n = out["m"], out["n"]
    assert m == bs and n == (nb - 1) * bs and out["kept"].tolist() == [1, 2, 3, 4]
    # numpy: stack the factors that touch block 0 (block order: dropped first, then by address)
    order = [0, 1, 2, 3, 4]
    Jrows, rrows = [], []
    for k, (i, j) in enumerate(pairs):
        if i != 0 and j != 0:
            continue
        Jk = np.zeros((nr, nb * bs))
        Jk[:, order.index(i) * bs:(order.index(i) + 1) * bs] += W[k, 0]
        Jk[:, order.index(j) * bs:(order.index(j) + 1) * bs] += W[k, 1]
        rk = W[k, 0] @ x[i] + W[k, 1] @ x[j] - y[k]
        if cauchy:                                  # ResidualBlockInfo::Evaluate :69-95 (rho'' < 0 branch)
            w = 1.0 / np.sqrt(1.0 + rk @ rk)
            Jk, rk = Jk * w, rk * w
        Jrows.append(Jk); rrows.append(rk)
    Jn, rn = np.vstack(Jrows), np.concatenate(rrows)
    A, b = Jn.T @ Jn, Jn.T @ rn
    assert np.abs(out["A"] - A).max() <= 1e-12 * np.abs(A).max() and np.abs(out["b"] - b).max() <= 1e-12 * np.abs(b).max()
    Amm, Amr, Arr = A[:m, :m], A[:m, m:], A[m:, m:]
    A2 = Arr - Amr.T @ np.linalg.pinv(Amm, hermitian=True) @ Amr
    b2 = b[m:] - Amr.T @ np.linalg.pinv(Amm, hermitian=True) @ b[:m]
    J, r = out["J"], out["r"]
    assert np.abs(J.T @ J - A2).max() <= 1e-10 * np.abs(A2).max()            # prior information = Schur complement
    assert np.abs(J.T @ r - b2).max() <= 1e-9 * max(1.0, np.abs(b2).max())   # prior gradient
    # the square root is the eigen one: rows are sqrt(lambda_k) v_k^T with ascending lambda, zero rows below eps
    lam = np.linalg.eigvalsh(A2)
    assert np.allclose(np.sort((J * J).sum(1)), np.where(lam > 1e-8, lam, 0.0), rtol=1e-9, atol=1e-12)


def test_marginalize_drops_directions_below_eps(oracle):
    """A gauge direction (information only on x_1 - x_2) gives a singular Schur complement: the eigen square root keeps
    rank, the null direction gets an exactly-zero row and residual (eps = 1e-8 test, MarginalizationFactor.h)."""
    nb, bs, nr = 3, 2, 2
    pairs = [(0, 1), (0, 2)]
    I = np.eye(2)
    W = np.array([[I, -I], [I, -I]])            # r = x_0 - x_j - y
    x = np.array([[0.3, -0.2], [1.0, 2.0], [-1.0, 0.5]])
    y = np.zeros((2, 2))
    out = oracle.toy_marginalize(nb, bs, np.array([0, 0]), np.array([1, 2]), W, y, x)
    J, r = out["J"], out["r"]
    sq = (J * J).sum(1)
    assert (sq[:2] == 0).all() and (r[:2] == 0).all() and (sq[2:] > 0.1).all()
    # marginalising x_0 out of two relative constraints leaves 1/2 (x_1 - x_2 + const)^2 per axis
    d = np.array([1.0, 0.0, -1.0, 0.0])
    assert abs(d @ (J.T @ J) @ d - 2.0) < 1e-12 and np.abs((J.T @ J) @ np.array([1.0, 0, 1.0, 0])).max() < 1e-12


# ---- 3. fp32 front-end stability properties ------------------------------------------------------------------------------------
def test_knn_ties_resolve_by_index(oracle):
    """Duplicated map points are exact distance ties: the k-NN answer lists them by ascending index, independent of where
    the duplicates sit in the cloud (what lets the device deal candidates to lanes in any order)."""
    rng = np.random.default_rng(3)
    base = rng.uniform(-5, 5, (300, 3)).astype(np.float32)
    q = rng.uniform(-5, 5, (40, 3)).astype(np.float32)
    dup = np.concatenate([base, base[:120]], 0)                       # point i and 300 + i coincide
    perm = rng.permutation(len(dup))
    cloud = np.concatenate([dup[perm], np.zeros((len(dup), 1), np.float32)], 1)
    queries = np.concatenate([q, np.zeros((len(q), 1), np.float32)], 1)
    idx, d2 = oracle.knn(cloud, queries, 5)
    for i in range(len(q)):
        dd = ((cloud[:, :3] - q[i]) ** 2).astype(np.float32)
        dist = (dd[:, 0] + dd[:, 1]) + dd[:, 2]
        ref = np.lexsort((np.arange(len(cloud)), dist))[:5]
        assert np.array_equal(idx[i], ref), i
        assert (np.diff(d2[i]) >= 0).all()
        ties = np.nonzero(np.diff(d2[i]) == 0)[0]
        assert all(idx[i][t] < idx[i][t + 1] for t in ties)


def test_voxel_grid_is_invariant_to_input_order_up_to_float_sum(oracle):
    """Voxel membership is decided by floor((p - min) / leaf): points placed exactly on voxel boundaries and a permuted input
    give the same occupied voxels in the same (ascending index) output order; centroids differ only by summation order."""
    rng = np.random.default_rng(5)
    leaf = 0.4
    pts = rng.uniform(-3, 3, (2000, 3)).astype(np.float32)
    edge = (np.round(rng.uniform(-3, 3, (200, 3)) / leaf) * leaf).astype(np.float32)     # on the lattice of the leaf size
    cloud = np.concatenate([np.concatenate([pts, edge], 0), np.ones((2200, 1), np.float32)], 1)
    a = oracle.voxel_grid(cloud, leaf)
    b = oracle.voxel_grid(cloud[rng.permutation(len(cloud))], leaf)
    assert a.shape == b.shape
    assert np.abs(a[:, :3] - b[:, :3]).max() <= 2e-6 * 3.0
    # ascending voxel index == lexicographic (z, y, x) order of the voxel coordinates
    mn = cloud[:, :3].min(0)
    ijk = np.floor(a[:, :3] / leaf).astype(np.int64) - np.floor(mn / leaf).astype(np.int64)
    key = (ijk[:, 2] * 10000 + ijk[:, 1]) * 10000 + ijk[:, 0]
    assert (np.diff(key) > 0).all()


@pytest.mark.parametrize("seed,bscale,amp", [(20, 6.0, 2.0), (23, 6.0, 2.0), (25, 3.0, 3.0), (28, 3.0, 3.0), (30, 6.0, 2.0), (26, 6.0, 2.0)])
def test_trust_region_rejections_and_dogleg_interpolation(oracle, seed, bscale, amp):
    """High-frequency residuals: about half of the 50 iterations are rejected, the radius collapses from 1e4 to the step scale
    and the steps become Cauchy-limited / interpolated dogleg steps - the branches the smooth problems above never reach."""
    rng = np.random.default_rng(seed)
    n, m = 10, 50
    A = rng.normal(size=(m, n)); B = bscale * rng.normal(size=(m, n)); xt = rng.normal(size=n)
    y = A @ xt + amp * np.sin(B @ xt)
    x0 = xt + rng.normal(size=n)
    xo, so = oracle.toy_solve(A, B, y, x0, amp=amp, use_cauchy=False, max_iter=50)
    xn, sn = numpy_ceres_dogleg(A, B, y, x0, amp, False, 50)
    assert so["iterations"] - so["successful"] >= 10
    assert (sn["iterations"], sn["successful"], sn["termination"]) == (so["iterations"], so["successful"], so["termination"]), (sn, so)
    assert abs(sn["final_cost"] - so["final_cost"]) <= 1e-9 * so["final_cost"]
    assert np.abs(xn - xo).max() <= 1e-9 * max(1.0, np.abs(xo).max())
