"""GPU parity where the linear algebra is hard:
  * an ill-conditioned lasso (correlated columns, rho = 1e-3, cond(D'D + rho I) ~ 4e5): the inverse-factor
    x-update must keep the 1e-9 parity, and must agree with blocked substitution on the same factor;
  * the conditioning guard: a factor whose diagonal spans > 1e9 switches the x-update to the reference's own
    two substitutions (getProxOps.m:1200);
  * rank-deficient D for the unwrapped x-update: all-zero columns (constant-zero MNIST pixels) give x_j = 0 exactly
    as pinv(D) does (unwrappedadmm.m:76-78, linearsvm.m:185); any other rank deficiency raises and says why."""
import os

import numpy as np
import pytest
import scipy.linalg as sla

import oracle
from admm_project_b200 import EngineError, lasso, linearsvm, linearsvm_onevsall
from admm_project_b200 import _lib as L
from admm_project_b200 import generators as gen

pytestmark = pytest.mark.gpu
TOL = 1e-9
HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def correlated_design(seed, m, n, corr):
    rs = np.random.RandomState(seed)
    D = np.sqrt(1 - corr) * rs.randn(m, n) + np.sqrt(corr) * rs.randn(m, 1)
    D /= np.linalg.norm(D, axis=0)
    xt = rs.randn(n) * (rs.rand(n) < 0.3)
    s = D @ xt + 0.03 * rs.randn(m)
    return np.asfortranarray(D), s


@pytest.mark.parametrize("corr,rho", [(0.99, 1e-3), (0.999, 1e-3)])
def test_ill_conditioned_lasso_keeps_parity(engine, corr, rho):
    D, s = correlated_design(0, 2048, 512, corr)
    lam = 0.01 * float(np.max(np.abs(D.T @ s)))
    opts = {"rho": rho, "maxiters": 300, "history": 0}
    ref = oracle.lasso(D, s, lam, opts)
    res = lasso(D, s, lam, opts, engine=engine)
    assert res["steps"] == ref["steps"]
    for k in ("xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr"):
        assert rel(res[k], ref[k]) < TOL, (k, rel(res[k], ref[k]))
    info = engine.info()
    assert info["xsolve_effective"] == L.XSOLVE_INVFACTOR and info["diag_ratio"] > 1e2
    # both x-update realisations against SciPy's substitutions on the ENGINE's factor
    Lf = engine.get_factor()
    b = np.random.RandomState(1).randn(512)
    exact = sla.solve_triangular(Lf.T, sla.solve_triangular(Lf, b, lower=True), lower=False)
    assert rel(engine.factor_solve(b, L.XSOLVE_INVFACTOR), exact) < 1e-10
    assert rel(engine.factor_solve(b, L.XSOLVE_SUBST), exact) < 1e-10
    G = D.T @ D + rho * np.eye(512)
    assert rel(Lf, np.linalg.cholesky(G)) < 1e-10


def test_conditioning_guard_switches_to_substitution(engine):
    rs = np.random.RandomState(3)
    m, n = 1024, 256
    D = rs.randn(m, n) * (10.0 ** (-6.0 * np.arange(n) / n))[None, :]     # column scales span 1e6: diag(L) spans > 3e4
    s = D @ rs.randn(n)
    lam = 1e-3 * float(np.max(np.abs(D.T @ s)))
    opts = {"rho": 1e-13, "maxiters": 50, "domaxiters": 1, "history": 0}
    res = lasso(D, s, lam, opts, engine=engine)
    info = engine.info()
    assert info["diag_ratio"] > 1e9 and info["xsolve_effective"] == L.XSOLVE_SUBST
    ref = oracle.lasso(D, s, lam, opts)
    assert res["steps"] == ref["steps"] == 50
    # cond ~ 1e12: two Cholesky implementations agree to ~cond*eps only; the point here is the switch
    assert rel(res["pnorm"][:5], ref["pnorm"][:5]) < 1e-3
    # the guard is a property of the setup: a well-conditioned problem afterwards is back on the inverse factor
    D2, s2, lam2, _ = gen.lasso_problem(0, 256, 64)
    lasso(D2, s2, lam2, {}, engine=engine)
    assert engine.info()["xsolve_effective"] == L.XSOLVE_INVFACTOR


def mnist_like_with_dead_pixels(rows):
    """20 x 20 crops (400 features, mnistsvm.m:237) whose border pixels are always zero, labels = the first
    `rows` entries of the reference's own train-labels file (tests/golden/make_mnist_labels.py)."""
    labels = np.load(os.path.join(HERE, "golden", "mnist_train_labels_6000.npy")).astype(np.int64)[:rows]
    D, ELL = gen.svm_mnist_like(7, rows, 400, labels=labels)
    dead = np.zeros((20, 20), dtype=bool)
    dead[:2, :] = dead[-2:, :] = dead[:, :2] = dead[:, -2:] = True
    D[:, dead.reshape(-1)] = 0.0
    return D, ELL, dead.reshape(-1)


def test_zero_columns_match_pinv_semantics(engine):
    D, ELL, dead = mnist_like_with_dead_pixels(6000)
    assert np.linalg.matrix_rank(D) == 400 - dead.sum()
    opts = {"objevals": 1, "history": 0}
    np.random.seed(11)
    ref = oracle.linearsvm(D, ELL[:, 4], 0.5, opts)        # serial reference path: Dplus = pinv(D) (linearsvm.m:185)
    np.random.seed(11)
    res = linearsvm(D, ELL[:, 4], 0.5, opts, engine=engine)
    assert engine.info()["zero_cols"] == int(dead.sum())
    assert res["steps"] == ref["steps"]
    for k in ("xopt", "zopt", "uopt", "pnorm", "perr", "objevals"):
        assert rel(res[k], ref[k]) < TOL, (k, rel(res[k], ref[k]))
    assert np.all(res["xopt"][dead] == 0.0)                # the minimum-norm solution: exactly zero, not merely small


def test_zero_columns_in_the_class_batch(engine):
    D, ELL, dead = mnist_like_with_dead_pixels(3000)
    np.random.seed(12)
    out = linearsvm_onevsall(D, ELL[:, :4], 0.5, {"objevals": 1}, engine=engine)
    np.random.seed(12)
    for k in range(4):
        ref = oracle.linearsvm(D, ELL[:, k], 0.5, {"objevals": 1, "history": 0})
        assert out[k]["steps"] == ref["steps"]
        assert rel(out[k]["xopt"], ref["xopt"]) < TOL and np.all(out[k]["xopt"][dead] == 0.0)


def test_other_rank_deficiency_raises_with_the_reference_context(engine):
    D, ELL = gen.svm_mnist_like(8, 500, 40)
    D[:, 7] = D[:, 3] + D[:, 5]                            # collinear columns: D'D singular, no zero column
    with pytest.raises(EngineError, match="pinv"):
        linearsvm(D, ELL[:, 0], 0.5, {}, engine=engine)
