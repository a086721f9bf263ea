"""GPU parity of the building-block kernels (through the C-ABI) against NumPy/SciPy FP64."""
import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 257, 1001), (5, 3, 7), (129, 640, 33)])
def test_dgemm_matches_numpy(engine, ta, tb, M, N, K):
    rs = np.random.RandomState(M * 7 + N * 3 + K + ta * 2 + tb)
    A = rs.randn(K, M) if ta else rs.randn(M, K)
    B = rs.randn(N, K) if tb else rs.randn(K, N)
    C0 = rs.randn(M, N)
    ref = 1.5 * (A.T if ta else A) @ (B.T if tb else B) - 0.5 * C0
    got = engine.dgemm(ta, tb, 1.5, A, B, beta=-0.5, Cmat=C0)
    assert rel(got, ref) < 1e-13


def test_dgemm_splitk_and_lower_only(engine):
    rs = np.random.RandomState(5)
    A = rs.randn(20000, 200)
    ref = A.T @ A
    got = engine.dgemm(1, 0, 1.0, A, A)                     # few tiles, long K -> split-K path
    assert rel(got, ref) < 1e-13
    got2 = engine.dgemm(1, 0, 1.0, A, A)
    assert np.array_equal(got, got2)                        # deterministic reduction order
    A = rs.randn(500, 700)
    low = engine.dgemm(1, 0, 1.0, A, A, lower_only=True)
    ref = A.T @ A
    assert rel(np.tril(low), np.tril(ref)) < 1e-13


@pytest.mark.parametrize("m,n", [(1000, 300), (257, 129), (150, 500)])
def test_gram_both_orientations(engine, m, n):
    rs = np.random.RandomState(m + n)
    D = rs.randn(m, n)
    G = engine.gram(D, trans=True, scale=1.0, shift=0.75)
    assert rel(G, D.T @ D + 0.75 * np.eye(n)) < 1e-13
    G = engine.gram(D, trans=False, scale=0.5, shift=1.0)
    assert rel(G, 0.5 * D @ D.T + np.eye(m)) < 1e-13


@pytest.mark.parametrize("k", [1, 7, 128, 129, 512, 513, 700, 1024, 1500, 2048, 2500, 4100])   # > 512: look-ahead driver
def test_potrf_and_inverse_factor(engine, k):
    rs = np.random.RandomState(k)
    X = rs.randn(k + 50, k)
    A = X.T @ X + np.eye(k)
    Lref = np.linalg.cholesky(A)
    Lg, W = engine.potrf(A, want_inverse=True)
    assert np.all(np.triu(Lg, 1) == 0)
    assert rel(Lg, Lref) < 1e-12                             # SURVEY.md section 7 step 5 gate
    assert np.all(np.triu(W, 1) == 0)
    assert rel(W @ Lref, np.eye(k)) < 1e-11


def test_potrf_not_positive_definite_fails_loudly(engine):
    from admm_project_b200 import EngineError
    A = np.eye(200)
    A[150, 150] = -1.0
    with pytest.raises(EngineError) as e:
        engine.potrf(A)
    assert "positive definite" in str(e.value)


@pytest.mark.parametrize("m,n", [(600, 200), (100, 333), (3000, 1027)])
def test_lasso_setup_factor_and_solve(engine, m, n):
    rs = np.random.RandomState(m)
    D = rs.randn(m, n) / np.sqrt(m)
    s = rs.randn(m)
    rho = 1.3
    engine.setup_lasso(D, s, rho)
    Lg = engine.get_factor()
    G = D.T @ D + rho * np.eye(n) if m >= n else D @ D.T / rho + np.eye(m)
    Lref = np.linalg.cholesky(G)
    assert rel(Lg, Lref) < 1e-12
    b = rs.randn(Lref.shape[0])
    xref = sla.solve_triangular(Lref.T, sla.solve_triangular(Lref, b, lower=True), lower=False)
    assert rel(engine.factor_solve(b), xref) < 1e-12


@pytest.mark.parametrize("m,n", [(900, 300), (200, 515), (2000, 1027)])
def test_substitution_mode_equals_inverse_factor_mode(engine, m, n):
    """ADMM_B200_XSOLVE_SUBST (blocked forward/back substitution on L) and the default inverse-factor
    products give the same x-update; both match SciPy's triangular solves."""
    from admm_project_b200 import _lib as L
    rs = np.random.RandomState(n)
    D = rs.randn(m, n) / np.sqrt(m)
    s = rs.randn(m)
    engine.setup_lasso(D, s, 0.7)
    Lref = engine.get_factor()
    b = rs.randn(Lref.shape[0])
    xref = sla.solve_triangular(Lref.T, sla.solve_triangular(Lref, b, lower=True), lower=False)
    x_inv = engine.factor_solve(b, L.XSOLVE_INVFACTOR)
    x_sub = engine.factor_solve(b, L.XSOLVE_SUBST)
    assert rel(x_inv, xref) < 1e-12 and rel(x_sub, xref) < 1e-12
    assert rel(x_inv, x_sub) < 1e-12


def test_lasso_runs_in_substitution_mode(engine):
    import oracle
    from admm_project_b200 import lasso
    from admm_project_b200.generators import lasso_problem
    from admm_project_b200 import _lib as L
    for rows, cols in ((600, 260), (150, 400), (3000, 1300)):      # 1300: three 512-wide steps (512, 512, 276) per solve
        D, s, lam, _ = lasso_problem(0, rows, cols)
        ref = oracle.lasso(D, s, lam, {"history": 0})
        res = lasso(D, s, lam, {"history": 0, "xsolve": L.XSOLVE_SUBST}, engine=engine)
        assert res["steps"] == ref["steps"]
        assert rel(res["xopt"], ref["xopt"]) < 1e-9 and rel(res["uopt"], ref["uopt"]) < 1e-9
