"""CPU tests of the oracle (the checker) itself.  The reference is MATLAB and ships no golden
vectors, so the oracle is PARITY UNPINNED (oracle/__init__.py); what can be pinned is pinned here:
hand-derived known answers of each operator, closed-form optima, KKT conditions of the problems
the solvers claim to solve, the reference testers' own pass criteria (testers/*.m), and the quirks
of admm.m that SURVEY.md section 5 lists."""
import math

import numpy as np
import pytest
import scipy.optimize as sopt

import oracle
from admm_project_b200 import generators as gen


# ---- operators: hand-derived values ---------------------------------------------------------
def test_soft_threshold_known_values():                      # getProxOps.m:933-938
    v = np.array([-3.0, -1.0, -0.5, 0.0, 0.5, 1.0, 3.0])
    assert np.array_equal(oracle.zminSoftThresholding(v, 1.0), [-2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 2.0])


def test_minz01_known_values():                              # getProxOps.m:1158-1180
    s = np.array([2.0, 1.0, 0.99, 0.0, -0.5, -3.0])
    # t = rho/C = 2 -> threshold 1 - sqrt(2/2) = 0: s >= 1 or s < 0 keep s, the rest become 1
    assert np.array_equal(oracle.minz01(s, 2.0), [2.0, 1.0, 1.0, 1.0, -0.5, -3.0])


def test_box_and_nonneg():                                   # getProxOps.m:1378-1382, 1470-1474
    x, u = np.array([-2.0, 0.5, 3.0]), np.array([0.5, 0.5, 0.5])
    assert np.array_equal(oracle.zminNonNegative(x, None, u, 1.0), [0.0, 1.0, 3.5])
    box = oracle.make_zminBox(np.array([-1.0, -1.0, -1.0]), np.array([1.0, 0.75, 1.0]))
    assert np.array_equal(box(x, None, u, 1.0), [-1.0, 0.75, 1.0])


def test_huber_prox_is_the_minimiser():                      # getProxOps.m:1529-1539
    # z = argmin 1/2*huber(z) + rho/2 (z - v)^2 checked by brute force on a grid
    rho, s = 0.7, np.zeros(1)
    _, minz, _ = oracle.getproxops("huberfit", dict(R=np.eye(1), D=np.eye(1), s=s))
    for v in (-5.0, -0.3, 0.0, 0.9, 4.0):
        z = minz(np.array([v]), None, np.zeros(1), rho)[0]
        grid = np.linspace(-6, 6, 240001)
        f = 0.5 * oracle.huber(grid) / 2 + rho / 2 * (grid - v) ** 2
        # huberfit's objective is 1/2*sum(huber(z)) with CVX huber = z^2 inside |z|<=1: g(z) = z^2/2
        f = 0.5 * oracle.huber(grid) + rho / 2 * (grid - v) ** 2
        assert abs(z - grid[np.argmin(f)]) < 1e-4


def test_slicemaker_cases():                                 # errorcheck.m:216-267
    assert oracle.slicemaker(0, 3, 10) == [4, 3, 3]
    assert oracle.slicemaker(0, 4, 8) == [2, 2, 2, 2]
    assert oracle.slicemaker(3, 2, 10) == [3, 3, 3, 1]
    assert oracle.slicemaker(5, 2, 10) == [5, 0]             # the reference's off-by-one, kept
    assert oracle.slicemaker([6, 4], 2, 10) == [6, 4]
    with pytest.raises(oracle.MatlabError):
        oracle.slicemaker([6, 5], 2, 10)


# ---- admm.m quirks ----------------------------------------------------------------------------
def test_setopt_hnormtol_reads_hreltol():                    # admm.m:927-928
    assert oracle.setopt({}, "Hnormtol", 1e-6) == 1e-6
    assert oracle.setopt({"Hnormtol": 1, "Hreltol": 5e-3}, "Hnormtol", 1e-6) == 5e-3
    with pytest.raises(oracle.MatlabError):
        oracle.setopt({"Hnormtol": 1e-3}, "Hnormtol", 1e-6)


def test_nonpositive_maxiters_becomes_1000():                # admm.m:334-339
    D, s, lam, _ = gen.lasso_problem(0, 40, 10)
    r = oracle.lasso(D, s, lam, {"maxiters": -5, "domaxiters": 1})
    assert r["steps"] == 1000


def test_divergence_return_leaves_results_unfinished():      # admm.m:692-700
    # a broken x-prox (in the spirit of examples/convergencechecking.m:198) trips the H-norm test
    n = 8
    rs = np.random.RandomState(0)
    q = rs.randn(n)
    minx_broken = lambda x, z, u, rho: 2 * x + (z - u) + q
    minz = lambda x, z, u, rho: oracle.zminSoftThresholding(x + u, 0.1)
    r = oracle.admm(minx_broken, minz, dict(A=1, At=1, B=-1, c=0, m=n, nA=n, nB=n, convtest=1))
    assert "steps" not in r and "xopt" not in r and len(r["Hnormsq"]) == 2


def test_missing_constraint_fields_raise_reference_messages():
    f = lambda x, z, u, rho: x
    with pytest.raises(oracle.MatlabError, match="Must specify a matrix A"):
        oracle.admm(f, f, dict(B=-1, c=0, m=3))
    with pytest.raises(oracle.MatlabError, match="Must specify a vector c"):
        oracle.admm(f, f, dict(A=1, B=-1))
    with pytest.raises(oracle.MatlabError, match="not a struct"):
        oracle.admm(f, f, None)


# ---- closed forms and optimality conditions ------------------------------------------------------
def test_lasso_identity_design_has_closed_form():
    # D = I  =>  argmin 1/2||x - s||^2 + lam||x||_1 = soft(s, lam)
    rs = np.random.RandomState(3)
    s = rs.randn(30)
    r = oracle.lasso(np.eye(30), s, 0.4, {"abstol": 1e-12, "reltol": 1e-12, "maxiters": 5000})
    assert np.allclose(r["zopt"], oracle.zminSoftThresholding(s, 0.4), atol=1e-9)


@pytest.mark.parametrize("rows,cols", [(256, 64), (60, 200)])
def test_lasso_kkt_and_tester_criterion(rows, cols):        # lassotest.m:143-147
    D, s, lam, testx = gen.lasso_problem(0, rows, cols)
    r = oracle.lasso(D, s, lam, {"objevals": 1, "abstol": 1e-11, "reltol": 1e-11, "maxiters": 20000})
    z = r["zopt"]
    g = D.T @ (D @ z - s)
    on = np.abs(z) > 1e-8
    assert np.allclose(g[on], -lam * np.sign(z[on]), atol=1e-6)      # stationarity on the support
    assert np.all(np.abs(g[~on]) <= lam + 1e-6)                       # subgradient bound off it
    obj = lambda x: 0.5 * np.sum((D @ x - s) ** 2) + lam * np.sum(np.abs(x))
    assert obj(r["xopt"]) < obj(testx)


@pytest.mark.parametrize("seed,rows,cols", [(2, 200, 50), (3, 60, 150)])
def test_lasso_optimum_matches_coordinate_descent(seed, rows, cols):   # lasso.m:227: 1/2||Dx - s||^2 + lambda||x||_1
    """An independent ALGORITHM on the same objective: scikit-learn's coordinate descent minimises
    1/(2m)||s - Dw||^2 + alpha||w||_1, i.e. alpha = lambda/m.  Tall (Cholesky of D'D + rho I) and fat (Woodbury) branch."""
    lm = pytest.importorskip("sklearn.linear_model")
    D, s, lam, _ = gen.lasso_problem(seed, rows, cols)
    r = oracle.lasso(D, s, lam, {"abstol": 1e-12, "reltol": 1e-12, "maxiters": 200000})
    sk = lm.Lasso(alpha=lam / rows, fit_intercept=False, tol=1e-14, max_iter=5000000).fit(D, s)
    assert np.abs(sk.coef_ - r["zopt"]).max() < 1e-8
    assert np.array_equal(sk.coef_ != 0, np.abs(r["zopt"]) > 1e-10)     # same support
    obj = lambda x: 0.5 * np.sum((D @ x - s) ** 2) + lam * np.sum(np.abs(x))
    assert abs(obj(r["zopt"]) - obj(sk.coef_)) <= 1e-10 * obj(sk.coef_)


def test_huberfit_optimum_matches_quasi_newton():           # huberfit.m:180: 1/2*sum(huber(Dx - s))
    D, s, _ = gen.huber_problem(1, 300, 8)
    r = oracle.huberfit(D, s, {"abstol": 1e-11, "reltol": 1e-11, "maxiters": 20000})
    f = lambda x: 0.5 * np.sum(oracle.huber(D @ x - s))
    g = lambda x: D.T @ np.clip(D @ x - s, -1.0, 1.0)
    o = sopt.minimize(f, np.zeros(D.shape[1]), jac=g, method="BFGS", options={"gtol": 1e-11})
    assert np.abs(o.x - r["xopt"]).max() < 1e-6 and abs(f(o.x) - f(r["xopt"])) <= 1e-10 * f(o.x)


def test_bounded_qp_optimum_matches_lbfgsb():                # quadraticprogram.m:210-216, getProxOps.m:1441-1474
    rs = np.random.RandomState(4)
    n = 25
    Mx = rs.randn(n, n)
    P, q = Mx @ Mx.T + 0.5 * np.eye(n), rs.randn(n)
    lb, ub = -0.3 * np.ones(n), 0.2 * np.ones(n)
    r = oracle.quadraticprogram(P, q, 0.0, lb, ub, {"abstol": 1e-12, "reltol": 1e-12, "maxiters": 50000})
    f = lambda x: 0.5 * x @ P @ x + q @ x
    o = sopt.minimize(f, np.zeros(n), jac=lambda x: P @ x + q, method="L-BFGS-B", bounds=list(zip(lb, ub)),
                      options={"ftol": 1e-15, "gtol": 1e-12, "maxiter": 10000})
    assert np.abs(o.x - r["zopt"]).max() < 1e-6 and f(r["zopt"]) <= f(o.x) + 1e-10
    assert np.all(r["zopt"] >= lb - 1e-12) and np.all(r["zopt"] <= ub + 1e-12)
    assert np.any(r["zopt"] == lb) and np.any(r["zopt"] == ub)          # both bounds active somewhere: a real projection


def test_lasso_fat_and_tall_branches_agree_on_square_problem():  # getProxOps.m:1199-1205 (Woodbury)
    D, s, lam, _ = gen.lasso_problem(1, 50, 50)
    tall = oracle.lasso(D, s, lam, {"domaxiters": 1, "maxiters": 25})
    # force the fat branch by appending a zero column (m < n) -- the extra coordinate stays 0
    D2 = np.hstack([D, np.zeros((50, 1))])
    fat = oracle.lasso(D2, s, lam, {"domaxiters": 1, "maxiters": 25})
    assert np.allclose(fat["xopt"][:50], tall["xopt"], rtol=1e-9, atol=1e-12)
    assert abs(fat["xopt"][50]) < 1e-14


def test_lad_matches_linear_program():                       # ladtest.m:149-168
    D, s, xtrue = gen.lad_problem(0, 120, 6)
    r = oracle.lad(D, s, {"abstol": 1e-10, "reltol": 1e-10, "maxiters": 20000})
    m, n = D.shape
    # min sum t  s.t. -t <= Dx - s <= t
    c = np.concatenate([np.zeros(n), np.ones(m)])
    A = np.block([[D, -np.eye(m)], [-D, -np.eye(m)]])
    lp = sopt.linprog(c, A_ub=A, b_ub=np.concatenate([s, -s]), bounds=[(None, None)] * n + [(0, None)] * m)
    assert lp.status == 0
    assert abs(np.sum(np.abs(D @ r["xopt"] - s)) - lp.fun) <= 1e-5 * lp.fun
    assert np.linalg.norm(r["xopt"] - xtrue) < 1e-3


def test_huberfit_stationarity_and_tester_criterion():      # huberfittest.m:154-158
    D, s, testx = gen.huber_problem(0, 400, 12)
    r = oracle.huberfit(D, s, {"abstol": 1e-11, "reltol": 1e-11, "maxiters": 20000})
    res = D @ r["xopt"] - s
    psi = np.clip(res, -1.0, 1.0)            # derivative of 1/2*huber
    assert np.linalg.norm(D.T @ psi) < 1e-6
    f = lambda x: 0.5 * np.sum(oracle.huber(D @ x - s))
    assert f(r["xopt"]) <= f(testx)


def test_basispursuit_feasible_and_not_worse_than_truth():   # basispursuittest.m:136-143
    D, s, testx = gen.bp_problem(0, 30, 90, density=0.05)
    r = oracle.basispursuit(D, s, {"maxiters": 10000})
    assert np.linalg.norm(D @ r["xopt"] - s) <= 1e-8 * np.linalg.norm(s)
    assert np.sum(np.abs(r["zopt"])) <= np.sum(np.abs(testx)) * (1 + 1e-3)


def test_totalvariation_beats_truth_and_tridiagonal_solve():  # totalvariationtest.m:151-155
    s, truth = gen.tv_problem(0, 128)
    lam = 5.0
    r = oracle.totalvariation(s, lam, {"maxiters": 10000})
    obj = lambda x: 0.5 * np.sum((x - s) ** 2) + lam * np.sum(np.abs(np.diff(x)))
    assert obj(r["xopt"]) < obj(truth)
    # x-update solves (I + rho D'D) x = s + rho D'(z-u) with D'D = tridiag(-1, 2, -1), (1,1) entry 1
    n, rho = 6, 0.7
    Dm = np.eye(n) - np.eye(n, k=1)
    DtD = Dm.T @ Dm
    assert DtD[0, 0] == 1 and DtD[1, 1] == 2 and DtD[n - 1, n - 1] == 2 and DtD[0, 1] == -1


def test_linearsvm_serial_equals_transpose_reduction():      # unwrappedadmm.m:76-78 vs :96-141
    D, ell = gen.svm_problem(0, 64, 64)
    np.random.seed(1)
    a = oracle.linearsvm(D, ell, 0.5, {"objevals": 1})
    np.random.seed(1)
    b = oracle.linearsvm(D, ell, 0.5, {"objevals": 1, "parallel": "both", "workers": 3})
    assert a["steps"] == b["steps"]
    assert np.allclose(a["xopt"], b["xopt"], rtol=1e-9, atol=1e-12)
    x = a["xopt"]
    assert abs(1 + x[1] / x[0]) <= 0.05                      # linearsvmtest.m:180-192 (errtol 0.05)


def test_linearsvm_0_1_string_is_hinge():                    # getProxOps.m:1094 strcmp(loss,'01')
    D, ell = gen.svm_problem(0, 32, 32)
    np.random.seed(2)
    a = oracle.linearsvm(D, ell, 0.5, {"lossfunction": "0-1"})
    np.random.seed(2)
    b = oracle.linearsvm(D, ell, 0.5, {"lossfunction": "hinge"})
    assert a["steps"] == b["steps"] and np.array_equal(a["xopt"], b["xopt"])


def test_unwrapped_forced_options():                         # unwrappedadmm.m:81-92
    D, ell = gen.svm_problem(0, 16, 16)
    np.random.seed(0)
    r = oracle.linearsvm(D, ell, 0.5, {"maxiters": 5, "stopcond": "standard"})
    o = r["options"]
    assert o["maxiters"] == 1000 and o["stopcond"] == "both" and o["nodualerror"] == 1
    assert np.all(np.isnan(r["dnorm"])) and np.all(np.isnan(r["derr"]))


def test_model_converges_to_the_least_squares_solution_and_passes_modeltest():
    # the minimiser of 1/2||Px-r||^2 + 1/2||Qx-s||^2 is (P'P + Q'Q) \\ (P'r + Q's) (modeltest.m:117);
    # pass criterion modeltest.m:133-141 with errtol = 1e-3 and the tester's options (:124-129)
    from admm_project_b200.generators import model_problem
    P, Q, r, s, truex = model_problem(0, 128, 128)
    res = oracle.model(P, Q, r, s, {"objevals": 1, "maxiters": 10000, "convtest": 1, "stopcond": "both"})
    obj = lambda x: 0.5 * np.sum((P @ x - r) ** 2) + 0.5 * np.sum((Q @ x - s) ** 2)
    assert abs(1 - obj(res["xopt"]) / obj(truex)) <= 1e-3 and np.linalg.norm(truex - res["xopt"]) <= 1e-3
    assert abs(res["objopt"] - (0.5 * np.sum((P @ res["xopt"] - r) ** 2) + 0.5 * np.sum((Q @ res["zopt"] - s) ** 2))) < 1e-9
    # a fixed point of both prox operators is the optimum: one iteration started there stays there
    rho = 2.0
    u = -(P.T @ (P @ truex - r)) / rho                     # stationarity of the x-update at (truex, truex, u)
    minx, minz, _ = oracle.getproxops("Model", dict(PtP=P.T @ P, Ptr=P.T @ r, QtQ=Q.T @ Q, Qts=Q.T @ s, n=128))
    assert np.allclose(minx(None, truex, u, rho), truex, atol=1e-9)
    assert np.allclose(minz(truex, None, u, rho), truex, atol=1e-9)
    # the accelerated variant reaches the optimum to rounding (examples/fasteradmmcomparison.m)
    acc = oracle.model(P, Q, r, s, {"fast": 1, "maxiters": 2000})
    assert np.linalg.norm(acc["xopt"] - truex) < 1e-10


def test_basispursuit_factored_form_equals_explicit_projector():
    """The oracle's test aid for n = 32768 (basispursuit_factored: v - D'((DD')\\(Dv - s))) against the reference's
    explicit P, q (basispursuit.m:116-120, getProxOps.m:1031): same steps, iterates within 1e-12."""
    from admm_project_b200 import generators as gen
    for rows, cols in ((64, 128), (100, 501)):
        D, s, _ = gen.bp_problem(0, rows, cols, density=0.05)
        opts = {"objevals": 1, "maxiters": 10000, "convtest": 0}
        a, b = oracle.basispursuit(D, s, opts), oracle.basispursuit_factored(D, s, opts)
        assert a["steps"] == b["steps"]
        for k in ("xopt", "zopt", "uopt", "pnorm", "dnorm", "objevals"):
            assert np.linalg.norm(a[k] - b[k]) <= 1e-12 * np.linalg.norm(a[k]), k


# ---- independent solvers (not ADMM): the PROBLEM each solver states is the one a third-party algorithm solves ------------
def test_linearsvm_iteration_minimises_the_hinge_term_only():   # linearsvm.m:185,233; unwrappedadmm.m:76-78
    """A reference quirk worth pinning: the x-update of the unwrapped formulation is x = pinv(D)(z - u) -- the ridge term
    1/2||x||^2 of the objective the solver REPORTS (linearsvm.m:233) is not part of the iteration.  Its fixed point
    minimises C*sum(hinge) alone, so against liblinear (which solves the stated problem) the hinge term is lower and the
    stated objective higher.  The drop-in reproduces the iteration, not the textbook SVM."""
    svm = pytest.importorskip("sklearn.svm")
    D, ell = oracle_gen().svm_problem(3, 150, 150, 0.3)
    C = 0.5
    np.random.seed(0)
    r = oracle.linearsvm(D, ell, C, {"objevals": 1})
    clf = svm.LinearSVC(loss="hinge", C=C, fit_intercept=False, dual=True, tol=1e-12, max_iter=2000000).fit(D, ell)
    w = clf.coef_.ravel()
    hinge = lambda x: float(np.sum(np.maximum(1 - ell * (D @ x), 0)))
    obj = lambda x: 0.5 * float(x @ x) + C * hinge(x)
    assert hinge(r["xopt"]) < hinge(w)                      # ADMM's answer fits the hinge term better ...
    assert obj(r["xopt"]) > obj(w) * 1.01                   # ... and is NOT the minimiser of the stated objective
    assert abs(r["objopt"] - obj(r["xopt"])) <= 1e-9 * obj(r["xopt"])    # the reported objective is the stated one


def test_basispursuit_optimum_matches_linear_program():     # basispursuit.m:140: min ||x||_1 s.t. Dx = s
    opt = pytest.importorskip("scipy.optimize")
    D, s, _ = oracle_gen().bp_problem(2, 30, 80, density=0.08)
    r = oracle.basispursuit(D, s, {"maxiters": 20000, "abstol": 1e-10, "reltol": 1e-9, "convtest": 0})
    n = D.shape[1]                                            # x = p - q, p, q >= 0: min 1'(p + q) s.t. D(p - q) = s
    lp = opt.linprog(np.ones(2 * n), A_eq=np.hstack([D, -D]), b_eq=s, bounds=(0, None), method="highs")
    assert lp.status == 0
    assert abs(np.sum(np.abs(r["xopt"])) - lp.fun) <= 1e-5 * lp.fun
    assert np.linalg.norm(D @ r["xopt"] - s) <= 1e-8 * np.linalg.norm(s)


def _tv1d_direct(y, lam):
    """Condat's direct O(n) algorithm for min 1/2||x - y||^2 + lam * sum|x_i - x_{i+1}| (L. Condat, 'A direct algorithm for
    1-D total variation denoising', IEEE SPL 2013) -- no iteration, no tolerance: an exact reference."""
    n = len(y)
    x = np.empty(n)
    k = k0 = km = kp = 0
    vmin, vmax = y[0] - lam, y[0] + lam
    umin, umax = lam, -lam
    while True:
        if k == n - 1:
            if umin < 0.0:
                x[k0:km + 1] = vmin
                k = k0 = km = km + 1
                vmin = y[k]; umin = lam; umax = y[k] + lam - vmax
            elif umax > 0.0:
                x[k0:kp + 1] = vmax
                k = k0 = kp = kp + 1
                vmax = y[k]; umax = -lam; umin = y[k] - lam - vmin
            else:
                vmin += umin / (k - k0 + 1)
                x[k0:k + 1] = vmin
                return x
            continue
        if y[k + 1] + umin < vmin - lam:
            x[k0:km + 1] = vmin
            k = k0 = km = kp = km + 1
            vmin = y[k]; vmax = y[k] + 2 * lam; umin = lam; umax = -lam
        elif y[k + 1] + umax > vmax + lam:
            x[k0:kp + 1] = vmax
            k = k0 = km = kp = kp + 1
            vmax = y[k]; vmin = y[k] - 2 * lam; umin = lam; umax = -lam
        else:
            k += 1
            umin += y[k] - vmin
            umax += y[k] - vmax
            if umin >= lam:
                vmin += (umin - lam) / (k - k0 + 1)
                umin = lam
                km = k
            if umax <= -lam:
                vmax += (umax + lam) / (k - k0 + 1)
                umax = -lam
                kp = k


def test_totalvariation_fixed_point_and_the_last_row_of_D():    # totalvariation.m:122-134, getProxOps.m:172-199
    """The reference's difference operator is square: (Dx)_i = x_i - x_{i+1} and a LAST ROW (Dx)_n = x_n.  The iteration
    therefore minimises 1/2||x - s||^2 + lam*(sum|x_i - x_{i+1}| + |x_n|).  (a) an exact dual certificate of THAT problem
    at the oracle's answer: with g = cumsum((s - x)/lam) (so that x - s + lam*D'g = 0), g must be a subgradient of |Dx|;
    (b) against Condat's direct algorithm, which solves the plain problem, each answer wins on its own objective."""
    s, _ = oracle_gen().tv_problem(5, 400)
    s = np.asarray(s, dtype=float)
    lam = 1.5
    r = oracle.totalvariation(s, lam, {"maxiters": 20000, "abstol": 1e-10, "reltol": 1e-9})
    x = r["xopt"]
    Dx = np.append(x[:-1] - x[1:], x[-1])
    g = np.cumsum((s - x) / lam)                             # (D'g)_i = g_i - g_{i-1}
    assert np.max(np.abs(g)) <= 1 + 1e-6
    active = np.abs(Dx) > 1e-6
    assert np.allclose(g[active], np.sign(Dx[active]), atol=1e-5)
    xd = _tv1d_direct(s, lam)
    plain = lambda v: 0.5 * np.sum((v - s) ** 2) + lam * np.sum(np.abs(np.diff(v)))
    with_last = lambda v: plain(v) + lam * abs(v[-1])
    assert with_last(x) < with_last(xd) and plain(xd) < plain(x)
    # and the direct algorithm is itself certified on the plain problem: same construction without the last row
    gd = np.cumsum((s - xd) / lam)[:-1]
    assert np.max(np.abs(gd)) <= 1 + 1e-9 and abs(np.sum(s - xd)) <= 1e-9 * np.sum(np.abs(s))


def oracle_gen():
    from admm_project_b200 import generators
    return generators
