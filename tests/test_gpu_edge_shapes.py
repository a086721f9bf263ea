"""Edge shapes through the C-ABI: one-row / one-column / tiny / empty problems for every in-scope solver.  The
reference's testers only draw comfortable shapes (testers/lassotest.m:109-122 and friends); MATLAB itself has no
trouble with a 1 x 1 `chol` or an empty product, so the engine must not either.  For each shape the oracle is run
first: where it raises (the reference's own checks, e.g. basispursuit.m's square-matrix error), the engine's host
mirror must raise as well; otherwise the usual parity holds -- same steps, iterates within 1e-9."""
import warnings

import numpy as np
import pytest

import oracle
from admm_project_b200 import MatlabError, basispursuit, huberfit, lad, lasso, lasso_path, linearsvm, totalvariation
from admm_project_b200._lib import EngineError

pytestmark = pytest.mark.gpu
TOL = 1e-9


def close(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.linalg.norm(a - b)) <= TOL * float(np.linalg.norm(b)) + 1e-13


def both(ref_fn, eng_fn):
    """Run the oracle, then the engine: both raise, or both agree."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ref = ref_fn()
        except (oracle.MatlabError, np.linalg.LinAlgError, ValueError, TypeError, IndexError, ZeroDivisionError) as ex:
            with pytest.raises((MatlabError, EngineError, ValueError, TypeError, IndexError)):
                eng_fn()
            return None, ex
        res = eng_fn()
    if "steps" not in ref:          # the H-norm early return (admm.m:686-701): neither side reports an optimum
        assert "steps" not in res
        return res, ref
    assert res["steps"] == ref["steps"], (res["steps"], ref["steps"])
    for k in ("xopt", "zopt", "uopt"):
        assert close(res[k], ref[k]), (k, res[k], ref[k])
    assert close(res["pnorm"], ref["pnorm"]) and close(res["perr"], ref["perr"])
    return res, ref


@pytest.mark.parametrize("rows,cols", [(1, 1), (2, 1), (1, 3), (5, 3), (3, 5), (3, 3), (17, 2), (2, 17)])
def test_lasso_tiny(engine, rows, cols):
    rs = np.random.RandomState(100 * rows + cols)
    D, s = rs.randn(rows, cols), rs.randn(rows)
    both(lambda: oracle.lasso(D, s, 0.1, {"relax": 1.2}), lambda: lasso(D, s, 0.1, {"relax": 1.2}, engine=engine))


def test_empty_inputs_are_refused_loudly(engine):
    """MATLAB runs lasso on a 0 x 3 matrix (one iteration, x = 0); the engine refuses empty problems with a clear
    error instead of launching empty grids (deviations)."""
    for fn in (lambda: lasso(np.zeros((0, 3)), np.zeros(0), 0.1, {}, engine=engine),
               lambda: huberfit(np.zeros((0, 3)), np.zeros(0), {}, engine=engine),
               lambda: totalvariation(np.zeros(0), 0.5, {}, engine=engine)):
        with pytest.raises((EngineError, MatlabError), match="bad dimensions|not a vector|empty"):
            fn()
    res = lasso(np.eye(3), np.ones(3), 0.1, {}, engine=engine)       # the handle is still usable afterwards
    assert res["steps"] > 0


def test_lasso_path_of_one_and_on_a_tiny_factor(engine):
    rs = np.random.RandomState(5)
    D, s = rs.randn(9, 3), rs.randn(9)
    lams = np.array([0.3])
    rb = lasso_path(D, s, lams, {}, engine=engine)
    ref = oracle.lasso(D, s, 0.3, {"history": 0})
    assert int(rb["steps"][0]) == ref["steps"] and close(rb["xopt"][:, 0], ref["xopt"])
    lams = np.array([1.0, 0.1, 0.01, 0.0])          # lambda = 0 is a legal value (getProxOps.m:455: nonnegative)
    rb = lasso_path(D, s, lams, {}, engine=engine)
    for j, lam in enumerate(lams):
        ref = oracle.lasso(D, s, lam, {"history": 0})
        assert int(rb["steps"][j]) == ref["steps"] and close(rb["xopt"][:, j], ref["xopt"]), j


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 33])
def test_totalvariation_tiny(engine, n):
    s = np.random.RandomState(n).randn(n)
    both(lambda: oracle.totalvariation(s, 0.5, {}), lambda: totalvariation(s, 0.5, {}, engine=engine))


@pytest.mark.parametrize("fit", ["huber", "lad"])
@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 1), (2, 2), (4, 3), (33, 1)])
def test_robust_fit_tiny(engine, fit, rows, cols):
    rs = np.random.RandomState(10 * rows + cols)
    D, s = rs.randn(rows, cols), rs.randn(rows)
    ofn, efn = (oracle.huberfit, huberfit) if fit == "huber" else (oracle.lad, lad)
    both(lambda: ofn(D, s, {}), lambda: efn(D, s, {}, engine=engine))


@pytest.mark.parametrize("rows,cols", [(1, 2), (2, 5), (1, 1), (3, 3), (1, 40)])
def test_basispursuit_tiny(engine, rows, cols):
    rs = np.random.RandomState(7 * rows + cols)
    D, s = rs.randn(rows, cols), rs.randn(rows)
    both(lambda: oracle.basispursuit(D, s, {}), lambda: basispursuit(D, s, {}, engine=engine))


@pytest.mark.parametrize("rows,cols", [(2, 1), (3, 2), (2, 2), (9, 1), (40, 3)])
def test_linearsvm_tiny(engine, rows, cols):
    rs = np.random.RandomState(3 * rows + cols)
    D = rs.randn(rows, cols)
    ell = np.where(np.arange(rows) % 2 == 0, 1.0, -1.0)

    def ref_fn():
        np.random.seed(11)          # unwrappedadmm.m draws x0 / z0 / u0 with rand
        return oracle.linearsvm(D, ell, 0.5, {})

    def eng_fn():
        np.random.seed(11)
        return linearsvm(D, ell, 0.5, {}, engine=engine)
    both(ref_fn, eng_fn)
