"""GPU parity of the model problem (solvers/model.m, getProxOps.m:55-110, 952-1012) against the oracle:
two cached-Cholesky prox operators, vanilla / relaxed / fast / accelerated ADMM, the rho-change path
(getProxOps.m:967-970) and the pass criterion of testers/modeltest.m:133-160."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import MatlabError, admm, getproxops, model
from admm_project_b200 import generators as gen

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def compare(res, ref, keys=("xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr", "objevals")):
    assert res["steps"] == ref["steps"], (res["steps"], ref["steps"])
    for k in keys:
        if k in ref:
            assert rel(res[k], ref[k]) < TOL, (k, rel(res[k], ref[k]))


@pytest.mark.parametrize("rows,cols", [(128, 128), (300, 60), (1000, 257)])
def test_model_matches_oracle_with_the_testers_options(engine, rows, cols):
    P, Q, r, s, truex = gen.model_problem(0, rows, cols)
    opts = {"objevals": 1, "maxiters": 10000, "convtest": 1, "stopcond": "both"}       # modeltest.m:124-129
    ref = oracle.model(P, Q, r, s, opts)
    res = model(P, Q, r, s, opts, engine=engine)
    compare(res, ref)
    assert rel(res["Hnormsq"], ref["Hnormsq"]) < TOL, rel(res["Hnormsq"], ref["Hnormsq"])
    # elementwise too; the late entries are squared differences of nearly equal iterates (||dz|| ~ 1e-6 ||z||), so they
    # carry ~6 fewer digits than the iterates in ANY implementation
    assert np.allclose(res["Hnormsq"], ref["Hnormsq"], rtol=1e-6, atol=1e-22)
    for k in ("xvals", "zvals", "uvals"):
        assert rel(res[k], ref[k]) < TOL, k
    # modeltest.m:133-141: objective and x within errtol = 1e-3 of the true least-squares solution
    obj = lambda x: 0.5 * np.sum((P @ x - r) ** 2) + 0.5 * np.sum((Q @ x - s) ** 2)
    assert abs(1 - obj(res["xopt"]) / obj(truex)) <= 1e-3
    assert np.linalg.norm(truex - res["xopt"]) <= 1e-3


@pytest.mark.parametrize("relax,rho", [(1.6, 1.0), (1.0, 5.0), (1.8, 0.3)])
def test_model_relaxed_and_other_rho(engine, relax, rho):
    P, Q, r, s, _ = gen.model_problem(1, 400, 90)
    opts = {"objevals": 1, "relax": relax, "rho": rho, "history": 0}
    compare(model(P, Q, r, s, opts, engine=engine), oracle.model(P, Q, r, s, opts))


@pytest.mark.parametrize("fasttype", ["weak", "strong"])
def test_model_fast_variants(engine, fasttype):
    # examples/fasteradmmcomparison.m:70-87 runs the model problem with vanilla, fast and accelerated ADMM
    P, Q, r, s, _ = gen.model_problem(2, 256, 64)
    opts = {"fast": 1, "fasttype": fasttype, "objevals": 1, "history": 0, "maxiters": 50, "domaxiters": 1}
    ref = oracle.model(P, Q, r, s, opts)
    res = model(P, Q, r, s, opts, engine=engine)
    assert res["steps"] == ref["steps"]
    for k in ("xopt", "zopt", "uopt", "avals", "objevals"):
        assert rel(res[k], ref[k]) < TOL, (k, rel(res[k], ref[k]))
    if fasttype == "weak":
        assert np.array_equal(res["restarted"], ref["restarted"])
        assert rel(res["dvals"], ref["dvals"]) < TOL, rel(res["dvals"], ref["dvals"])
    else:
        assert rel(res["pnorm"], ref["pnorm"]) < TOL


def test_model_rho_change_refactors_from_the_cached_gram(engine):
    # getProxOps.m:967-970 re-adds rho to diag(PtP) when it changes; the engine keeps PtP / QtQ and only
    # refactors -- same handles, new options.rho (examples/stepsizetesting.m:56-70 sweeps rho)
    P, Q, r, s, _ = gen.model_problem(3, 500, 120)
    minx, minz, _ = getproxops("Model", {"engine": engine, "P": P, "Q": Q, "r": r, "s": s, "n": 120})
    base = dict(A=1, B=-1, c=0, m=120, nA=120, nB=120, obj="engine", objevals=1, history=0)
    for rho in (1.0, 4.0, 0.25, 4.0):
        res = admm(minx, minz, dict(base, rho=rho))
        compare(res, oracle.model(P, Q, r, s, {"objevals": 1, "rho": rho, "history": 0}))


def test_model_reference_error_messages(engine):
    P, Q, r, s, _ = gen.model_problem(0, 20, 5)
    with pytest.raises(MatlabError, match="rows in P do not match number of rows in Q"):
        model(P, Q[:10], r, s, {}, engine=engine)
    with pytest.raises(MatlabError, match="columns in P do not match number of columns in Q"):
        model(P, Q[:, :3], r, s, {}, engine=engine)
    with pytest.raises(MatlabError, match="rows in P does not match length of vector r"):
        model(P, Q, r[:7], s, {}, engine=engine)
    with pytest.raises(MatlabError, match="rows in Q does not match length of vector s"):
        model(P, Q, r, s[:7], {}, engine=engine)
    with pytest.raises(MatlabError, match="not a struct"):
        model(P, Q, r, s, None, engine=engine)
