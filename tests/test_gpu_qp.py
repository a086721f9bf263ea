"""GPU parity: bounded quadratic program (cached Cholesky x-update + box projection) and the nonneg
projection, against the oracle."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import quadraticprogram
from admm_project_b200 import _lib as L

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def qp_problem(seed, n):
    rs = np.random.RandomState(seed)
    X = rs.randn(n + 5, n)
    P = X.T @ X / n
    q = rs.randn(n)
    lb = -0.3 * rs.rand(n)
    ub = 0.3 * rs.rand(n)
    return P, q, 0.7, lb, ub


@pytest.mark.parametrize("n,relax,rho", [(64, 1.0, 1.0), (300, 1.5, 2.0), (1025, 1.0, 0.5)])
def test_bounded_qp_matches_oracle(engine, n, relax, rho):
    P, q, r, lb, ub = qp_problem(n, n)
    opts = {"objevals": 1, "relax": relax, "rho": rho, "convtest": 1}
    ref = oracle.quadraticprogram(P, q, r, lb, ub, opts)
    res = quadraticprogram(P, q, r, ub, lb, opts, engine=engine)      # swapped bounds are re-ordered (:312-315)
    assert res["steps"] == ref["steps"]
    for k in ("xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr", "objevals", "xvals", "zvals", "uvals"):
        assert rel(res[k], ref[k]) < 1e-9, (k, rel(res[k], ref[k]))
    assert np.all(res["zopt"] >= lb - 1e-15) and np.all(res["zopt"] <= ub + 1e-15)


def test_nonneg_projection_kind(engine):
    # same x-update, z = pos(x + u): minimise 1/2 x'Px + q'x subject to x >= 0; check the KKT conditions
    P, q, r, _, _ = qp_problem(3, 80)
    engine.setup_quadratic(L.PROX_NONNEG, P, q, r, 1.0)
    o = engine.default_options()
    o.abstol, o.reltol, o.maxiters, o.history = 1e-11, 1e-11, 20000, 0
    res = engine.solve(o, want_history=False)
    z = res["zopt"]
    g = P @ z + q
    assert np.all(z >= 0)
    assert np.all(g[z > 1e-9] < 1e-6) and np.all(np.abs(g[z > 1e-9]) < 1e-6)
    assert np.all(g[z <= 1e-9] > -1e-6)


def test_qp_reference_errors(engine):
    from admm_project_b200 import EngineError, MatlabError
    P, q, r, lb, ub = qp_problem(1, 8)
    with pytest.raises(MatlabError, match="do not match"):
        quadraticprogram(P, q[:-1], r, lb, ub, {}, engine=engine)
    with pytest.raises(MatlabError, match="upper and lower bound"):
        mixed = lb.copy(); mixed[0] = 1.0
        quadraticprogram(P, q, r, mixed, ub * 0, {}, engine=engine)
    with pytest.raises(EngineError, match="standard"):
        quadraticprogram(P, q, r, np.eye(8), np.ones(8), {}, engine=engine)
