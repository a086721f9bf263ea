"""GPU parity: 1-D total variation (windowed tridiagonal scans + fused stencil prox) vs the oracle."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import totalvariation
from admm_project_b200 import generators as gen

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("n,rho,relax,lam", [(128, 1.0, 1.0, 5.0), (1000, 1.0, 1.5, 2.0), (50000, 1.0, 1.0, 3.0),
                                             (20001, 50.0, 1.0, 3.0), (30000, 7.0, 1.7, 1.0), (2, 2.0, 1.3, 0.5),
                                             (17, 1.0, 1.0, 0.5), (30011, 20.0, 1.0, 2.0), (12345, 20.0, 1.2, 2.0)])
def test_totalvariation_matches_oracle(engine, n, rho, relax, lam):
    s, truth = gen.tv_problem(0, n)
    opts = {"objevals": 1, "maxiters": 3000, "rho": rho, "relax": relax, "history": int(n <= 1000)}
    if relax != 1.0:
        # admm.m:521 hands Axhat to the z-prox in x's slot and getProxOps.m:199 applies D to it again;
        # with that quirk the relaxed TV iteration grows without bound (oracle and engine alike), so
        # parity is checked on a fixed number of early iterations.
        opts.update(maxiters=30, domaxiters=1)
    ref = oracle.totalvariation(s, lam, opts)
    res = totalvariation(s, lam, opts, engine=engine)
    assert res["steps"] == ref["steps"], (res["steps"], ref["steps"])
    for k in ("xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr", "objevals"):
        assert rel(res[k], ref[k]) < 1e-9, (k, rel(res[k], ref[k]))
    if n <= 1000:
        for k in ("xvals", "zvals", "uvals"):
            assert rel(res[k], ref[k]) < 1e-9, k
    if n >= 128:
        obj = lambda x: 0.5 * np.sum((x - s) ** 2) + lam * np.sum(np.abs(np.diff(x)))
        assert (obj(res["xopt"]) < obj(truth)) == (obj(ref["xopt"]) < obj(truth))    # totalvariationtest.m:151-155


@pytest.mark.parametrize("n,relax", [(9001, 1.0), (6000, 1.3)])
def test_totalvariation_history_through_the_fused_kernel(engine, n, relax):
    # n > 2 windows of the fused kernel: interior segments take its predicate-free body, the first and
    # last its general body; the x history is written by the kernel itself (x never reaches h->x)
    s, _ = gen.tv_problem(3, n)
    opts = {"objevals": 1, "maxiters": 25, "domaxiters": 1, "relax": relax, "history": 1}
    ref = oracle.totalvariation(s, 2.0, opts)
    res = totalvariation(s, 2.0, opts, engine=engine)
    assert res["steps"] == ref["steps"] == 25
    for k in ("xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr", "objevals", "xvals", "zvals", "uvals"):
        assert rel(res[k], ref[k]) < 1e-9, (k, rel(res[k], ref[k]))


def test_totalvariation_x_update_is_the_tridiagonal_solve(engine):
    # one iteration from random (z0, u0): x must solve (I + rho D'D) x = s + rho D'(z0 - u0)
    n, rho = 40000, 3.0
    rs = np.random.RandomState(1)
    s, z0, u0 = rs.randn(n), rs.randn(n), rs.randn(n)
    res = totalvariation(s, 1.0, {"rho": rho, "maxiters": 1, "z0": z0, "u0": u0, "x0": np.zeros(n)}, engine=engine)
    x = res["xopt"]
    Dt = lambda w: w - np.concatenate([[0.0], w[:-1]])
    D = lambda v: v - np.concatenate([v[1:], [0.0]])
    lhs = x + rho * Dt(D(x))
    rhs = s + rho * Dt(z0 - u0)
    assert rel(lhs, rhs) < 1e-13


@pytest.mark.parametrize("n,rho,relax", [(256, 1e7, 1.0), (50000, 1e4, 1.0), (20001, 1e5, 1.0), (8192, 3e3, 1.0), (8193, 1e5, 1.3),
                                         (100003, 2e5, 1.0)])
def test_totalvariation_any_rho_matches_oracle(engine, n, rho, relax):
    """The reference accepts any rho (getProxOps.m:1047 factorises I + rho*D'D whatever it is).  Past rho ~ 2500 the
    recurrences no longer forget their carry within a window, and the engine chains its 8192-element segments through
    their affine aggregates instead (tv_exact_kernel): same steps, iterates within 1e-9.  (rho is kept where
    cond(I + rho*D'D) <= 4*rho leaves two different factorisations 1e-9 of agreement; beyond, see the residual test.)"""
    s, _ = gen.tv_problem(1, n)
    opts = {"objevals": 1, "maxiters": 40, "domaxiters": 1, "rho": rho, "relax": relax, "history": 0}
    ref = oracle.totalvariation(s, 2.0, opts)
    res = totalvariation(s, 2.0, opts, engine=engine)
    assert res["steps"] == ref["steps"] == 40
    for k in ("xopt", "zopt", "uopt", "perr", "derr", "objevals"):
        assert rel(res[k], ref[k]) < 1e-9, (k, rel(res[k], ref[k]))
    # at large rho the residual norms are differences of nearly equal vectors (||Dx - z|| ~ 1e-8 ||x|| at rho = 1e7), so
    # they carry that many fewer digits than the iterates in ANY implementation: 1e-9 relative to the iterates' scale
    scale = np.linalg.norm(ref["xopt"]) + np.linalg.norm(ref["zopt"])
    for k in ("pnorm", "dnorm"):
        assert np.max(np.abs(res[k] - ref[k])) <= 1e-9 * max(scale, np.max(np.abs(ref[k]))), k
        assert rel(res[k], ref[k]) < 1e-6, (k, rel(res[k], ref[k]))


def test_totalvariation_exact_solve_equals_windowed_solve(engine, monkeypatch):
    # the chained solve forced at a small rho must reproduce the fused windowed kernel, and solve the tridiagonal system
    n, rho = 70001, 3.0
    rs = np.random.RandomState(2)
    s, z0, u0 = rs.randn(n), rs.randn(n), rs.randn(n)
    opts = {"rho": rho, "maxiters": 12, "domaxiters": 1, "z0": z0, "u0": u0, "x0": np.zeros(n), "history": 0}
    win = totalvariation(s, 1.0, opts, engine=engine)
    monkeypatch.setenv("ADMM_B200_TV_EXACT", "1")
    ex = totalvariation(s, 1.0, opts, engine=engine)
    for k in ("xopt", "zopt", "uopt", "pnorm", "dnorm"):
        assert rel(ex[k], win[k]) < 1e-12, (k, rel(ex[k], win[k]))
    one = totalvariation(s, 1.0, dict(opts, maxiters=1), engine=engine)
    x = one["xopt"]
    Dt = lambda w: w - np.concatenate([[0.0], w[:-1]])
    D = lambda v: v - np.concatenate([v[1:], [0.0]])
    assert rel(x + rho * Dt(D(x)), s + rho * Dt(z0 - u0)) < 1e-13


def test_totalvariation_huge_rho_solves_the_system(engine):
    # rho = 1e9 on 300001 elements (cond ~ 4e9: no two solvers agree to 1e-9 there): the x-update must still solve ITS system
    n, rho = 300001, 1e9
    rs = np.random.RandomState(4)
    s, z0, u0 = rs.randn(n), rs.randn(n), rs.randn(n)
    res = totalvariation(s, 1.0, {"rho": rho, "maxiters": 1, "z0": z0, "u0": u0, "x0": np.zeros(n), "history": 0}, engine=engine)
    x = res["xopt"]
    Dt = lambda w: w - np.concatenate([[0.0], w[:-1]])
    D = lambda v: v - np.concatenate([v[1:], [0.0]])
    rhs = s + rho * Dt(z0 - u0)
    assert rel(x + rho * Dt(D(x)), rhs) < 1e-10
