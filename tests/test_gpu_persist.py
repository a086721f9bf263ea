"""GPU parity of the persistent A = D iteration (csrc/persist.cuh): with nodualerror, no objective and no history
the whole loop of linearsvm / huberfit / lad runs as one cooperative kernel per burst on Q = D*inv(R)' -- no
triangular solve and no x inside the loop (x = inv(R)' t after it).  The iterates must still be the reference's:
same `steps`, x / z / u and the residual histories within 1e-9 of the oracle (which runs pinv(D) for the SVM,
unwrappedadmm.m:76-78, and R'\\(R\\.) for Huber / LAD, getProxOps.m:1514)."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import huberfit, lad, linearsvm
from admm_project_b200 import generators as gen

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def compare(res, ref):
    assert res["steps"] == ref["steps"], (res["steps"], ref["steps"])
    for k in ("xopt", "zopt", "uopt", "pnorm", "perr"):
        assert rel(res[k], ref[k]) < TOL, (k, rel(res[k], ref[k]))
    assert np.all(np.isnan(res["dnorm"])) and np.all(np.isnan(res["derr"]))


@pytest.mark.parametrize("rows,cols", [(3000, 700), (20000, 200), (9001, 333), (60000, 784)])
def test_svm_persistent_loop_matches_oracle(engine, rows, cols):
    D, ELL = gen.svm_mnist_like(1, rows, cols)
    D = D + 1e-3 * np.random.RandomState(2).randn(rows, cols) if rows < 20000 else D
    opts = {"history": 0}                                    # unwrappedadmm.m:90-92 forces nodualerror, stopcond both
    np.random.seed(4)
    ref = oracle.linearsvm(D, ELL[:, 2], 0.5, opts)
    np.random.seed(4)
    launches = engine.launch_count()
    res = linearsvm(D, ELL[:, 2], 0.5, opts, engine=engine)
    compare(res, ref)
    # the loop really ran as bursts of the persistent kernel: a handful of launches, not ~5 per iteration
    assert engine.launch_count() - launches < 40 + ref["steps"] // 4


def test_svm_persistent_matches_the_stepwise_path(engine, monkeypatch):
    """Same problem through the persistent kernel and through the per-iteration kernels (objevals = 1 forces them):
    the two formulations (Q t versus D (R'\\(R\\d))) agree to rounding."""
    D, ELL = gen.svm_mnist_like(3, 12000, 256)
    np.random.seed(6)
    a = linearsvm(D, ELL[:, 0], 0.5, {"history": 0}, engine=engine)
    np.random.seed(6)
    b = linearsvm(D, ELL[:, 0], 0.5, {"history": 0, "objevals": 1}, engine=engine)
    assert a["steps"] == b["steps"]
    for k in ("xopt", "zopt", "uopt", "pnorm", "perr"):
        assert rel(a[k], b[k]) < 1e-11, k


@pytest.mark.parametrize("problem", ["huber", "lad"])
@pytest.mark.parametrize("relax", [1.0, 1.6])
def test_robust_fit_persistent_loop_matches_oracle(engine, problem, relax):
    rows, cols = 30000, 160
    D, s, _ = (gen.huber_problem if problem == "huber" else gen.lad_problem)(5, rows, cols)
    opts = {"history": 0, "nodualerror": 1, "convtest": 1, "relax": relax, "stopcond": "both", "maxiters": 400}
    ref = (oracle.huberfit if problem == "huber" else oracle.lad)(D, s, opts)
    res = (huberfit if problem == "huber" else lad)(D, s, opts, engine=engine)
    compare(res, ref)
    assert rel(res["Hnormsq"], ref["Hnormsq"]) < TOL


def test_persistent_loop_domaxiters_and_small_bursts(engine):
    D, ELL = gen.svm_mnist_like(7, 8000, 300)
    opts = {"history": 0, "check_every": 3}
    np.random.seed(8)
    ref = oracle.linearsvm(D, ELL[:, 1], 0.5, opts)
    np.random.seed(8)
    res = linearsvm(D, ELL[:, 1], 0.5, opts, engine=engine)
    compare(res, ref)
