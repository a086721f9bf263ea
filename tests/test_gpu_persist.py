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


@pytest.mark.parametrize("rows,cols,classes", [(3000, 300, 4), (20000, 200, 10), (5001, 129, 3), (12000, 784, 10)])
def test_class_batch_persistent_loop_matches_oracle(engine, rows, cols, classes):
    """linearsvm_onevsall without the objective: the batch runs as the persistent kernel of csrc/persist_batch.cuh
    (one 16-row tile of Q per iteration for ALL classes); every class must reproduce its stand-alone oracle run."""
    from admm_project_b200 import linearsvm_onevsall
    D, ELL = gen.svm_mnist_like(11, rows, cols, nclass=classes)
    if cols < 784:
        D = D + 1e-3 * np.random.RandomState(12).randn(rows, cols)
    np.random.seed(13)
    launches = engine.launch_count()
    outs = linearsvm_onevsall(D, ELL, 0.5, {}, engine=engine)
    used = engine.launch_count() - launches
    np.random.seed(13)
    worst = 0
    for k in range(classes):
        ref = oracle.linearsvm(D, ELL[:, k], 0.5, {"history": 0})
        assert outs[k]["steps"] == ref["steps"], (k, outs[k]["steps"], ref["steps"])
        for key in ("xopt", "zopt", "uopt", "pnorm", "perr"):
            assert rel(outs[k][key], ref[key]) < TOL, (k, key, rel(outs[k][key], ref[key]))
        worst = max(worst, ref["steps"])
    if cols % 4 == 0:                      # (other widths take the stepwise batch kernels)
        assert used < 60 + worst // 4      # bursts of the persistent kernel, not ~9 launches per iteration


def test_class_batch_c3_size_three_of_ten_classes(engine):
    """BASELINE.json configs[2] at full size: 60000 x 784, ten one-vs-all classes in one batch; three of the
    classes are checked against the oracle (each oracle run costs a pinv of D)."""
    from admm_project_b200 import linearsvm_onevsall
    D, ELL = gen.svm_mnist_like(0, 60000, 784)
    np.random.seed(21)
    outs = linearsvm_onevsall(D, ELL, 0.5, {}, engine=engine)
    for k in (0, 5, 9):
        # replay class k's own init: the batch drew rand(n), rand(m), rand(m) per class, in class order
        np.random.seed(21)
        for _ in range(k):
            np.random.rand(784), np.random.rand(60000), np.random.rand(60000)
        ref = oracle.linearsvm(D, ELL[:, k], 0.5, {"history": 0})
        assert outs[k]["steps"] == ref["steps"], (k, outs[k]["steps"], ref["steps"])
        for key in ("xopt", "zopt", "uopt", "pnorm", "perr"):
            assert rel(outs[k][key], ref[key]) < TOL, (k, key, rel(outs[k][key], ref[key]))
