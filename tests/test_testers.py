"""The reference's tester scripts (testers/*.m) re-expressed in Python (admm_project_b200/testers.py): the
same recipes, forced options and pass criteria.  CPU: run against the oracle -- the restatement passes the
reference's own acceptance tests at the demo sizes.  GPU: run against the engine through the C-ABI."""
import numpy as np
import pytest

from admm_project_b200 import testers

DEMOS = ["lasso", "linearsvm", "huberfit", "totalvariation", "basispursuit", "model"]


def seeded(name, solvers, **kw):
    np.random.seed(7)                                  # unwrappedadmm.m:87-89 draws the SVM's initial iterates
    return testers.TESTERS[name](0, solvers=solvers, **kw)


@pytest.mark.parametrize("name", DEMOS)
def test_oracle_passes_the_reference_testers_at_demo_size(name):
    import oracle
    results, test = seeded(name, oracle)
    assert test["failed"] == 0, test["failreason"]
    assert results["steps"] >= 1


def test_ladtest_criterion_needs_the_tolerances_the_tester_cannot_set():
    # ladtest.m:149 asks ||xtrue - xopt|| < 1e-3 with xtrue = 10*randn: ADMM at the default reltol 1e-3 does not
    # get there (oracle and engine alike); with tight tolerances it does.  Recorded so the mirror is honest.
    import oracle
    _, loose = testers.ladtest(0, 256, 16, solvers=oracle)
    _, tight = testers.ladtest(0, 256, 16, options={"abstol": 1e-9, "reltol": 1e-9, "maxiters": 20000}, solvers=oracle)
    assert tight["failed"] == 0, tight["failreason"]
    assert loose["xresidual"] >= tight["xresidual"]


def test_solvertester_sweeps_sizes_with_the_oracle():
    import oracle
    out = testers.solvertester("lasso", 4, 6, trials=1, solvers=oracle)
    assert [t["scale"] for t in out["trials"]] == [4, 5, 6] and out["failures"] == 0
    with pytest.raises(ValueError):
        testers.solvertester("covarianceselection", solvers=oracle)


@pytest.mark.parametrize("solver", ["huberfit", "basispursuit", "totalvariation", "model", "linearsvm"])
def test_solvertester_every_in_scope_solver_with_the_oracle(solver):
    import oracle
    np.random.seed(3)
    lo = 7 if solver == "linearsvm" else 4             # 2^5 + 2^5 jittered points do not pin the 45-degree line to 5 %
    out = testers.solvertester(solver, lo, lo + 1, trials=1, seed=1, solvers=oracle)
    assert len(out["trials"]) == 2
    assert out["failures"] == 0, out


@pytest.mark.gpu
@pytest.mark.parametrize("name", DEMOS)
def test_engine_passes_the_reference_testers_and_agrees_with_the_oracle(engine, name):
    import oracle
    res, test = seeded(name, None, engine=engine)
    ref, rtest = seeded(name, oracle)
    assert test["failed"] == 0, test["failreason"]
    assert res["steps"] == ref["steps"]
    assert abs(test["objopt"] - rtest["objopt"]) <= 1e-9 * abs(rtest["objopt"])


@pytest.mark.gpu
def test_solvertester_on_the_engine(engine):
    for solver in ("huberfit", "basispursuit", "totalvariation"):
        out = testers.solvertester(solver, 5, 7, trials=1, engine=engine)
        assert out["failures"] == 0, out
