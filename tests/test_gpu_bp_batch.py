"""GPU parity: basis pursuit (factored projector) and the lasso regularisation path (batch of lambdas
as multi-RHS triangular GEMMs) against the oracle."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import basispursuit, lasso
from admm_project_b200 import generators as gen

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("rows,cols", [(64, 128), (100, 501), (256, 1280)])
def test_basispursuit_matches_oracle(engine, rows, cols):
    D, s, testx = gen.bp_problem(0, rows, cols, density=0.05)
    opts = {"objevals": 1, "maxiters": 10000, "convtest": 0}          # basispursuittest.m:121-123
    ref = oracle.basispursuit(D, s, opts)
    res = basispursuit(D, s, opts, engine=engine)
    assert res["steps"] == ref["steps"]
    for k in ("xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr", "objevals"):
        assert rel(res[k], ref[k]) < TOL, (k, rel(res[k], ref[k]))
    # basispursuittest.m:136-143: the engine's answer passes/fails the tester exactly as the oracle's does
    crit = lambda r: (np.sum(np.abs(r["xopt"])) <= np.sum(np.abs(testx)),
                      np.mean(np.abs(D @ r["xopt"] - s) / np.abs(s)) <= 1e-10)
    assert crit(res) == crit(ref)
    assert np.linalg.norm(D @ res["xopt"] - s) <= 1e-9 * np.linalg.norm(s)


def test_basispursuit_reference_errors(engine):
    from admm_project_b200 import MatlabError
    with pytest.raises(MatlabError, match="Square matrix problem"):
        basispursuit(np.eye(4), np.ones(4), {}, engine=engine)
    with pytest.raises(MatlabError, match="Overdetermined"):
        basispursuit(np.ones((6, 3)), np.ones(6), {}, engine=engine)
    with pytest.raises(MatlabError, match="must match the number of rows"):
        basispursuit(np.ones((3, 6)), np.ones(4), {}, engine=engine)


@pytest.mark.parametrize("rows,cols,nb", [(512, 128, 5), (3000, 700, 64), (1030, 257, 3)])
def test_lasso_lambda_batch_matches_per_lambda_oracle(engine, rows, cols, nb):
    D, s, lam_max10, _ = gen.lasso_problem(0, rows, cols)
    lam_max = lam_max10 * 10.0
    lams = lam_max * 10.0 ** (-np.arange(nb) / 21.0)
    engine.setup_lasso(D, s, 1.0)
    o = engine.default_options()
    o.reltol = 1e-4
    out = engine.solve_lasso_batch(o, lams)
    for j in range(nb):
        ref = oracle.lasso(D, s, float(lams[j]), {"reltol": 1e-4, "history": 0})
        assert out["steps"][j] == ref["steps"], (j, out["steps"][j], ref["steps"])
        for k in ("xopt", "zopt", "uopt"):
            assert rel(out[k][:, j], ref[k]) < TOL, (j, k)
        n = ref["steps"]
        assert rel(out["pnorm"][:n, j], ref["pnorm"]) < TOL
        assert rel(out["derr"][:n, j], ref["derr"]) < TOL


def test_lasso_batch_on_a_large_factor_uses_the_tall_triangular_gemm(engine):
    """n >= 2048: the two triangular products of the batch run on 256 x 64 tiles with uniform K chunks whose empty
    (row tile, chunk) units are skipped; every column must still reproduce the oracle's stand-alone run."""
    D, s, lam, _ = gen.lasso_problem(3, 5000, 2304)
    lams = (lam / 0.1) * 10.0 ** (-np.arange(9) / 4.0)
    engine.setup_lasso(D, s, 1.0)
    o = engine.default_options()
    rb = engine.solve_lasso_batch(o, lams)
    for j in (0, 4, 8):
        ref = oracle.lasso(D, s, lams[j], {"history": 0})
        assert rb["steps"][j] == ref["steps"]
        k = ref["steps"]
        for key in ("xopt", "zopt", "uopt"):
            assert rel(rb[key][:, j], ref[key]) < TOL, key
        for key in ("pnorm", "dnorm", "perr", "derr"):
            assert rel(rb[key][:k, j], ref[key]) < TOL, key
