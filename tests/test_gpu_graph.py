"""options.graph: from the third burst of check_every iterations on, the loop is replayed as one CUDA
graph launch per burst (engine.cu:run_bursts).  The kernels and their arguments are the same, so the
results must be BITWISE those of the eager loop -- and therefore the oracle parity of the other GPU
tests carries over -- and graph launches must really have happened."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import basispursuit, huberfit, lad, lasso, linearsvm, totalvariation
from admm_project_b200 import generators as gen

pytestmark = pytest.mark.gpu
KEYS = ("xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr")


def both(engine, run, opts):
    r0 = engine.graph_replays()
    eager = run(dict(opts, graph=0))
    assert engine.graph_replays() == r0                       # graph = 0 really is the eager loop
    graphed = run(dict(opts, graph=1))
    return eager, graphed, engine.graph_replays() - r0


def same(a, b, keys=KEYS):
    assert a["steps"] == b["steps"]
    for k in keys:
        if k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("check_every", [1, 4, 8])
def test_lasso_graph_is_bitwise_the_eager_loop(engine, check_every):
    D, s, lam, _ = gen.lasso_problem(3, 1500, 5000)            # lassotest.m shape (fat: Woodbury x-update)
    opts = {"objevals": 1, "check_every": check_every}
    eager, graphed, replays = both(engine, lambda o: lasso(D, s, lam, o, engine=engine), opts)
    same(eager, graphed, KEYS + ("objevals", "xvals", "zvals", "uvals"))
    assert replays == max(0, -(-graphed["steps"] // check_every) - 2)   # every burst after the first two
    ref = oracle.lasso(D, s, lam, {"objevals": 1})
    assert graphed["steps"] == ref["steps"]
    assert np.linalg.norm(graphed["xopt"] - ref["xopt"]) <= 1e-9 * np.linalg.norm(ref["xopt"])


def test_lasso_graph_partial_last_burst_and_maxiters(engine):
    D, s, lam, _ = gen.lasso_problem(4, 600, 200)
    opts = {"domaxiters": 1, "maxiters": 37, "check_every": 8}    # 2 eager + 2 graph bursts + 5 eager iterations
    eager, graphed, replays = both(engine, lambda o: lasso(D, s, lam, o, engine=engine), opts)
    assert graphed["steps"] == 37 and replays == 2
    same(eager, graphed, KEYS + ("xvals", "zvals", "uvals"))


def test_fast_admm_graph(engine):
    D, s, lam, _ = gen.lasso_problem(5, 800, 300)
    for fasttype in ("weak", "strong"):
        opts = {"fast": 1, "fasttype": fasttype, "maxiters": 40, "domaxiters": 1, "check_every": 4}
        eager, graphed, replays = both(engine, lambda o: lasso(D, s, lam, o, engine=engine), opts)
        assert replays == -(-graphed["steps"] // 4) - 2 and replays >= 1
        same(eager, graphed, KEYS + ("avals", "dvals"))


def test_totalvariation_graph_needs_an_even_burst(engine):
    s, _ = gen.tv_problem(1, 40000)
    run = lambda o: totalvariation(s, 3.0, o, engine=engine)
    eager, graphed, replays = both(engine, run, {"maxiters": 200, "history": 0, "check_every": 6})
    assert replays >= 1
    same(eager, graphed)
    # an odd burst would replay the wrong z/u halves: the engine stays eager
    eager, graphed, replays = both(engine, run, {"maxiters": 200, "history": 0, "check_every": 5})
    assert replays == 0
    same(eager, graphed)


def test_unwrapped_family_graph(engine):
    D, ell = gen.svm_problem(2, 700, 650)
    def run(o):
        np.random.seed(5)                                      # unwrappedadmm.m:87-89 draws x0, z0, u0
        return linearsvm(D, ell, 0.5, o, engine=engine)
    eager, graphed, replays = both(engine, run, {"check_every": 4, "history": 0})
    assert replays >= 1
    same(eager, graphed)
    Dh, sh, _ = gen.huber_problem(1, 3000, 50)
    eager, graphed, replays = both(engine, lambda o: huberfit(Dh, sh, o, engine=engine), {"check_every": 2, "relax": 1.5})
    assert replays >= 1
    same(eager, graphed, KEYS + ("xvals", "zvals", "uvals"))
    Dl, sl, _ = gen.lad_problem(1, 900, 30)
    eager, graphed, replays = both(engine, lambda o: lad(Dl, sl, o, engine=engine), {"check_every": 3, "convtest": 1})
    assert replays >= 1
    same(eager, graphed)


def test_basispursuit_and_batches_graph(engine):
    D, s, _ = gen.bp_problem(0, 64, 256)
    eager, graphed, replays = both(engine, lambda o: basispursuit(D, s, o, engine=engine), {"check_every": 4})
    assert replays >= 1
    same(eager, graphed)
    # lambda batch
    Dl, sl, lam10, _ = gen.lasso_problem(0, 512, 128)
    lams = lam10 * 10.0 * 10.0 ** (-np.arange(6) / 21.0)
    engine.setup_lasso(Dl, sl, 1.0)
    outs = []
    for g in (0, 1):
        o = engine.default_options()
        o.reltol, o.check_every, o.graph = 1e-4, 4, g
        r0 = engine.graph_replays()
        outs.append(engine.solve_lasso_batch(o, lams))
        assert (engine.graph_replays() > r0) == bool(g)
    for k in ("steps", "xopt", "zopt", "uopt", "pnorm"):
        assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k
