"""CPU test of the host mirror of admm.m's option handling (admm.m:46-76, setopt :780-971): a recording
stand-in for the Engine receives the C options struct that admm() builds, so defaults, quirks and the result
struct assembly are checked without a GPU.  (The stand-in computes nothing: it is test scaffolding, not a
CPU path -- the product has none.)"""
import ctypes

import numpy as np
import pytest

from admm_project_b200 import _lib as L
from admm_project_b200 import EngineError, MatlabError, admm
from admm_project_b200.getproxops import EngineProx


class RecordingEngine:
    nranks, rank, row_range, m_total = 1, 0, None, None

    def __init__(self, n=4, steps=3, status=L.CONVERGED_STD):
        self.n, self.steps, self.status, self.seen, self.init = n, steps, status, None, None

    def default_options(self):
        o = L.Options()
        L.load().admm_b200_default_options(ctypes.byref(o))        # no GPU needed for the defaults
        return o

    def dims(self):
        return self.n, self.n, self.n

    def set_init(self, x0, z0, u0):
        self.init = (x0, z0, u0)

    def solve(self, o, want_history=True):
        self.seen = {k: getattr(o, k) for k, _ in L.Options._fields_}
        k, n = self.steps, self.n
        out = dict(steps=k, status=self.status, xopt=np.ones(n), zopt=np.ones(n), uopt=np.zeros(n), objopt=2.5,
                   setup_ms=0.0, loop_ms=0.0)
        for name in ("pnorm", "dnorm", "perr", "derr", "hnormsq", "objevals", "dvals", "avals", "restarted"):
            out[name] = np.arange(1.0, k + 1)
        if want_history and o.history:
            out["xvals"], out["zvals"], out["uvals"] = (np.ones((n, k)) for _ in range(3))
        return out


def run(options, **kw):
    eng = RecordingEngine(**kw)
    minx, minz = EngineProx("xminf", "lasso", "xminLASSO", eng, {}), EngineProx("zming", "lasso", "zminSoftThresholding", eng, {})
    base = dict(A=1, B=-1, c=0, m=eng.n, nA=eng.n, nB=eng.n)
    return admm(minx, minz, dict(base, **options)), eng


def test_defaults_are_admm_m_51_76():
    res, eng = run({})
    s = eng.seen
    assert (s["rho"], s["relax"], s["abstol"], s["reltol"], s["convtol"], s["hnormtol"]) == (1.0, 1.0, 1e-5, 1e-3, 1e-10, 1e-6)
    assert (s["maxiters"], s["domaxiters"], s["stopcond"], s["nodualerror"], s["convtest"], s["objevals"]) == (1000, 0, L.STOP_STANDARD, 0, 0, 0)
    assert (s["history"], s["check_every"], s["graph"], s["fast"]) == (1, 8, 1, 0)
    assert res["steps"] == 3 and res["xvals"].shape == (4, 3) and "Hnormsq" not in res and "objevals" not in res
    assert np.array_equal(res["x0"], np.zeros(4)) and eng.init == (None, None, None)        # admm.m:252-254


def test_quirks_of_setopt_and_the_loop_bounds():
    _, eng = run({"maxiters": 0})                                   # admm.m:334-339
    assert eng.seen["maxiters"] == 1000
    _, eng = run({"maxiters": 12.3})
    assert eng.seen["maxiters"] == 13
    _, eng = run({"stopcond": "nonsense"})                          # strcmp matches neither test: runs to maxiters
    assert eng.seen["domaxiters"] == 1
    res, eng = run({"stopcond": "both", "Hnormtol": 5.0, "Hreltol": 1e-9})     # setopt reads Hreltol, admm.m:927-928
    assert eng.seen["hnormtol"] == 1e-9 and eng.seen["stopcond"] == L.STOP_BOTH and res["Hnormtol"] == 1e-9
    assert "Hnormsq" in res and res["wvals"].shape == (12, 3)      # w = [x; z; rho*u], admm.m:679
    with pytest.raises(MatlabError, match="Hreltol"):
        run({"Hnormtol": 1.0})
    _, eng = run({"objevals": 1})                                   # no objective handle: nothing to evaluate
    assert eng.seen["objevals"] == 0
    res, eng = run({"objevals": 1, "obj": "engine"})
    assert eng.seen["objevals"] == 1 and res["objopt"] == 2.5 and res["objevals"].shape == (3,)


def test_fast_variants_and_history_switch():
    res, eng = run({"fast": 1})                                     # fasttype defaults to 'weak': accelerated, alg 2
    assert (eng.seen["fast"], eng.seen["fasttype"], eng.seen["restart"], eng.seen["dvaltol"]) == (1, 1, 0.999, 1e-8)
    assert res["pnorm"].size == 0 and "dvals" in res and "avals" in res and "perr" not in res
    res, eng = run({"fast": 1, "fasttype": "strong", "restart": 7})
    assert eng.seen["fasttype"] == 0 and "avals" in res and "perr" in res
    res, eng = run({"history": 0, "graph": 0, "check_every": 3})
    assert (eng.seen["history"], eng.seen["graph"], eng.seen["check_every"]) == (0, 0, 3) and "xvals" not in res


def test_divergence_return_leaves_the_struct_unfinished(capsys):
    res, _ = run({"convtest": 1}, steps=2, status=L.DIVERGED_RETURN)
    assert "steps" not in res and "xopt" not in res and "Hnormsq" in res       # admm.m:692-700
    assert "not be converging" in capsys.readouterr().out


def test_reference_error_texts_and_refusals():
    with pytest.raises(MatlabError, match="not a struct"):
        admm(None, None, 3)
    with pytest.raises(MatlabError, match="Must specify a matrix A"):
        eng = RecordingEngine()
        p = EngineProx("xminf", "lasso", "x", eng, {})
        admm(p, EngineProx("zming", "lasso", "z", eng, {}), {"B": -1, "c": 0})
    with pytest.raises(EngineError, match="adaptive"):
        run({"adaptive": 1})
    with pytest.raises(EngineError, match="consensus"):
        run({"parallel": "both"})
    # host handles the reference evaluates inside every iteration (admm.m:556-558, 612-616) are refused, not ignored
    with pytest.raises(EngineError, match="options.altu"):
        run({"altu": lambda u, ax, bz, c: u})
    with pytest.raises(EngineError, match="options.specialnorms"):
        run({"specialnorms": lambda x, z, u, rho: (0.0, 0.0)})
    calls = []
    res, eng = run({"preprocess": lambda: calls.append(1)})         # admm.m:473-476: called once, before the loop
    assert calls == [1] and res["steps"] == 3
    with pytest.raises(EngineError, match="different problems"):
        e = RecordingEngine()
        admm(EngineProx("xminf", "lasso", "x", e, {}), EngineProx("zming", "lad", "z", e, {}), {})
