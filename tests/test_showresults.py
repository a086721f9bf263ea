"""The reporting edge (showresults.m:13-411): the results struct the engine / oracle fill carries every field the
reference's showresults reads, the text report follows the reference line for line, the plot plan follows its subplot
rules, and the .mat round trip hands MATLAB-shaped structs to the reference's own showresults.m."""
import io

import numpy as np
import pytest

import oracle
from admm_project_b200 import MatlabError, testers
from admm_project_b200.showresults import load_mat, num2str, save_mat, showresults

READ_BY_SHOWRESULTS = ("steps", "runtime", "solverruntime", "xopt", "options", "objevals", "pnorm", "perr", "dnorm",
                       "derr")                                        # showresults.m:139-154,169,201-233,297-298,330-331
WITH_CONVTEST = ("Hnormsq", "Hnormtol")                               # :218-221, 270-272: only when the H-norm was evaluated


def test_num2str_follows_matlab():
    assert num2str(3) == "3" and num2str(-12.0) == "-12"
    assert num2str(np.pi) == "3.1416"
    assert num2str(123.456) == "123.456"
    assert num2str(0.000012345678) == "1.2346e-05"
    assert num2str(1234567.891) == "1234567.891"
    assert num2str(float("nan")) == "NaN" and num2str(float("inf")) == "Inf"


def check_report(results, test, options, header):
    buf = io.StringIO()
    rep = showresults(results, test, options, file=buf)
    text = buf.getvalue().splitlines()
    assert text == rep["lines"]
    assert text[0] == " " and text[1] == header
    assert "Number of iteration steps performed: %d" % results["steps"] in text
    assert any(t.startswith("Runtime of ADMM call: ") and t.endswith(" seconds.") for t in text)
    assert any(t.startswith("Overall runtime of solver: ") for t in text)
    assert ("Test successful!" in text) == (not test["failed"])
    return rep


def test_lasso_report_and_plot_plan_from_the_tester():
    results, test = testers.lassotest(seed=3, rows=96, cols=32, quiet=1, solvers=oracle, options={"convtest": 1})
    for k in READ_BY_SHOWRESULTS + WITH_CONVTEST:
        assert k in results, k
    rep = check_report(results, test, {"solver": "lasso"}, "LASSO EXECUTION AND TEST RESULTS ---")
    assert "True optimal objective value: " + num2str(test["trueobjopt"]) in rep["lines"]
    assert "ADMM's optimal objective value for (x, z): " + num2str(test["admmopt"]) in rep["lines"]
    N = results["steps"]
    kinds = [p["kind"] for p in rep["plots"]]
    assert kinds == ["signal", "objective", "hnorm", "pnorm", "dnorm"]          # D, s, testx present -> the signal figure
    assert rep["nplots"] == 4
    sub = [p["subplot"] for p in rep["plots"][1:]]
    assert sub == [1, 2, 3, 4]
    assert [p["xlabel"] for p in rep["plots"][1:]] == [None, None, None, "Iteration k"]      # only the last subplot
    obj = rep["plots"][1]
    assert obj["series"][0][0] == "True optimal objective value" and np.all(obj["series"][0][2] == test["trueobjopt"])
    assert obj["series"][1][2].shape == (N,)
    hn = rep["plots"][2]
    assert np.all(hn["series"][0][2] >= 1e-8) and np.all(hn["series"][1][2] == results["Hnormtol"])
    pn = rep["plots"][3]
    assert np.array_equal(pn["series"][1][2], np.asarray(results["perr"]).reshape(-1))


def test_quiet_zero_runs_the_report_like_the_reference(capsys):
    results, test = testers.huberfittest(seed=1, rows=200, cols=16, quiet=0, solvers=oracle)
    out = capsys.readouterr().out
    assert "HUBER FITTING EXECUTION AND TEST RESULTS ---" in out
    assert test["report"]["nplots"] == 4
    _, t2 = testers.linearsvmtest(seed=2, mpos=40, mneg=40, quiet=0, solvers=oracle)
    out = capsys.readouterr().out
    assert "ADMM EXECUTION AND TEST RESULTS FOR HINGE ---" in out and "ADMM EXECUTION AND TEST RESULTS FOR 01 ---" in out
    assert t2["reports"][0]["plots"][1]["title"].endswith("for hinge loss function")       # options.tester == 'linearsvm'


def test_reference_quirks():
    res = {"steps": 2, "xopt": np.zeros(3), "options": {}, "runtime": 0.5}
    with pytest.raises(MatlabError, match="No structs given"):
        showresults()
    with pytest.raises(MatlabError, match="Undefined function or variable 'solver'"):          # showresults.m:65-66
        showresults(res, {}, {"solver": "unwrappedadmm"}, file=io.StringIO())
    rep = showresults(res, {"testobj": 7.0, "testobjx": 9.0}, {"solver": "nonsense"}, file=io.StringIO())
    assert rep["lines"][1] == "ADMM EXECUTION AND TEST RESULTS ---"
    assert "Test's original objective value for x: 7" in rep["lines"]                        # :93-96 prints test.testobj
    assert rep["nplots"] == 0 and rep["plots"] == []
    # fast / weak: three subplots at most (:202-210); the d-value plot then lands beyond them, as in the reference
    r = oracle.lasso(*_small_lasso(), {"fast": 1, "fasttype": "weak", "objevals": 1})
    assert "dvals" in r and "dvaltol" in r
    ro = r.get("options") or {}
    if ro.get("algorithm") == "fast":
        assert showresults(r, {}, {}, file=io.StringIO())["nplots"] <= 3


def _small_lasso():
    from admm_project_b200.generators import lasso_problem
    D, s, lam, _ = lasso_problem(5, 64, 24)
    return D, s, lam


def test_mat_round_trip(tmp_path):
    results, test = testers.lassotest(seed=4, rows=80, cols=24, quiet=1, solvers=oracle)
    path = str(tmp_path / "run.mat")
    save_mat(path, results, test, {"solver": "lasso"})
    back = load_mat(path)
    r = back["results"]
    for k in READ_BY_SHOWRESULTS:
        assert k in r, k
    assert int(r["steps"]) == results["steps"]
    assert np.allclose(r["xopt"], results["xopt"]) and np.allclose(r["pnorm"], results["pnorm"])
    assert isinstance(r["options"], dict)
    assert back["options"]["solver"] == "lasso" and int(back["test"]["failed"]) == test["failed"]
    from scipy.io import loadmat
    raw = loadmat(path)                                  # MATLAB shapes: iterates and histories are COLUMN vectors
    assert raw["results"]["xopt"][0, 0].shape == (24, 1)
    assert raw["results"]["pnorm"][0, 0].shape == (results["steps"], 1)
    # the report from the reloaded structs is the report from the originals
    a = showresults(results, test, {"solver": "lasso"}, file=io.StringIO())["lines"]
    b = showresults(back["results"], back["test"], back["options"], file=io.StringIO())["lines"]
    assert a == b


@pytest.mark.gpu
def test_engine_results_struct_feeds_showresults(engine, tmp_path):
    results, test = testers.lassotest(seed=3, rows=96, cols=32, quiet=1, engine=engine, options={"convtest": 1})
    ref, _ = testers.lassotest(seed=3, rows=96, cols=32, quiet=1, solvers=oracle, options={"convtest": 1})
    for k in READ_BY_SHOWRESULTS + WITH_CONVTEST:
        assert k in results, k
    rep = check_report(results, test, {"solver": "lasso"}, "LASSO EXECUTION AND TEST RESULTS ---")
    assert [p["kind"] for p in rep["plots"]] == ["signal", "objective", "hnorm", "pnorm", "dnorm"]
    ref_rep = showresults(ref, _, {"solver": "lasso"}, file=io.StringIO())
    for p, q in zip(rep["plots"], ref_rep["plots"]):
        for (la, xa, ya, _sa), (lb, xb, yb, _sb) in zip(p["series"], q["series"]):
            assert la == lb and np.allclose(ya, yb, rtol=1e-6, atol=1e-12)
    path = str(tmp_path / "gpu.mat")
    save_mat(path, results, test, {"solver": "lasso"})
    assert int(load_mat(path)["results"]["steps"]) == ref["steps"]
