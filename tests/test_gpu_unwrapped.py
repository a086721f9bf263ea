"""GPU parity of the A = D problems (linear SVM via unwrapped ADMM / transpose reduction, Huber
fitting, LAD) against the oracle: same iteration count, iterates and histories within 1e-9."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import huberfit, lad, linearsvm
from admm_project_b200 import generators as gen

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def compare(res, ref, hist=True, tol=TOL):
    assert res["steps"] == ref["steps"], (res["steps"], ref["steps"])
    for k in ("xopt", "zopt", "uopt"):
        assert rel(res[k], ref[k]) < tol, (k, rel(res[k], ref[k]))
    for k in ("pnorm", "perr"):
        assert rel(res[k], ref[k]) < tol, k
    for k in ("dnorm", "derr"):
        assert np.array_equal(np.isnan(res[k]), np.isnan(ref[k])), k
        ok = ~np.isnan(ref[k])
        assert rel(res[k][ok], ref[k][ok]) < tol, k
    if "objevals" in ref:
        assert rel(res["objevals"], ref["objevals"]) < tol
        assert abs(res["objopt"] - ref["objopt"]) <= tol * max(abs(ref["objopt"]), 1e-300)
    if "Hnormsq" in ref:
        assert rel(res["Hnormsq"], ref["Hnormsq"]) < TOL, rel(res["Hnormsq"], ref["Hnormsq"])
        # elementwise too; the late entries are squared differences of nearly equal iterates (||dz|| ~ 1e-6 ||z||), so they
        # carry ~6 fewer digits than the iterates in ANY implementation
        assert np.allclose(res["Hnormsq"], ref["Hnormsq"], rtol=1e-6, atol=1e-20)
    if hist:
        for k in ("xvals", "zvals", "uvals"):
            assert res[k].shape == ref[k].shape, k
            assert rel(res[k], ref[k]) < tol, k


@pytest.mark.parametrize("mpos,mneg", [(128, 128), (257, 130), (2048, 2048)])
def test_linearsvm_hinge_matches_oracle(engine, mpos, mneg):
    D, ell = gen.svm_problem(0, mpos, mneg)
    opts = {"objevals": 1, "convtest": 1}                       # linearsvmtest.m:148-149
    np.random.seed(11)
    ref = oracle.linearsvm(D, ell, 0.5, opts)                   # serial pinv path of the reference
    np.random.seed(11)
    res = linearsvm(D, ell, 0.5, opts, engine=engine)
    compare(res, ref)
    x = res["xopt"]
    if mpos == mneg:                                            # the tester's geometry is symmetric
        assert abs(1 + x[1] / x[0]) <= 0.05                     # linearsvmtest.m:180-192, errtol 0.05
    o = res["options"]
    assert o["maxiters"] == 1000 and o["stopcond"] == "both" and o["nodualerror"] == 1   # unwrappedadmm.m:81-92
    assert np.all(np.isnan(res["dnorm"]))


def test_linearsvm_parallel_option_equals_transpose_reduction_oracle(engine):
    D, ell = gen.svm_problem(1, 300, 300)
    np.random.seed(5)
    ref = oracle.linearsvm(D, ell, 0.5, {"parallel": "both", "workers": 4, "objevals": 1})
    np.random.seed(5)
    res = linearsvm(D, ell, 0.5, {"parallel": "both", "objevals": 1}, engine=engine)
    compare(res, ref)


def test_linearsvm_mnist_shaped(engine):
    D, ell = gen.svm_mnist_like(0, 6000, 96, nclass=3)
    D = D + 1e-3 * np.random.RandomState(1).randn(*D.shape)      # full column rank
    np.random.seed(3)
    ref = oracle.linearsvm(D, ell[:, 1], 0.5, {"parallel": "both", "workers": 2, "history": 0})
    np.random.seed(3)
    res = linearsvm(D, ell[:, 1], 0.5, {"history": 0}, engine=engine)
    compare(res, ref, hist=False)


def test_linearsvm_01_loss_trips_divergence_return(engine, capsys):
    D, ell = gen.svm_problem(2, 64, 64)
    opts = {"lossfunction": "01", "convtest": 1}
    np.random.seed(7)
    ref = oracle.linearsvm(D, ell, 0.5, opts)
    np.random.seed(7)
    res = linearsvm(D, ell, 0.5, opts, engine=engine)
    assert ("steps" in res) == ("steps" in ref)
    assert len(res["Hnormsq"]) == len(ref["Hnormsq"])
    assert rel(res["pnorm"], ref["pnorm"]) < TOL
    if "steps" not in ref:                                       # admm.m:692-700 early return
        assert "xopt" not in res and "runtime" not in res
        assert "ADMM seems to not be converging" in capsys.readouterr().out


def test_linearsvm_relax_is_a_dimension_error_like_the_reference(engine):
    from admm_project_b200 import EngineError
    D, ell = gen.svm_problem(3, 32, 32)
    with pytest.raises(EngineError, match="dimensions must agree"):
        linearsvm(D, ell, 0.5, {"relax": 1.5}, engine=engine)


@pytest.mark.parametrize("rows,cols", [(2048, 128), (1001, 37), (20000, 64)])
@pytest.mark.parametrize("relax", [1.0, 1.6])
def test_huberfit_matches_oracle(engine, rows, cols, relax):
    D, s, testx = gen.huber_problem(0, rows, cols)
    opts = {"objevals": 1, "convtest": 1, "relax": relax}       # huberfittest.m:137-139
    ref = oracle.huberfit(D, s, opts)
    res = huberfit(D, s, opts, engine=engine)
    compare(res, ref, hist=rows <= 2048)
    f = lambda x: 0.5 * np.sum(oracle.huber(D @ x - s))
    assert f(res["xopt"]) <= f(testx)                           # huberfittest.m:154-158


@pytest.mark.parametrize("rows,cols", [(1024, 128), (777, 20)])
@pytest.mark.parametrize("relax", [1.0, 1.4])
def test_lad_matches_oracle(engine, rows, cols, relax):
    D, s, xtrue = gen.lad_problem(0, rows, cols)
    opts = {"objevals": 1, "convtest": 1, "relax": relax}       # ladtest.m:130-132
    ref = oracle.lad(D, s, opts)
    res = lad(D, s, opts, engine=engine)
    compare(res, ref)


def test_robustfit_reference_error_messages(engine):
    from admm_project_b200 import MatlabError
    D, s, _ = gen.lad_problem(0, 64, 4)
    with pytest.raises(MatlabError, match="do not match size of s"):
        lad(D, s[:-1], {}, engine=engine)
    with pytest.raises(MatlabError, match="sizes incompatible"):
        linearsvm(D, s[:-1], 0.5, {}, engine=engine)
    with pytest.raises(MatlabError, match="nonnegative number"):
        linearsvm(D, np.sign(s), -1, {}, engine=engine)
