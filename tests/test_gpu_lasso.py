"""GPU parity of the lasso path (solvers/lasso.m -> getproxops -> admm.m) against the oracle:
same iteration count, iterates within 1e-9 relative (BASELINE.json north_star), per-iteration
residual histories within 1e-9, and the reference tester's pass criterion (lassotest.m:143-147)."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import lasso
from admm_project_b200.generators import lasso_problem

pytestmark = pytest.mark.gpu
TOL = 1e-9      # north_star: iterates within 1e-9 relative error in FP64


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def compare(res, ref, hist=True):
    assert res["steps"] == ref["steps"], (res["steps"], ref["steps"])
    for k in ("xopt", "zopt", "uopt"):
        assert rel(res[k], ref[k]) < TOL, k
    for k in ("pnorm", "dnorm", "perr", "derr"):
        assert res[k].shape == ref[k].shape
        assert rel(res[k], ref[k]) < TOL, k
    if "objevals" in ref:
        assert rel(res["objevals"], ref["objevals"]) < TOL
        assert abs(res["objopt"] - ref["objopt"]) <= TOL * abs(ref["objopt"])
    if "Hnormsq" in ref:
        assert rel(res["Hnormsq"], ref["Hnormsq"]) < TOL, rel(res["Hnormsq"], ref["Hnormsq"])
        # elementwise too; the late entries are squared differences of nearly equal iterates (||dz|| ~ 1e-6 ||z||), so they
        # carry ~6 fewer digits than the iterates in ANY implementation
        assert np.allclose(res["Hnormsq"], ref["Hnormsq"], rtol=1e-7, atol=1e-22)
    if hist:
        for k in ("xvals", "zvals", "uvals"):
            assert res[k].shape == ref[k].shape, k
            assert rel(res[k], ref[k]) < TOL, k


@pytest.mark.parametrize("rows,cols", [(256, 64), (2048, 256), (150, 500), (1500, 5000), (333, 129)])
@pytest.mark.parametrize("relax", [1.0, 1.5])
def test_lasso_matches_oracle(engine, rows, cols, relax):
    D, s, lam, testx = lasso_problem(0, rows, cols)
    opts = {"objevals": 1, "relax": relax}                       # lassotest.m:131
    ref = oracle.lasso(D, s, lam, opts)
    res = lasso(D, s, lam, opts, engine=engine)
    compare(res, ref)
    # lassotest.m:143-147 pass criterion, evaluated on the engine's answer
    obj = lambda x: 0.5 * np.sum((D @ x - s) ** 2) + lam * np.sum(np.abs(x))
    assert obj(res["xopt"]) < obj(testx)


def test_lasso_8192x1024_tight_tolerance(engine):
    D, s, lam, _ = lasso_problem(1, 8192, 1024)
    opts = {"reltol": 1e-4, "history": 0}
    ref = oracle.lasso(D, s, lam, opts)
    res = lasso(D, s, lam, opts, engine=engine)
    compare(res, ref, hist=False)
    assert "xvals" not in res


@pytest.mark.parametrize("stopcond", ["hnorm", "both"])
def test_lasso_hnorm_stop_and_convtest(engine, stopcond):
    D, s, lam, _ = lasso_problem(2, 512, 128)
    opts = {"stopcond": stopcond, "convtest": 1, "Hreltol": 1e-9, "Hnormtol": 1}    # setopt quirk: reads Hreltol
    ref = oracle.lasso(D, s, lam, opts)
    res = lasso(D, s, lam, opts, engine=engine)
    compare(res, ref)
    assert res["Hnormtol"] == 1e-9
    assert res["wvals"].shape == ref["wvals"].shape
    assert rel(res["wvals"], ref["wvals"]) < TOL


def test_lasso_domaxiters_and_warm_start(engine):
    D, s, lam, _ = lasso_problem(3, 300, 100)
    rs = np.random.RandomState(9)
    opts = {"domaxiters": 1, "maxiters": 37, "x0": rs.randn(100), "z0": rs.randn(100), "u0": rs.randn(100),
            "rho": 2.5}
    ref = oracle.lasso(D, s, lam, opts)
    res = lasso(D, s, lam, opts, engine=engine)
    assert res["steps"] == 37
    compare(res, ref)


def test_lasso_device_resident_matrix(engine):
    torch = pytest.importorskip("torch")
    from admm_project_b200 import DeviceMatrix
    D, s, lam, _ = lasso_problem(4, 1024, 192)
    t = torch.from_numpy(np.ascontiguousarray(D.T)).cuda()          # row-major (n x m) == column-major m x n
    ref = oracle.lasso(D, s, lam, {})
    res = lasso(DeviceMatrix(t.data_ptr(), 1024, 192, 1024, keepalive=t), s, lam, {}, engine=engine)
    compare(res, ref)


def test_lasso_errors_match_reference_messages(engine):
    from admm_project_b200 import EngineError, MatlabError
    D, s, lam, _ = lasso_problem(5, 64, 16)
    with pytest.raises(MatlabError, match="nonnegative real number"):
        lasso(D, s, -1.0, {}, engine=engine)
    with pytest.raises(MatlabError, match="positive real number"):
        lasso(D, s, lam, {"rho": 0}, engine=engine)
    with pytest.raises(MatlabError, match="Hreltol"):
        lasso(D, s, lam, {"Hnormtol": 1e-3}, engine=engine)
    with pytest.raises(EngineError):
        lasso(D, s, lam, {"parallel": "both"}, engine=engine)


def test_lasso_tall_host_matrix_pipelined_upload(engine):
    """m >= 32768 with a HOST matrix takes the panel-pipelined path (H2D of row panels overlapped with
    the accumulation of D_p'D_p and D_p's_p); the result must not depend on it."""
    D, s, lam, _ = lasso_problem(6, 40000, 48)      # 3 panels: 16384 + 16384 + 7232 rows
    ref = oracle.lasso(D, s, lam, {"history": 0, "objevals": 1})
    res = lasso(D, s, lam, {"history": 0, "objevals": 1}, engine=engine)
    compare(res, ref, hist=False)
    G = D.T @ D + np.eye(48)
    assert rel(engine.get_factor(), np.linalg.cholesky(G)) < 1e-12


def test_host_upload_paths_agree(engine, monkeypatch):
    """A tall HOST matrix goes up in row panels pipelined with the Gram (engine.cu:setup_lasso), and from pageable
    memory each panel is gathered by host threads into a pinned staging ring first (upload_rows).  Factor and
    solution must not depend on where D lived: pageable NumPy array (staged), the same array with staging switched off
    (driver bounce buffer), and a device-resident copy; ragged sizes on purpose."""
    import torch
    from admm_project_b200 import DeviceMatrix
    from admm_project_b200.generators import lasso_problem_big
    m, n = 18433, 2051                                           # 302 MB: above the panel / staging thresholds; odd m and n
    D, s, lam, _ = lasso_problem_big(7, m, n)
    opts = {"reltol": 1e-4, "history": 0}
    ref = oracle.lasso(D, s, lam, opts)
    res = lasso(D, s, lam, opts, engine=engine)                   # pageable: pinned staging ring
    compare(res, ref, hist=False)
    L_staged = engine.get_factor()
    ld = m + 1                                                    # device matrices need an even leading dimension
    Dt = torch.zeros(n, ld, dtype=torch.float64, device="cuda")   # row-major n x ld == column-major ld x n on the device
    Dt[:, :m] = torch.from_numpy(np.ascontiguousarray(D.T)).cuda()
    dev = lasso(DeviceMatrix(Dt.data_ptr(), m, n, ld, keepalive=Dt), s, lam, opts, engine=engine)
    compare(dev, ref, hist=False)
    L_dev = engine.get_factor()
    assert rel(L_staged, L_dev) < 1e-12
    monkeypatch.setenv("ADMM_B200_NO_PINNED_STAGING", "1")       # read at every upload
    plain = lasso(D, s, lam, opts, engine=engine)
    compare(plain, ref, hist=False)
    assert rel(engine.get_factor(), L_dev) < 1e-12
