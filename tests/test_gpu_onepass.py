"""The single-pass A = D iteration (csrc/onepass.cuh: D*x, prox and D'*[rhs, dz, u] from ONE read of D
through a shared-memory tile) against the oracle and against the two-pass kernels, at shapes that
exercise every tile height (32 / 24 / 16 rows), ragged last tiles, odd row counts, relaxation, the
dual-residual (3-vector) and nodualerror (1-vector) forms and the iterate history."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import huberfit, lad, linearsvm
from admm_project_b200 import generators as gen

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def compare(res, ref, keys=("xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr", "objevals", "xvals", "zvals", "uvals")):
    assert res["steps"] == ref["steps"], (res["steps"], ref["steps"])
    for k in keys:
        if k in ref and k in res:
            ok = ~np.isnan(np.asarray(ref[k], dtype=float))
            assert rel(np.asarray(res[k])[ok], np.asarray(ref[k])[ok]) < TOL, (k, rel(np.asarray(res[k])[ok], np.asarray(ref[k])[ok]))


@pytest.fixture
def force(monkeypatch):
    monkeypatch.setenv("ADMM_B200_FORCE_ONEPASS", "1")


@pytest.mark.parametrize("rows,cols", [(333, 20), (2049, 130), (4000, 64), (31, 7)])
@pytest.mark.parametrize("relax", [1.0, 1.6])
def test_huber_onepass_matches_oracle(engine, force, rows, cols, relax):
    D, s, _ = gen.huber_problem(0, rows, cols)
    opts = {"objevals": 1, "convtest": 1, "relax": relax}
    compare(huberfit(D, s, opts, engine=engine), oracle.huberfit(D, s, opts))


@pytest.mark.parametrize("rows,cols,relax", [(777, 20, 1.4), (1500, 200, 1.0)])
def test_lad_onepass_matches_oracle(engine, force, rows, cols, relax):
    D, s, _ = gen.lad_problem(0, rows, cols)
    opts = {"objevals": 1, "convtest": 1, "relax": relax}
    compare(lad(D, s, opts, engine=engine), oracle.lad(D, s, opts))


@pytest.mark.parametrize("mpos,mneg", [(257, 130), (2048, 2048)])
def test_svm_onepass_matches_oracle(engine, force, mpos, mneg):
    D, ell = gen.svm_problem(0, mpos, mneg)
    opts = {"objevals": 1, "convtest": 1}
    np.random.seed(11)
    ref = oracle.linearsvm(D, ell, 0.5, opts)
    np.random.seed(11)
    compare(linearsvm(D, ell, 0.5, opts, engine=engine), ref)


@pytest.mark.parametrize("cols,tile", [(700, 32), (1000, 24), (1400, 16)])
def test_every_tile_height(engine, force, cols, tile):
    # 32 rows while the n-column tile fits 227 KB of shared memory, then 24, then 16 (onepass.cuh)
    D, s, _ = gen.huber_problem(1, 2 * cols + 37, cols)
    opts = {"history": 0, "maxiters": 40}
    compare(huberfit(D, s, opts, engine=engine), oracle.huberfit(D, s, opts))


def test_onepass_is_the_default_at_size_and_agrees_with_two_pass(engine, monkeypatch):
    D, s, _ = gen.huber_problem(2, 20000, 128)                  # 2.56 M elements, n >= 128: single-pass by default
    opts = {"history": 0, "objevals": 1}
    l0 = engine.launch_count()
    one = huberfit(D, s, opts, engine=engine)
    l1 = engine.launch_count()
    monkeypatch.setenv("ADMM_B200_NO_ONEPASS", "1")
    two = huberfit(D, s, opts, engine=engine)
    l2 = engine.launch_count()
    assert one["steps"] == two["steps"]
    assert (l1 - l0) != (l2 - l1)                               # a different kernel sequence really ran
    for k in ("xopt", "zopt", "uopt", "pnorm", "dnorm", "objevals"):
        assert rel(one[k], two[k]) < 1e-11, k
    compare(one, oracle.huberfit(D, s, opts))
