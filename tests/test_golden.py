"""Golden fixtures (tests/golden/*.npz, frozen ORACLE outputs -- see make_golden.py for why they are
not reference outputs).  CPU: the oracle still reproduces them.  GPU: the engine matches them through
the C-ABI without importing the oracle."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
KEYS = ("xopt", "zopt", "uopt", "pnorm", "perr", "objevals")


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def run(name, g, mod, **kw):
    o = {"objevals": 1}
    if name.startswith("lasso"):
        o["relax"] = float(g["in_relax"])
        return mod.lasso(g["in_D"], g["in_s"], float(g["in_lam"]), o, **kw)
    if name.startswith("svm"):
        np.random.seed(int(g["in_seed"]))
        return mod.linearsvm(g["in_D"], g["in_ell"], float(g["in_C"]), o, **kw)
    if name.startswith("huber"):
        return mod.huberfit(g["in_D"], g["in_s"], dict(o, convtest=1), **kw)
    if name.startswith("lad"):
        return mod.lad(g["in_D"], g["in_s"], dict(o, convtest=1, relax=float(g["in_relax"])), **kw)
    if name.startswith("tv"):
        return mod.totalvariation(g["in_s"], float(g["in_lam"]), dict(o, maxiters=2000), **kw)
    if name.startswith("bp"):
        return mod.basispursuit(g["in_D"], g["in_s"], dict(o, maxiters=5000), **kw)
    if name.startswith("model"):
        return mod.model(g["in_P"], g["in_Q"], g["in_r"], g["in_s"], dict(o, relax=float(g["in_relax"])), **kw)
    raise AssertionError(name)


def test_fixtures_exist():
    assert len(FILES) == 8


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    import oracle
    g = np.load(path)
    res = run(os.path.basename(path), g, oracle)
    assert res["steps"] == int(g["steps"])
    for k in KEYS:
        assert rel(res[k], g[k]) < 1e-12, k


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_engine_matches_golden(engine, path):
    import admm_project_b200 as eng_pkg
    g = np.load(path)
    res = run(os.path.basename(path), g, eng_pkg, engine=engine)
    assert res["steps"] == int(g["steps"])
    for k in KEYS:
        assert rel(res[k], g[k]) < 1e-9, (k, rel(res[k], g[k]))
