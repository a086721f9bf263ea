"""CPU tests: the C-ABI library loads and exports every symbol include/admm_b200.h declares, the
host mirror reproduces the reference's argument checks, and the product never imports oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "admm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(admm_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from admm_project_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libadmm_b200.so does not export " + n
        assert n in _lib.SYMBOLS, "ctypes binding missing for " + n
    assert sorted(_lib.SYMBOLS) == names
    assert lib.admm_b200_version() == 200


def test_struct_layouts_match_header():
    from admm_project_b200 import _lib
    # options: 6 doubles + int64 + 10 int32 + 2 doubles + 2 int32; result: see header
    assert ctypes.sizeof(_lib.Options) == 6 * 8 + 8 + 10 * 4 + 2 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.Result) == 8 + 4 + 4 + 3 * 8 + 15 * 8
    o = _lib.Options()
    _lib.load().admm_b200_default_options(ctypes.byref(o))          # admm.m:51-76
    assert (o.rho, o.relax, o.abstol, o.reltol, o.convtol, o.hnormtol) == (1.0, 1.0, 1e-5, 1e-3, 1e-10, 1e-6)
    assert (o.maxiters, o.domaxiters, o.stopcond, o.nodualerror, o.convtest, o.objevals) == (1000, 0, 0, 0, 0, 0)
    assert (o.fast, o.fasttype, o.restart, o.dvaltol) == (0, 1, 0.999, 1e-8)              # admm.m:59-60, 287, 291
    assert (o.check_every, o.graph) == (8, 1)


def test_slicemaker_through_cabi_matches_oracle():
    import oracle
    from admm_project_b200 import slicemaker
    from admm_project_b200.errorcheck import errorcheck
    for length, workers in [(10, 3), (8, 4), (60000, 8), (7, 8), (4194304, 8), (1, 1)]:
        assert slicemaker(length, workers) == oracle.slicemaker(0, workers, length)
    for spec in (0, 3, 5, [6, 4]):
        assert errorcheck(spec, "slices", "s", {"workers": 2, "slicelength": 10}) == \
            oracle.errorcheck(spec, "slices", "s", {"workers": 2, "slicelength": 10})


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from admm_project_b200 import Engine, EngineError, lasso
    with pytest.raises(EngineError, match="no CPU fallback"):
        Engine(0)
    with pytest.raises(EngineError):
        lasso(np.eye(4), np.ones(4), 0.1, {})


def test_host_mirror_argument_checks():
    from admm_project_b200 import EngineError, MatlabError, admm, getproxops, lasso, setopt
    with pytest.raises(MatlabError, match="nonnegative real number"):
        lasso(np.eye(4), np.ones(4), -1, {})
    with pytest.raises(MatlabError, match="do not match size of s"):
        lasso(np.eye(4), np.ones(5), 0.1, {})
    with pytest.raises(MatlabError, match="not a struct"):
        lasso(np.eye(4), np.ones(4), 0.1, None)
    with pytest.raises(MatlabError, match="not a solver"):
        getproxops("nosuchproblem", {})
    with pytest.raises(MatlabError, match="not a string"):
        getproxops(3, {})
    with pytest.raises(EngineError, match="outside the engine"):
        getproxops("CovarianceSelection", {})
    with pytest.raises(EngineError, match="CPU path"):
        admm(lambda x, z, u, r: x, lambda x, z, u, r: z, {})
    assert setopt({"Hnormtol": 1, "Hreltol": 2.0}, "Hnormtol", 1e-6) == 2.0


def test_model_mirror_argument_checks_run_before_any_device_work():
    # solvers/model.m:158-218 -- the reference's errorcheck texts, raised by the mirror without a GPU
    from admm_project_b200 import EngineError, MatlabError, getproxops, model
    P, Q, r, s = np.ones((6, 3)), np.ones((6, 3)), np.ones(6), np.ones(6)
    with pytest.raises(MatlabError, match="rows in P do not match number of rows in Q"):
        model(P, Q[:4], r, s, {})
    with pytest.raises(MatlabError, match="columns in P do not match number of columns in Q"):
        model(P, Q[:, :2], r, s, {})
    with pytest.raises(MatlabError, match="does not match length of vector r"):
        model(P, Q, r[:5], s, {})
    with pytest.raises(MatlabError, match="does not match length of vector s"):
        model(P, Q, r, s[:5], {})
    with pytest.raises(MatlabError, match="Argument r is not a vector"):
        model(P, Q, np.ones((6, 2)), s, {})
    with pytest.raises(MatlabError, match="not a struct"):
        model(P, Q, r, s, 3)
    with pytest.raises(EngineError, match="device-resident"):
        getproxops("Model", {"PtP": P.T @ P})                   # no engine: there is no CPU path


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "admm_project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
    code = "import sys; import admm_project_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
