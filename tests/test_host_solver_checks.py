"""CPU test: every solver mirror raises the reference's errorcheck texts (solvers/<name>.m local errorcheck
functions) BEFORE any device work -- so without a GPU the message is the reference's, not the engine's
"no CPU fallback".  The oracle raises the same texts (parity of the error behaviour)."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import (MatlabError, basispursuit, huberfit, lad, lasso, linearsvm, linearsvm_onevsall,
                               quadraticprogram, totalvariation, unwrappedadmm)
from admm_project_b200.errorcheck import errorcheck

D = np.arange(12.0).reshape(4, 3)
CASES = [
    ("lasso", lambda S: S.lasso(D, np.ones(5), 0.1, {}), "do not match size of s"),
    ("lasso", lambda S: S.lasso(D, np.ones(4), -0.1, {}), "lambda is not a nonnegative real number"),
    ("lasso", lambda S: S.lasso(D, np.ones(4), 0.1, {"rho": -1}), "options.rho is not a positive real number"),
    ("huberfit", lambda S: S.huberfit(D, np.ones(3), {}), "do not match size of s"),
    ("lad", lambda S: S.lad(D, np.ones(3), {}), "do not match size of s"),
    ("linearsvm", lambda S: S.linearsvm(D, np.ones(3), 0.5, {}), "Product ell\\*D is not possible"),
    ("linearsvm", lambda S: S.linearsvm(D, np.ones(4), -1.0, {}), "C is not a nonnegative number"),
    ("linearsvm", lambda S: S.linearsvm(D[:, :1], np.array([1.0, -1, 1, -1]), 0.5, {}), "Number of rows in At"),
    ("linearsvm", lambda S: S.linearsvm(np.ones((1, 1)), np.ones(1), 0.5, {}), "Given scalar as matrix A"),
    ("totalvariation", lambda S: S.totalvariation(np.ones(1), 1.0, {}), "Given scalar as matrix A"),
    ("totalvariation", lambda S: S.totalvariation(np.ones(5), -2.0, {}), "lambda parameter is not a nonnegative number"),
    ("totalvariation", lambda S: S.totalvariation(np.ones((3, 3)), 1.0, {}), "Argument s is not a vector"),
    ("basispursuit", lambda S: S.basispursuit(np.eye(3), np.ones(3), {}), "Square matrix problem"),
    ("basispursuit", lambda S: S.basispursuit(D, np.ones(4), {}), "Overdetermined system"),
    ("basispursuit", lambda S: S.basispursuit(D.T, np.ones(4), {}), "must match the number of rows in signal vector s"),
    ("quadraticprogram", lambda S: S.quadraticprogram(np.eye(3), np.ones(2), 0.0, np.zeros(3), np.ones(3), {}),
     "square matrix P and vector q do not match"),
    ("quadraticprogram", lambda S: S.quadraticprogram(np.eye(3), np.ones(3), 0.0, np.zeros(3), np.ones(2), {}),
     "Lengths of lower and upper bound"),
    ("quadraticprogram", lambda S: S.quadraticprogram(np.eye(3), np.ones(3), 0.0, np.array([0, 2.0, 0]), np.array([1, 1.0, 1]), {}),
     "do not specify an upper and lower bound"),
]


@pytest.mark.parametrize("name,call,text", CASES, ids=["%s-%d" % (c[0], i) for i, c in enumerate(CASES)])
def test_mirror_and_oracle_raise_the_reference_text(name, call, text):
    import admm_project_b200 as pkg
    with pytest.raises(MatlabError, match=text):
        call(pkg)
    # lasso.m has no size check of its own: MATLAB's `D'*s` raises its built-in dimension error there, NumPy's
    # matmul does in the oracle; the mirror words it like huberfit.m / lad.m do
    with pytest.raises((oracle.MatlabError, ValueError), match=text if name != "lasso" or "size of s" not in text else "mismatch"):
        call(oracle)


def test_options_must_be_a_struct_everywhere():
    calls = [lambda: lasso(D, np.ones(4), 0.1, None), lambda: huberfit(D, np.ones(4), 3), lambda: lad(D, np.ones(4), "x"),
             lambda: linearsvm(D, np.ones(4), 0.5, None), lambda: linearsvm_onevsall(D, np.ones((4, 2)), 0.5, None),
             lambda: totalvariation(np.ones(4), 1.0, None), lambda: basispursuit(D.T, np.ones(3), None),
             lambda: quadraticprogram(np.eye(3), np.ones(3), 0.0, np.zeros(3), np.ones(3), None),
             lambda: unwrappedadmm(None, D, None)]
    for c in calls:
        with pytest.raises(MatlabError, match="not a struct"):
            c()


def test_errorcheck_slices_texts():
    with pytest.raises(MatlabError, match="not a numeric vector or integer"):
        errorcheck("ab", "slices", "s", {"workers": 2, "slicelength": 10})
    with pytest.raises(MatlabError, match="does not match length of x"):
        errorcheck([3, 3], "slices", "s", {"workers": 2, "slicelength": 10})
    with pytest.raises(MatlabError, match="Did not provide"):
        errorcheck(0, "slices", "s", {"workers": 2})
    assert errorcheck(0, "slices", "s", {"workers": 3, "slicelength": 10}) == [4, 3, 3]        # errorcheck.m:249-259
