"""results = totalvariation(s, lambda, options) -- mirror of solvers/totalvariation.m:62-213."""
from __future__ import annotations

import time

import numpy as np

from ..admm import admm
from ..engine import Engine, acquire_engine
from ..errorcheck import MatlabError
from ..getproxops import getproxops


def totalvariation(s, lam, options, engine=None):
    t0 = time.perf_counter()
    s = np.asarray(s, dtype=np.float64)
    if s.ndim > 2 or (s.ndim == 2 and 1 not in s.shape):                    # totalvariation.m:186-188
        raise MatlabError("Argument s is not a vector!")
    s = s.reshape(-1)
    if not (np.isscalar(lam) and np.isreal(lam) and lam >= 0):              # :190-194
        raise MatlabError("Given lambda parameter is not a nonnegative number!")
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    n = s.shape[0]
    if n == 1:      # D = spdiags(.., 1, 1) is a scalar to admm.m:113-161 and totalvariation.m passes no options.nA
        raise MatlabError("Given scalar as matrix A with no number of columnsnA specified in options struct; cannot "
                          "infer nA - please specify nA in options!")
    eng = acquire_engine(engine, options)
    # :127-131 build D = spdiags([1 -1],0:1,n,n), Dt, DtD, Id; the engine keeps them implicit
    xmin, zmin, _ = getproxops("TotalVariation", {"engine": eng, "s": s, "lambda": float(lam)})   # :148
    options.update(A="D", At="D'", B=-1, mB=n, nB=n, c=0, m=n)              # :151-161
    options["obj"] = "engine"      # 1/2*norm(x-s)^2 + lambda*sum(abs(diff(x))), :133-134
    results = admm(xmin, zmin, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results
