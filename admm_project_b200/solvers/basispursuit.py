"""results = basispursuit(D, s, options) -- mirror of solvers/basispursuit.m:52-210."""
from __future__ import annotations

import time

import numpy as np

from ..admm import admm
from ..engine import DeviceMatrix, Engine, acquire_engine
from ..errorcheck import MatlabError
from ..getproxops import getproxops


def basispursuit(D, s, options, engine=None):
    t0 = time.perf_counter()
    if not isinstance(D, DeviceMatrix):
        D = np.asarray(D, dtype=np.float64)
        s = np.asarray(s, dtype=np.float64).reshape(-1)
        if D.ndim != 2:
            raise MatlabError("Argument D is not a matrix!")
        ms = s.shape[0]
    else:
        ms = D.shape[0]
    mD, nD = D.shape
    if mD == nD and mD == ms:                                               # basispursuit.m:192-203
        raise MatlabError("Square matrix problem Dx = s; don't need Basis Pursuit to solve this!")
    elif mD > nD and mD == ms:
        raise MatlabError("Overdetermined system Dx = s, as D has more rows thancolumns; use Unwrapped ADMM "
                          "solver for efficiency, instead.")
    elif mD != ms:
        raise MatlabError("The number of rows in matrix D must match the number of rows in signal vector s!")
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    n = nD
    eng = acquire_engine(engine, options)
    # basispursuit.m:116-120 builds P = I - D'((DD')\D), q = D'((DD')\s); the engine factors D*D' instead
    minx, minz, _ = getproxops("BasisPursuit", {"engine": eng, "D": D, "s": s})    # :127
    options.update(A=1, B=-1, c=0, m=n, nA=n, nB=n, solver="basispursuit")  # :130-137
    options["obj"] = "engine"                                               # norm(x,1), :140
    results = admm(minx, minz, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results
