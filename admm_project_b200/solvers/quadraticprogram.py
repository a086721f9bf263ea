"""results = quadraticprogram(P, q, r, cons1, cons2, options) -- mirror of the 'bounded' branch of
solvers/quadraticprogram.m:99-257 (lb <= x <= ub).  The 'standard' branch needs a dense indefinite
KKT solve per iteration (getProxOps.m:1410) and is outside the engine's hot path."""
from __future__ import annotations

import time

import numpy as np

from .. import _lib as L
from ..admm import admm
from ..engine import Engine, acquire_engine
from ..errorcheck import MatlabError
from ..getproxops import getproxops


def quadraticprogram(P, q, r, cons1, cons2, options, engine=None):
    t0 = time.perf_counter()
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    P = np.asarray(P, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64).reshape(-1)
    nP = P.shape[0]
    if q.shape[0] != nP:                                                    # quadraticprogram.m:280-284
        raise MatlabError("The dimensions of square matrix P and vector q do not match!")
    c1, c2 = np.asarray(cons1, dtype=np.float64), np.asarray(cons2, dtype=np.float64)
    isvec = lambda a: a.ndim <= 1 or 1 in a.shape
    if not (isvec(c1) and isvec(c2)):
        raise L.EngineError(L.ERR_UNSUPPORTED, "quadraticprogram: the 'standard' constraint form (dense KKT solve "
                            "every iteration, getProxOps.m:1410) is outside the engine's hot path")
    c1, c2 = c1.reshape(-1), c2.reshape(-1)
    if c1.shape[0] != c2.shape[0]:                                          # :305-311
        raise MatlabError("Lengths of lower and upper bound constraints on solution x do not match!")
    if c1.shape[0] != nP:
        raise MatlabError("Bound vectors do not match predicted length of solution x!")
    if np.array_equal(np.maximum(c1, c2), c1):                              # :312-319
        c1, c2 = c2, c1
    elif not np.array_equal(np.maximum(c1, c2), c2):
        raise MatlabError("Given constraint variables do not specify an upper and lower bound on solution x!")
    n = nP
    rho = float(options["rho"]) if "rho" in options else 1.0                # :171-175
    eng = acquire_engine(engine, options)
    args = {"engine": eng, "P": P, "q": q, "r": float(r), "lb": c1, "ub": c2, "rho": rho, "n": n,
            "constraint": "bounded"}                                        # :210-216
    minx, minz, _ = getproxops("quadraticprogram", args)
    options.update(A=1, B=-1, c=0, m=n, nA=n, nB=n)                         # :229-234
    options["obj"] = "engine"                                               # 1/2*x'*P*x + q'*x + r, :237
    results = admm(minx, minz, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results
