"""results = model(P, Q, r, s, options) -- mirror of solvers/model.m:99-146 (error checks :158-218):
minimise 1/2*||P*x - r||^2 + 1/2*||Q*x - s||^2 as f(x) + g(z) subject to x - z = 0.  Used by
testers/modeltest.m and examples/{convergencechecking,fasteradmmcomparison,hnormdemo}.m."""
from __future__ import annotations

import time

import numpy as np

from ..admm import admm
from ..engine import DeviceMatrix, Engine, acquire_engine
from ..errorcheck import MatlabError
from ..getproxops import getproxops


def model(P, Q, r, s, options, engine=None):
    t0 = time.perf_counter()
    if not isinstance(P, DeviceMatrix):
        P = np.asarray(P, dtype=np.float64)
        if P.ndim != 2:
            raise MatlabError("Argument P is not a matrix!")
    if not isinstance(Q, DeviceMatrix):
        Q = np.asarray(Q, dtype=np.float64)
        if Q.ndim != 2:
            raise MatlabError("Argument Q is not a matrix!")
    r, s = np.asarray(r, dtype=np.float64), np.asarray(s, dtype=np.float64)
    isvec = lambda a: a.ndim <= 1 or (a.ndim == 2 and 1 in a.shape)
    if not isvec(r):
        raise MatlabError("Argument r is not a vector!")
    if not isvec(s):
        raise MatlabError("Argument s is not a vector!")
    r, s = r.reshape(-1), s.reshape(-1)                                     # isrow -> transpose, model.m:178-192
    if P.shape[0] != Q.shape[0]:                                            # model.m:196-203
        raise MatlabError("Number of rows in P do not match number of rows in Q!")
    if P.shape[1] != Q.shape[1]:
        raise MatlabError("Number of columns in P do not match number of columns in Q!")
    if P.shape[0] != r.shape[0]:
        raise MatlabError("Number of rows in P does not match length of vector r!")
    if Q.shape[0] != s.shape[0]:
        raise MatlabError("Number of rows in Q does not match length of vector s!")
    if not isinstance(options, dict):
        raise MatlabError("Given options argument is not a struct! Please check your arguments and try again.")
    options = dict(options)
    n = P.shape[1]
    eng = acquire_engine(engine, options)
    rho = float(options["rho"]) if "rho" in options else 1.0
    args = {"engine": eng, "P": P, "Q": Q, "r": r, "s": s, "n": n, "rho": rho}      # model.m:123-128
    minx, minz, _ = getproxops("Model", args)
    options.update(A=1, B=-1, c=0, m=n, nA=n, nB=n)                         # model.m:133-138
    options["obj"] = "engine"                                               # 1/2||Px-r||^2 + 1/2||Qz-s||^2, :139-140
    results = admm(minx, minz, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results
