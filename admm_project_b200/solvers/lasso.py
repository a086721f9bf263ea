"""results = lasso(D, s, lambda, options) -- mirror of solvers/lasso.m:77-245 (serial path)."""
from __future__ import annotations

import time

import numpy as np

from .. import _lib as L
from ..admm import admm
from ..engine import DeviceMatrix, Engine
from ..errorcheck import MatlabError
from ..getproxops import getproxops


def lasso(D, s, lam, options, engine=None):
    t0 = time.perf_counter()
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    if not isinstance(D, DeviceMatrix):
        D = np.asarray(D, dtype=np.float64)
        if D.ndim != 2:                                                     # lasso.m:132-136 (errorcheck ismatrix)
            raise MatlabError("Argument D is not a matrix!")
    m, n = D.shape
    if not isinstance(s, (int, np.integer)):
        s = np.asarray(s, dtype=np.float64).reshape(-1)
        if s.shape[0] != m:
            raise MatlabError("The number of rows in argument D do not match size of s!")
    if not (np.isscalar(lam) and np.isreal(lam) and lam >= 0):
        raise MatlabError("Argument lambda is not a nonnegative real number!")
    rho = float(options["rho"]) if "rho" in options else 1.0                # lasso.m:137-141
    if not rho > 0:
        raise MatlabError("Argument options.rho is not a positive real number!")
    if options.get("parallel") in ("both", "zming", "xminf"):               # lasso.m:144-156, 193-224
        raise L.EngineError(L.ERR_UNSUPPORTED, "lasso: the parfor consensus branch (lasso.m:193-224) is out of scope; "
                            "its reference implementation returns an all-zero z (getProxOps.m:1275-1276)")
    eng = engine or options.get("engine") or Engine(int(options.get("device", 0)))
    # lasso.m:159-176: Dts, chol(D'D + rho I) or chol(DD'/rho + I) -- on the device
    eng.setup_lasso(D, s, rho, int(options.get("xsolve", L.XSOLVE_INVFACTOR)))
    args = {"engine": eng, "m": m, "n": n, "parallel": 0, "rho": rho, "lambda": float(lam)}   # lasso.m:181-189
    minx, minz, _ = getproxops("LASSO", args)                               # lasso.m:192
    options["obj"] = "engine"      # 1/2*norm(D*x - s)^2 + lambda*norm(z,1) (lasso.m:227), evaluated on the device
    options.update(A=1, At=1, m=n, nA=n, nB=n, B=-1, c=0, parallel="none")  # lasso.m:231-239
    results = admm(minx, minz, options)                                     # lasso.m:242
    results["solverruntime"] = time.perf_counter() - t0                     # lasso.m:243
    return results
