"""results = huberfit(D, s, options) / lad(D, s, options) -- mirrors of solvers/huberfit.m:83-186
and solvers/lad.m:51-154."""
from __future__ import annotations

import time

import numpy as np

from ..admm import admm
from ..engine import DeviceMatrix, Engine, acquire_engine
from ..errorcheck import MatlabError
from ..getproxops import getproxops


def _robustfit(problem, D, s, options, engine):
    t0 = time.perf_counter()
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    if not isinstance(D, DeviceMatrix):
        D = np.asarray(D, dtype=np.float64)
        s = np.asarray(s, dtype=np.float64).reshape(-1)
        if D.ndim != 2:
            raise MatlabError("Argument D is not a matrix!")
        if D.shape[0] != s.shape[0]:
            raise MatlabError("The number of rows in argument D do not match size of s!")
        m, n = D.shape
    else:
        m, n = int(getattr(D, "m_total", D.shape[0])), D.shape[1]
    eng = acquire_engine(engine, options)
    args = {"engine": eng, "D": D, "s": s}
    if "relax" in options and options["relax"] != 1:                        # huberfit.m:156-158, lad.m:124-126
        args["userelax"] = 1
    minx, minz, _ = getproxops(problem, args)       # R = chol(D'*D,'lower') on the device (huberfit.m:166, lad.m:134)
    options.update(A="D", B=-1, c="s", m=m, nA=n, nB=m)                     # huberfit.m:172-177, lad.m:140-145
    options["obj"] = "engine"     # 1/2*sum(huber(z)) (huberfit.m:180) / norm(z,1) (lad.m:148), on the device
    results = admm(minx, minz, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results


def huberfit(D, s, options, engine=None):
    return _robustfit("huberfit", D, s, options, engine)


def lad(D, s, options, engine=None):
    return _robustfit("lad", D, s, options, engine)
