"""results = linearsvm(D, ell, C, options) -- mirror of solvers/linearsvm.m:92-308."""
from __future__ import annotations

import time

import numpy as np

from ..engine import DeviceMatrix, Engine
from ..errorcheck import MatlabError, errorcheck
from ..getproxops import getproxops
from .unwrappedadmm import unwrappedadmm


def linearsvm(D, ell, C, options, engine=None):
    t0 = time.perf_counter()
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    if not (np.isscalar(C) and np.isreal(C) and C >= 0):                    # linearsvm.m:270-274
        raise MatlabError("Given regularization parameter C is not a nonnegative number!")
    if not isinstance(D, DeviceMatrix):
        D = np.asarray(D, dtype=np.float64)
        ell = np.asarray(ell, dtype=np.float64).reshape(-1)
        if D.ndim != 2 or D.shape[0] != ell.shape[0]:                       # :298-300
            raise MatlabError("Product ell*D is not possible; sizes incompatible!")
        m = D.shape[0]
    else:
        m = int(getattr(D, "m_total", D.shape[0]))
    loss = options.get("lossfunction", "hinge")                             # :154-158
    eng = engine or options.get("engine") or Engine(int(options.get("device", 0)))
    args = {"engine": eng, "D": D, "ell": ell, "C": float(C), "lossfunction": loss}   # :210-214
    if options.get("parallel") in ("both", "zming", "xminf"):               # :170-205
        options["parallel"] = "both"
        workers = max(int(options.get("workers", eng.nranks)), 1)
        args["slices"] = errorcheck(options.get("slices", 0), "slices", "options.slices",
                                    {"slicelength": m, "workers": workers})
    _, minz, _ = getproxops("LinearSVM", args)                              # :217
    options["obj"] = "engine"   # 1/2*x'x + C*sum(pos(1 - ell.*(D*x))) (:233) or the sign variant (:236), on the device
    results = unwrappedadmm(minz, D, options)                               # :242
    results["solverruntime"] = time.perf_counter() - t0
    return results
