"""results = unwrappedadmm(zming, D, options) -- mirror of solvers/unwrappedadmm.m:1-143.

The reference's serial x-update is pinv(D)*(z-u) (:76-78) and its parfor x-update is the transpose
reduction W \\ sum_i D_i'(z_i-u_i), W = sum_i D_i'D_i (:96-141).  The engine always runs the second
form (one cached Cholesky of W instead of a re-factorisation per iteration); with one GPU there is
one slice, with torch.distributed initialised every rank is one slice."""
from __future__ import annotations

import time

import numpy as np

from .. import _lib as L
from ..admm import admm
from ..engine import DeviceMatrix
from ..errorcheck import MatlabError, errorcheck
from ..getproxops import EngineProx
from ..parallel import shared_draw


def check_single_column(m, n):
    """unwrappedadmm.m:81-82 hands A = D, At = D' to admm() WITHOUT options.nA, and admm.m:113-201 cannot infer nA from
    a single column: a one-feature D is a reference error, not a supported shape."""
    if n == 1:
        raise MatlabError("Given scalar as matrix A with no number of columnsnA specified in options struct; cannot "
                          "infer nA - please specify nA in options!" if m == 1 else
                          "Number of rows in At (A transpose) do not match number of columns in A, in constraint "
                          "Ax + Bz = c")


def unwrappedadmm(zming, D, options):
    t0 = time.perf_counter()
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    if not isinstance(zming, EngineProx) or zming.role != "zming":
        raise L.EngineError(L.ERR_UNSUPPORTED, "unwrappedadmm: zming must be a device-resident operator from "
                            "getproxops() (no CPU path for arbitrary function handles)")
    eng = zming.engine
    if isinstance(D, DeviceMatrix):
        m, n = int(getattr(D, "m_total", D.shape[0])), D.shape[1]
    else:
        m, n = np.asarray(D).shape                                          # :43
    check_single_column(m, n)
    if options.get("parallel") in ("xminf", "zming", "both"):               # :45-74
        workers = max(int(options.get("workers", eng.nranks)), 1)
        slices = options.get("slices", 0)
        options["slices"] = errorcheck(np.floor(np.real(np.atleast_1d(slices))), "slices", "options.slices",
                                       {"workers": workers, "slicelength": m})
        options["parallel"] = "none"     # the slices are the engine's row shards; admm() itself stays serial
    xminf = EngineProx("xminf", zming.problem, "proxf", eng, {})            # :125-141
    options["A"] = "D"                                                      # :81-92 (D lives on the device)
    options["At"] = "D'"
    options["B"] = -1
    options["nB"] = m
    options["c"] = 0
    options["m"] = m
    # same draw order as the reference: x0, z0, u0 (:87-89).  Under torch.distributed rank 0 draws and everybody
    # receives the same vectors, so every rank cuts its rows out of ONE global init and results['x0'|'z0'|'u0']
    # describe the init that was actually used.
    options["x0"], options["z0"], options["u0"] = shared_draw(lambda: (np.random.rand(n), np.random.rand(m),
                                                                       np.random.rand(m)))
    options["maxiters"] = 1000
    options["stopcond"] = "both"
    options["nodualerror"] = 1
    results = admm(xminf, zming, options)                                   # :94
    results["solverruntime"] = time.perf_counter() - t0
    return results
