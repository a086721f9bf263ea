"""Host-side mirror of ``results = admm(xminf, zming, options)`` (admm.m:24).

Option parsing, defaults and quirks follow admm.m:46-76 / setopt :780-971; the loop itself
(admm.m:496-767) runs on the device through admm_b200_solve.  ``options`` / ``results`` are dicts
with the reference's field names."""
from __future__ import annotations

import time

import numpy as np

from . import _lib as L
from .errorcheck import MatlabError
from .getproxops import EngineProx
from .parallel import gather_rows


def setopt(options, opttext, default):
    """admm.m:780-971, including the 'Hnormtol' case that reads the field Hreltol (:927-928)."""
    if opttext in options:
        if opttext == "Hnormtol":
            if "Hreltol" not in options:
                raise MatlabError("Reference to non-existent field 'Hreltol'.")
            return options["Hreltol"]
        return options[opttext]
    return default


_STOP = {"standard": L.STOP_STANDARD, "hnorm": L.STOP_HNORM, "both": L.STOP_BOTH}


def admm(xminf, zming, options):
    if not isinstance(options, dict):                                       # admm.m:46-49
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    if not (isinstance(xminf, EngineProx) and isinstance(zming, EngineProx)):
        raise L.EngineError(L.ERR_UNSUPPORTED, "admm: xminf/zming must be the device-resident operators returned "
                            "by getproxops(); arbitrary host function handles would need a CPU path, and the "
                            "engine has none")
    if xminf.engine is not zming.engine or xminf.problem != zming.problem:
        raise L.EngineError(L.ERR_INVALID, "admm: xminf and zming belong to different problems/engines")
    eng = xminf.engine
    gen = getattr(eng, "generation", None)
    if gen is not None and (xminf.generation != gen or zming.generation != gen):
        raise L.EngineError(L.ERR_STATE, "admm: these proximal operators were made for an earlier setup of this engine "
                            "(an Engine holds one problem; call getproxops / the solver again)")

    # admm.m:51-76
    if setopt(options, "adaptive", 0):
        raise L.EngineError(L.ERR_UNSUPPORTED, "options.adaptive is an unfinished experiment in the reference "
                            "(admm.m:724-741) and is not built")
    # admm.m:556-558, 612-616: host handles evaluated INSIDE every iteration -- the device loop cannot call back
    for name in ("altu", "specialnorms"):
        if callable(options.get(name)):
            raise L.EngineError(L.ERR_UNSUPPORTED, "options.%s is a host function handle evaluated in every iteration "
                                "(admm.m:%s); the device-resident loop has no CPU path for it" %
                                (name, "556-558" if name == "altu" else "612-616"))
    quiet = setopt(options, "quiet", 1)
    o = eng.default_options()
    o.rho = float(setopt(options, "rho", 1.0))
    N = setopt(options, "maxiters", 1000)
    o.maxiters = int(np.ceil(N)) if N > 0 else 1000                         # admm.m:334-339
    o.domaxiters = int(bool(setopt(options, "domaxiters", 0)))
    o.relax = float(setopt(options, "relax", 1))
    if setopt(options, "parallel", "none") in ("xminf", "zming", "both") and xminf.problem == "lasso":
        raise L.EngineError(L.ERR_UNSUPPORTED, "parallel consensus LASSO is out of scope (SURVEY.md section 2)")
    obj = setopt(options, "obj", 0)
    o.objevals = int(bool(setopt(options, "objevals", 0)) and (callable(obj) or obj == "engine"))
    o.convtest = int(bool(setopt(options, "convtest", 0)))
    o.convtol = float(setopt(options, "convtol", 1e-10))
    stopcond = setopt(options, "stopcond", "standard")
    # strcmp semantics: an unknown string matches neither test, so the loop runs to maxiters
    o.stopcond = _STOP.get(stopcond, L.STOP_STANDARD)
    if stopcond not in _STOP:
        o.domaxiters = 1
    o.nodualerror = int(bool(setopt(options, "nodualerror", 0)))
    o.abstol = float(setopt(options, "abstol", 1e-5))
    o.reltol = float(setopt(options, "reltol", 1e-3))
    o.hnormtol = float(setopt(options, "Hnormtol", 1e-6))
    o.history = int(bool(setopt(options, "history", 1)))                    # extension
    o.xsolve = int(setopt(options, "xsolve", L.XSOLVE_INVFACTOR))
    o.check_every = int(setopt(options, "check_every", 8))
    o.graph = int(bool(setopt(options, "graph", 1)))                        # extension: CUDA-graph bursts
    o.fast = int(bool(setopt(options, "fast", 0)))                          # admm.m:59-60, 267-298
    o.fasttype = int(setopt(options, "fasttype", "weak") == "weak")
    alg = (2 if o.fasttype else 1) if o.fast else 0
    if alg == 2:
        nrestart = setopt(options, "restart", 0.999)
        o.restart = float(nrestart) if 0 < nrestart < 1 else 0.999
        o.dvaltol = float(setopt(options, "dvaltol", 1e-8))
    for key in ("A", "B", "c"):                                             # admm.m:79-245
        if key not in options and not (key == "c" and options.get("m", 0) > 0):
            what = "vector c" if key == "c" else "matrix " + key
            raise MatlabError("Must specify a %s in constraint Ax + Bz = c!" % what)

    x0, z0, u0 = options.get("x0"), options.get("z0"), options.get("u0")    # admm.m:252-254
    nA, nB, m = eng.dims()
    sharded = eng.nranks > 1 and eng.row_range is not None
    lo, hi = eng.row_range if sharded else (0, None)
    mt = eng.m_total if sharded else m

    def rows(v, full):          # a full-length row vector is cut to this rank's rows
        if v is None or not sharded:
            return v
        v = L.fvec(v)
        return v[lo:hi] if v.size == full else v
    eng.set_init(x0, rows(z0, mt), rows(u0, mt))
    results = {"x0": np.zeros(nA) if x0 is None else L.fvec(x0).copy(),
               "z0": np.zeros(mt if sharded else nB) if z0 is None else L.fvec(z0).copy(),
               "u0": np.zeros(mt) if u0 is None else L.fvec(u0).copy()}
    use_hnorm = bool(o.convtest) or stopcond in ("hnorm", "both")
    if use_hnorm:
        results["Hnormtol"] = o.hnormtol

    start = time.perf_counter()
    if callable(options.get("preprocess")):     # admm.m:473-476: one host call before the loop (after the timer started, :319)
        options["preprocess"]()
    try:
        r = eng.solve(o, want_history=bool(o.history))
    finally:
        if getattr(eng, "_owned", False):       # made by the solver for this one call (engine.acquire_engine)
            eng.close()
    if sharded:                                     # results carry full-length z / u like the reference's
        for key in ("zopt", "uopt", "zvals", "uvals"):
            if key in r:
                r[key] = gather_rows(r[key], mt)
    if alg == 2:                            # the accelerated variant records d, not residual norms
        results["dvaltol"] = o.dvaltol
        results["pnorm"], results["dnorm"] = np.zeros(0), np.zeros(0)
        results["dvals"], results["restarted"] = r["dvals"], r["restarted"]
    else:
        results["pnorm"], results["dnorm"] = r["pnorm"], r["dnorm"]
        results["perr"], results["derr"] = r["perr"], r["derr"]
    if alg:
        results["avals"] = r["avals"]
    if use_hnorm:
        results["Hnormsq"] = r["hnormsq"]
    if o.objevals:
        results["objevals"] = r["objevals"]
    if o.history:
        results["xvals"], results["zvals"], results["uvals"] = r["xvals"], r["zvals"], r["uvals"]
        if use_hnorm:                                                       # w = [x; z; rho*u], admm.m:679
            results["wvals"] = np.vstack([r["xvals"], r["zvals"], o.rho * r["uvals"]])
    results["engine"] = dict(status=r["status"], setup_ms=r["setup_ms"], loop_ms=r["loop_ms"])
    if r["status"] == L.DIVERGED_RETURN:
        # admm.m:692-700: message + early `return` -- steps/xopt/zopt/uopt/runtime/options stay unset
        i = r["steps"]
        H2, H1 = r["hnormsq"][i - 1], r["hnormsq"][i - 2]
        print("Iteration %i: H norms not converging to given relative tolerance: %g is not less or equal to "
              "tol. %g" % (i, (H2 - H1) / (H1 + np.finfo(float).eps), o.convtol))
        print("ADMM seems to not be converging! Please check that your proximal operators are correct!")
        return results
    results["steps"] = r["steps"]                                           # admm.m:746-767
    results["xopt"], results["zopt"], results["uopt"] = r["xopt"], r["zopt"], r["uopt"]
    if o.objevals:
        results["objopt"] = r["objopt"]
    results["runtime"] = time.perf_counter() - start
    if not quiet:
        for i in range(r["steps"]):                                         # admm.m:661-673, printed after the run
            if o.objevals:
                print("%3d\t%10.4f\t%10.4f\t%10.4f\t%10.4f\t%10.2f" % (i + 1, r["pnorm"][i], r["perr"][i],
                                                                     r["dnorm"][i], r["derr"][i], r["objevals"][i]))
            else:
                print("%3d\t%10.4f\t%10.4f\t%10.4f\t%10.4f" % (i + 1, r["pnorm"][i], r["perr"][i], r["dnorm"][i],
                                                             r["derr"][i]))
        print("Elapsed time is %g seconds." % results["runtime"], end="")
        print("Number of steps to convergence: %d" % results["steps"], end="")
    results["options"] = options
    return results
