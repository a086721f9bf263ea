"""Host-side mirror of ``[minx, minz, extra] = getproxops(problem, args)`` (getProxOps.m:13).

The reference returns MATLAB closures; the engine returns ``EngineProx`` descriptors that name the
device kernel family standing for that closure.  ``admm`` recognises a matching (minx, minz) pair
and runs the WHOLE loop on the GPU -- a per-iteration host round trip would destroy the roofline
(SURVEY.md section 8b).  Calling a descriptor on the host raises: there is no CPU fallback."""
from __future__ import annotations

from ._lib import EngineError, ERR_UNSUPPORTED
from .errorcheck import MatlabError


class EngineProx:
    """Stands for one function handle of getProxOps.m (e.g. @xminLASSO, getProxOps.m:1192)."""

    def __init__(self, role, problem, name, engine, params):
        self.role, self.problem, self.name, self.engine, self.params = role, problem, name, engine, params

    def __call__(self, *a, **k):
        raise EngineError(ERR_UNSUPPORTED, "%s is a device-resident proximal operator; it is evaluated inside "
                          "admm() on the GPU and has no host (CPU) evaluation" % self.name)

    def __repr__(self):
        return "<EngineProx %s/%s>" % (self.problem, self.name)


_OUT = ("model", "linearprogram", "quadraticprogram", "covarianceselection")


def getproxops(problem, args):
    extra = {}
    if not isinstance(problem, str):
        raise MatlabError("Given problem argument is not a string specifying for which problem "
                          "proximal operators are needed!")
    problem = problem.lower()                                               # getProxOps.m:37
    if not isinstance(args, dict):
        raise MatlabError("Given struct args is not a struct containing arguments needed for "
                          "proximal operators for the given problem!")
    eng = args.get("engine")
    if problem == "lasso":                                                  # getProxOps.m:311-456
        if eng is None:
            raise EngineError(ERR_UNSUPPORTED, "getproxops('lasso'): args.engine is missing -- the factor L/U is "
                              "built on the device by solvers.lasso (admm_b200_setup_lasso)")
        if args.get("parallel"):
            raise EngineError(ERR_UNSUPPORTED, "consensus (parfor) LASSO is out of scope: the reference path "
                              "returns an all-zero minz (getProxOps.m:1275-1276)")
        eng.set_lambda(args["lambda"])
        minx = EngineProx("xminf", "lasso", "xminLASSO", eng, dict(m=args["m"], n=args["n"], rho=args["rho"]))
        minz = EngineProx("zming", "lasso", "zminSoftThresholding", eng, {"lambda": args["lambda"]})
    elif problem in _OUT:
        raise EngineError(ERR_UNSUPPORTED, "problem '%s' is outside the engine's hot path (SURVEY.md section 2)" % problem)
    elif problem in ("basispursuit", "totalvariation", "linearsvm", "lad", "huberfit"):
        raise EngineError(ERR_UNSUPPORTED, "problem '%s' is not built yet in this engine" % problem)
    else:
        raise MatlabError("Invalid input for problem - given string is not a solver!")
    return minx, minz, extra
