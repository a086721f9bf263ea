"""Host-side mirror of ``[minx, minz, extra] = getproxops(problem, args)`` (getProxOps.m:13).

The reference returns MATLAB closures; the engine returns ``EngineProx`` descriptors that name the
device kernel family standing for that closure.  ``admm`` recognises a matching (minx, minz) pair
and runs the WHOLE loop on the GPU -- a per-iteration host round trip would destroy the roofline
(SURVEY.md section 8b).  Calling a descriptor on the host raises: there is no CPU fallback."""
from __future__ import annotations

import numpy as np

from ._lib import EngineError, ERR_UNSUPPORTED, SVM_HINGE, SVM_01, HUBERFIT, LAD, PROX_BOX
from .engine import DeviceMatrix
from .errorcheck import MatlabError
from .parallel import attach_comm, row_range


class EngineProx:
    """Stands for one function handle of getProxOps.m (e.g. @xminLASSO, getProxOps.m:1192)."""

    def __init__(self, role, problem, name, engine, params):
        self.role, self.problem, self.name, self.engine, self.params = role, problem, name, engine, params
        # An Engine holds ONE problem: the pair is valid only for the setup it was made for.  admm() compares this
        # stamp with the engine's current one and refuses a stale pair (a later setup on the same engine).
        self.generation = engine.generation if engine is not None and hasattr(engine, "generation") else None

    def __call__(self, *a, **k):
        raise EngineError(ERR_UNSUPPORTED, "%s is a device-resident proximal operator; it is evaluated inside "
                          "admm() on the GPU and has no host (CPU) evaluation" % self.name)

    def __repr__(self):
        return "<EngineProx %s/%s>" % (self.problem, self.name)


def _need_engine(eng, problem):
    if eng is None:
        raise EngineError(ERR_UNSUPPORTED, "getproxops('%s'): args.engine is missing -- the operators are "
                          "device-resident (there is no CPU path)" % problem)
    return eng


def _setup_rows(eng, kind, D, aux, C):
    """One-time device setup of an A = D problem.  Under torch.distributed (world > 1) every rank
    keeps only its row block of D / aux (errorcheck.m:249-259) and W = sum_g D_g'D_g is allreduced
    (unwrappedadmm.m:114-122).  A DeviceMatrix is taken as THIS rank's rows already."""
    rank, world = attach_comm(eng)
    if isinstance(D, DeviceMatrix):
        m_total = int(getattr(D, "m_total", D.shape[0]))
        lo, hi = getattr(D, "row_range", (0, D.shape[0]))
        eng.setup_unwrapped(kind, D, aux, C, m_total=m_total)
    else:
        D = np.asarray(D, dtype=np.float64)
        m_total = D.shape[0]
        lo, hi = row_range(m_total, rank, world)
        aux = np.asarray(aux, dtype=np.float64).reshape(-1)
        eng.setup_unwrapped(kind, D[lo:hi, :], aux[lo:hi], C, m_total=m_total)
    eng.row_range = (lo, hi)


_OUT = ("linearprogram", "quadraticprogram", "covarianceselection")


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
    elif problem == "basispursuit":                                         # getProxOps.m:126-142
        eng = _need_engine(eng, problem)
        if "D" in args:                    # the engine keeps chol(D*D') instead of the dense projector P, q
            eng.setup_basispursuit(args["D"], args["s"])
        minx = EngineProx("xminf", "basispursuit", "xminBasisPursuit", eng, {})
        minz = EngineProx("zming", "basispursuit", "zminSoftThresholding", eng, {})
    elif problem == "totalvariation":                                       # getProxOps.m:172-199
        eng = _need_engine(eng, problem)
        eng.setup_totalvariation(args["s"], args["lambda"])      # D, Dt, DtD stay implicit on the device
        minx = EngineProx("xminf", "totalvariation", "xminTotalVariation", eng, {})
        minz = EngineProx("zming", "totalvariation", "zminSoftThresholding", eng, {"lambda": args["lambda"]})
    elif problem == "linearsvm":                                            # getProxOps.m:256-309
        D, ell, C, loss = args["D"], args["ell"], args["C"], args["lossfunction"]
        eng = _need_engine(eng, problem)
        kind = SVM_01 if loss == "01" else SVM_HINGE      # strcmp(loss,'01') -- '0-1' is hinge (:1094)
        _setup_rows(eng, kind, D, ell, C)
        minx = EngineProx("xminf", "linearsvm", "xminLinearSVM", eng, {})  # Dplus*(z-u) == W\D'(z-u)
        name = "zminParallelLinearSVM" if "slices" in args else "zminLinearSVM"
        minz = EngineProx("zming", "linearsvm", name, eng, {"C": C, "lossfunction": loss})
    elif problem in ("lad", "huberfit"):                                    # getProxOps.m:780-912
        D, s_ = args["D"], args["s"]
        eng = _need_engine(eng, problem)
        _setup_rows(eng, LAD if problem == "lad" else HUBERFIT, D, s_, 0.0)
        minx = EngineProx("xminf", problem, "xminLAD", eng, {})
        minz = EngineProx("zming", problem, "zminSoftThresholding" if problem == "lad" else
                          "zminHuberSoftThresholding", eng, {"userelax": int(bool(args.get("userelax", 0)))})
    elif problem == "quadraticprogram" and args.get("constraint") == "bounded":     # getProxOps.m:1441-1474
        eng = _need_engine(eng, problem)
        eng.setup_quadratic(PROX_BOX, args["P"], args["q"], args.get("r", 0.0), args["rho"], args["lb"], args["ub"])
        minx = EngineProx("xminf", "quadraticprogram", "xminQuadraticProgramBounded", eng, {})
        minz = EngineProx("zming", "quadraticprogram", "zminQuadraticProgramBounded", eng, {})
    elif problem == "model":                                                # getProxOps.m:55-110
        eng = _need_engine(eng, problem)
        # The reference hands over PtP, Ptr, QtQ, Qts (model.m:123-128); the engine forms them on the device
        # from P, Q, r, s (DMMA Gram), keeps both Gram matrices and caches chol(. + rho*I) -- the reference
        # re-adds rho and re-solves the dense systems in every iteration (:967-973, :1004-1011).
        for key in ("P", "Q", "r", "s"):
            if key not in args:
                raise EngineError(ERR_UNSUPPORTED, "getproxops('model'): pass args.P, args.Q, args.r, args.s -- the "
                                  "Gram matrices are formed on the device")
        eng.setup_model(args["P"], args["Q"], args["r"], args["s"], args.get("rho", 1.0))
        minx = EngineProx("xminf", "model", "xminModel", eng, {})
        minz = EngineProx("zming", "model", "zminModel", eng, {})
    elif problem in _OUT:
        raise EngineError(ERR_UNSUPPORTED, "problem '%s' is outside the engine's hot path (SURVEY.md section 2)" % problem)
    else:
        raise MatlabError("Invalid input for problem - given string is not a solver!")
    minx.generation = minz.generation = getattr(eng, "generation", None)      # stamp AFTER the setup this call ran
    return minx, minz, extra
