"""results = linearsvm(D, ell, C, options) -- mirror of solvers/linearsvm.m:92-308."""
from __future__ import annotations

import time

import numpy as np

from ..engine import DeviceMatrix, Engine, acquire_engine
from ..errorcheck import MatlabError, errorcheck
from ..getproxops import getproxops
from .unwrappedadmm import check_single_column, unwrappedadmm
from ..parallel import attach_comm, gather_rows, row_range, shared_draw


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
    check_single_column(m, D.shape[1])      # raised by admm.m at the end of the reference's call chain; same text, before device work
    loss = options.get("lossfunction", "hinge")                             # :154-158
    eng = acquire_engine(engine, options)
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


def linearsvm_onevsall(D, ELL, C, options, engine=None):
    """Engine extension: the one-vs-all loop of examples/mnistsvm.m:121-156 (one ``linearsvm(D, ell_k, C,
    options)`` call per class, hinge loss) run as ONE batch -- D is swept twice per iteration for all classes
    instead of twice per class.  ELL is m x K with +-1 columns.  Returns a list of K result dicts with the
    reference's fields (steps, xopt, zopt, uopt, pnorm, perr, dnorm = derr = NaN, objevals, x0, z0, u0).
    Initial iterates are drawn like unwrappedadmm.m:87-89 does for every call: rand(n), rand(m), rand(m)
    per class, in class order."""
    from .. import _lib as L
    t0 = time.perf_counter()
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    if not (np.isscalar(C) and np.isreal(C) and C >= 0):
        raise MatlabError("Given regularization parameter C is not a nonnegative number!")
    D = np.asarray(D, dtype=np.float64)
    ELL = np.asarray(ELL, dtype=np.float64)
    if ELL.ndim == 1:
        ELL = ELL[:, None]
    if D.ndim != 2 or D.shape[0] != ELL.shape[0]:
        raise MatlabError("Product ell*D is not possible; sizes incompatible!")
    if options.get("lossfunction", "hinge") == "01":
        raise L.EngineError(L.ERR_UNSUPPORTED, "linearsvm_onevsall: the class batch is built for the hinge loss")
    m, n = D.shape
    K = ELL.shape[1]
    check_single_column(m, n)
    eng = acquire_engine(engine, options)
    # Under torch.distributed every rank keeps its row block of D / ELL (errorcheck.m:249-259) and the batch
    # exchanges ONE allreduce of K x [D'r ; scalars] per iteration.  Rank 0 draws the initial iterates for all.
    rank, world = attach_comm(eng)
    lo, hi = row_range(m, rank, world)

    def draw():
        X0, Z0, U0 = np.zeros((n, K), order="F"), np.zeros((m, K), order="F"), np.zeros((m, K), order="F")
        for k in range(K):                   # same draw order as K successive unwrappedadmm calls (:87-89)
            X0[:, k], Z0[:, k], U0[:, k] = np.random.rand(n), np.random.rand(m), np.random.rand(m)
        return X0, Z0, U0
    X0, Z0, U0 = shared_draw(draw)
    eng.setup_unwrapped(L.SVM_HINGE, D[lo:hi, :], ELL[lo:hi, 0], float(C), m_total=m)
    eng.row_range = (lo, hi)
    o = eng.default_options()
    o.rho = float(options.get("rho", 1.0))
    o.abstol, o.reltol = float(options.get("abstol", 1e-5)), float(options.get("reltol", 1e-3))
    o.hnormtol = float(options["Hreltol"]) if "Hnormtol" in options else 1e-6
    o.objevals = int(bool(options.get("objevals", 0)))
    o.maxiters, o.stopcond, o.nodualerror = 1000, L.STOP_BOTH, 1          # unwrappedadmm.m:90-92
    try:
        r = eng.solve_unwrapped_batch(o, ELL[lo:hi, :], X0, Z0[lo:hi, :], U0[lo:hi, :])
    finally:
        if getattr(eng, "_owned", False):
            eng.close()
    if world > 1:                               # results carry full-length z / u like the reference's
        r["zopt"], r["uopt"] = gather_rows(r["zopt"], m), gather_rows(r["uopt"], m)
    out = []
    for k in range(K):
        s = int(r["steps"][k])
        res = dict(steps=s, xopt=r["xopt"][:, k].copy(), zopt=r["zopt"][:, k].copy(), uopt=r["uopt"][:, k].copy(),
                   pnorm=r["pnorm"][:s, k].copy(), perr=r["perr"][:s, k].copy(), dnorm=np.full(s, np.nan),
                   derr=np.full(s, np.nan), x0=X0[:, k], z0=Z0[:, k], u0=U0[:, k],
                   engine=dict(status=int(r["status"][k]), loop_ms=r["loop_ms"]))
        if o.objevals:
            res["objevals"] = r["objevals"][:s, k].copy()
            res["objopt"] = float(res["objevals"][-1]) if s else float("nan")
        out.append(res)
    for res in out:
        res["solverruntime"] = time.perf_counter() - t0
    return out
