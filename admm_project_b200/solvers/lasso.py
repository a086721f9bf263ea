"""results = lasso(D, s, lambda, options) -- mirror of solvers/lasso.m:77-245 (serial path)."""
from __future__ import annotations

import time

import numpy as np

from .. import _lib as L
from ..admm import admm
from ..engine import DeviceMatrix, Engine, RowShard, acquire_engine
from ..errorcheck import MatlabError
from ..getproxops import getproxops
from ..parallel import attach_comm, row_range, dist_info


def lasso(D, s, lam, options, engine=None):
    t0 = time.perf_counter()
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    if not isinstance(D, (DeviceMatrix, RowShard)):
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
    eng = acquire_engine(engine, options)
    # lasso.m:159-176: Dts, chol(D'D + rho I) or chol(DD'/rho + I) -- on the device.  Under torch.distributed
    # (one process per GPU) the tall problem is set up from ROW SHARDS: every rank forms D_g'D_g and D_g's_g on
    # its rows (errorcheck.m:249-259 partition), one allreduce sums them (the transpose reduction of
    # unwrappedadmm.m:114-122) and every rank factors the same n x n matrix; the n-sized iterations run replicated.
    xs = int(options.get("xsolve", L.XSOLVE_INVFACTOR))
    rank, world = attach_comm(eng)
    if isinstance(D, (DeviceMatrix, RowShard)) and getattr(D, "m_total", None) and world > 1:
        eng.setup_lasso_sharded(D, s, rho, int(D.m_total), xs)      # D / s hold THIS rank's rows
        m = int(D.m_total)
    elif world > 1 and m >= n and not isinstance(D, DeviceMatrix):
        lo, hi = row_range(m, rank, world)
        eng.setup_lasso_sharded(D[lo:hi, :], s[lo:hi], rho, m, xs)
    else:
        eng.setup_lasso(D, s, rho, xs)
    args = {"engine": eng, "m": m, "n": n, "parallel": 0, "rho": rho, "lambda": float(lam)}   # lasso.m:181-189
    minx, minz, _ = getproxops("LASSO", args)                               # lasso.m:192
    options["obj"] = "engine"      # 1/2*norm(D*x - s)^2 + lambda*norm(z,1) (lasso.m:227), evaluated on the device
    options.update(A=1, At=1, m=n, nA=n, nB=n, B=-1, c=0, parallel="none")  # lasso.m:231-239
    results = admm(minx, minz, options)                                     # lasso.m:242
    results["solverruntime"] = time.perf_counter() - t0                     # lasso.m:243
    return results


def lasso_path(D, s, lambdas, options, engine=None):
    """Engine extension (BASELINE.json configs[1]: "a batch of 64 lambda values as multi-RHS TRSM"): the lasso
    problems ``lasso(D, s, lambdas[j], options)`` for all j on ONE cached factor, the x-update of all columns as two
    triangular FP64 DMMA GEMMs.  Under torch.distributed the lambda COLUMNS are split over the ranks
    (errorcheck.m:249-259 balancing; no communication) after the row-sharded setup, and the
    per-column results are gathered.  Returns a dict of arrays: steps (K), status (K), xopt / zopt / uopt (n x K),
    pnorm / dnorm / perr / derr (maxiters x K, NaN-padded)."""
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)
    lambdas = np.asarray(lambdas, dtype=np.float64).reshape(-1)
    if lambdas.size == 0 or np.any(lambdas < 0):
        raise MatlabError("Argument lambda is not a nonnegative real number!")
    rho = float(options.get("rho", 1.0))
    if not rho > 0:
        raise MatlabError("Argument options.rho is not a positive real number!")
    eng = acquire_engine(engine, options)
    try:
        rank, world = attach_comm(eng)
        if isinstance(D, DeviceMatrix):
            m, n = int(getattr(D, "m_total", D.shape[0])), D.shape[1]
        else:
            D = np.asarray(D, dtype=np.float64)
            m, n = D.shape
        if m < n:
            raise L.EngineError(L.ERR_UNSUPPORTED, "lasso_path: only the tall (rows >= columns) lasso is built for a batch")
        if world > 1 and isinstance(D, DeviceMatrix) and getattr(D, "m_total", None):
            eng.setup_lasso_sharded(D, s, rho, m)
        elif world > 1 and not isinstance(D, DeviceMatrix):
            lo, hi = row_range(m, rank, world)
            eng.setup_lasso_sharded(D[lo:hi, :], np.asarray(s, dtype=np.float64).reshape(-1)[lo:hi], rho, m)
        else:
            eng.setup_lasso(D, s, rho)
        o = eng.default_options()
        o.rho, o.relax = rho, float(options.get("relax", 1))
        o.abstol, o.reltol = float(options.get("abstol", 1e-5)), float(options.get("reltol", 1e-3))
        N = options.get("maxiters", 1000)
        o.maxiters = int(np.ceil(N)) if N > 0 else 1000
        o.domaxiters = int(bool(options.get("domaxiters", 0)))
        o.check_every = int(options.get("check_every", 8))
        clo, chi = row_range(lambdas.size, rank, world)             # this rank's lambda columns
        mine = eng.solve_lasso_batch(o, lambdas[clo:chi]) if chi > clo else None
        if world == 1:
            return mine
        import torch.distributed as dist
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        parts = [q for q in parts if q is not None]
        out = {}
        for key in ("steps", "status"):
            out[key] = np.concatenate([q[key] for q in parts])
        for key in ("xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr"):
            out[key] = np.concatenate([q[key] for q in parts], axis=1)
        out["loop_ms"] = max(q["loop_ms"] for q in parts)
        return out
    finally:
        if getattr(eng, "_owned", False):
            eng.close()
