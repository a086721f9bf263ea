"""ORACLE (test infrastructure) -- restatement of the one-time setup done by the in-scope
``/root/reference/solvers/*.m`` before they call ``admm``.  Parity unpinned (oracle/admm.py).

The PCT pool size (``gcp().NumWorkers``) is taken from ``options['workers']`` (default 2).
``rand`` (unwrappedadmm.m:87-89) draws from NumPy's global RandomState in the same order
(x0, z0, u0), so a test that seeds ``np.random.seed`` before calling the oracle and the
product gives both the same start.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from .admm import MatlabError, admm, slice_ranges
from .errorcheck import errorcheck
from .getproxops import getproxops, huber


def _col(v):
    return np.asarray(v, dtype=np.float64).reshape(-1)


def lasso(D, s, lam, options):
    """solvers/lasso.m:77-245 (serial path; the consensus branch :193-224 is out of scope)."""
    t0 = time.perf_counter()
    options = dict(options)
    D = np.asarray(D, dtype=np.float64)
    s = _col(s)
    if not (np.isscalar(lam) and np.real(lam) >= 0):                        # :132
        raise MatlabError("Argument lambda is not a nonnegative real number!")
    rho = float(options["rho"]) if "rho" in options else 1.0                # :137-141
    if rho <= 0:
        raise MatlabError("Argument options.rho is not a positive real number!")
    parallel = 0
    if options.get("parallel") in ("both", "zming", "xminf"):               # :144-156
        raise MatlabError("oracle: consensus (parfor) LASSO is out of scope")
    Dts = D.T @ s                                                           # :160
    m, n = D.shape
    if m >= n:                                                              # :166-173
        L = sla.cholesky(D.T @ D + rho * np.eye(n), lower=True, check_finite=False)
    else:
        L = sla.cholesky(1 / rho * (D @ D.T) + np.eye(m), lower=True, check_finite=False)
    U = L.T                                                                 # :175-176
    args = dict(D=D, Dts=Dts, L=L, U=U, m=m, n=n, parallel=parallel, rho=rho)
    args["lambda"] = lam
    minx, minz, _ = getproxops("LASSO", args)                               # :192
    options["obj"] = lambda x, z: 0.5 * float(np.sum((D @ x - s) ** 2)) + lam * float(np.sum(np.abs(z)))
    options.update(A=1, At=1, m=n, nA=n, nB=n, B=-1, c=0, parallel="none")  # :231-239
    results = admm(minx, minz, options)                                     # :242
    results["solverruntime"] = time.perf_counter() - t0
    return results


def unwrappedadmm(zming, D, options):
    """solvers/unwrappedadmm.m:1-143."""
    t0 = time.perf_counter()
    options = dict(options)
    D = np.asarray(D, dtype=np.float64)
    m, n = D.shape                                                          # :43
    state = {}
    if options.get("parallel") in ("xminf", "zming", "both"):               # :45-74
        workers = int(options.get("workers", 2))
        options["parallel"] = "zming" if options["parallel"] == "both" else "none"
        if "slices" in options:
            options["slices"] = np.floor(np.real(np.atleast_1d(options["slices"])))
        else:
            options["slices"] = 0
        options["slices"] = errorcheck(options["slices"], "slices", "options.slices",
                                       {"workers": workers, "slicelength": m})
        options["workers"] = workers

        def nodepreprocessing():                                            # :96-123
            state["ranges"] = slice_ranges(options["slices"])
            state["Di"] = [D[a:b, :] for a, b in state["ranges"]]
            W = np.zeros((n, n))
            for Di in state["Di"]:
                W = W + Di.T @ Di
            state["W"] = W

        def proxf(_x, z, u, _rho):                                          # :125-141
            d = 0
            for (a, b), Di in zip(state["ranges"], state["Di"]):
                d = d + Di.T @ (z[a:b] - u[a:b])
            return np.linalg.solve(state["W"], d)                           # W \ d
        xminf = proxf
        options["preprocess"] = nodepreprocessing
    else:
        Dplus = np.linalg.pinv(D)                                           # :76
        xminf = lambda _x, z, u, _rho: Dplus @ (z - u)                      # :78
    options["A"] = D                                                        # :81-92
    options["At"] = D.T
    options["B"] = -1
    options["nB"] = m
    options["c"] = 0
    options["m"] = m
    options["x0"] = np.random.rand(n)
    options["z0"] = np.random.rand(m)
    options["u0"] = np.random.rand(m)
    options["maxiters"] = 1000
    options["stopcond"] = "both"
    options["nodualerror"] = 1
    results = admm(xminf, zming, options)                                   # :94
    results["solverruntime"] = time.perf_counter() - t0
    return results


def linearsvm(D, ell, C, options):
    """solvers/linearsvm.m:92-246."""
    t0 = time.perf_counter()
    options = dict(options)
    if not (np.isscalar(C) and np.real(C) >= 0):                            # :270-274
        raise MatlabError("Given regularization parameter C is not a nonnegative number!")
    C = float(np.real(C))
    ell = _col(ell)
    D = np.asarray(D, dtype=np.float64)
    if D.shape[0] != ell.shape[0]:                                          # :298-300
        raise MatlabError("Product ell*D is not possible; sizes incompatible!")
    loss = options.get("lossfunction", "hinge")                             # :154-158
    parallel = 0
    if options.get("parallel") in ("both", "zming", "xminf"):               # :170-181
        options["parallel"] = "both"
        parallel = 1
    args = {}
    if not parallel:
        args["Dplus"] = np.linalg.pinv(D)                                   # :185-186
    else:
        slices = options.get("slices", 0)                                   # :190-205
        workers = int(options.get("workers", 2))
        args["slices"] = errorcheck(slices, "slices", "options.slices",
                                    {"slicelength": D.shape[0], "workers": workers})
    args.update(D=D, Dt=D.T, ell=ell, C=C, lossfunction=loss)               # :210-214
    _, minz, _ = getproxops("LinearSVM", args)                              # :217
    if loss == "hinge":                                                     # :231-237
        options["obj"] = lambda x, z: 0.5 * float(x @ x) + C * float(np.sum(np.maximum(1 - ell * (D @ x), 0)))
    else:
        options["obj"] = lambda x, z: 0.5 * float(x @ x) + C * float(np.sum(np.maximum(np.sign(1 - ell * (D @ x)), 0)))
    results = unwrappedadmm(minz, D, options)                               # :242
    results["solverruntime"] = time.perf_counter() - t0
    return results


def _robustfit(D, s, options, problem):
    t0 = time.perf_counter()
    options = dict(options)
    D = np.asarray(D, dtype=np.float64)
    s = _col(s)
    m, n = D.shape
    if m != s.shape[0]:
        raise MatlabError("The number of rows in argument D do not match size of s!")
    args = {}
    if "relax" in options and options["relax"] != 1:                        # huberfit.m:156-158
        args["userelax"] = 1
    args["D"], args["s"] = D, s
    args["R"] = sla.cholesky(D.T @ D, lower=True, check_finite=False)       # huberfit.m:166 / lad.m:134
    minx, minz, _ = getproxops(problem, args)
    options.update(A=D, B=-1, c=s, m=m, nA=n, nB=m)                         # huberfit.m:172-177
    if problem == "huberfit":
        options["obj"] = lambda x, z: 0.5 * float(np.sum(huber(z)))         # huberfit.m:180
    else:
        options["obj"] = lambda x, z: float(np.sum(np.abs(z)))              # lad.m:148
    results = admm(minx, minz, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results


def huberfit(D, s, options):
    """solvers/huberfit.m:83-186."""
    return _robustfit(D, s, options, "huberfit")


def lad(D, s, options):
    """solvers/lad.m:51-154."""
    return _robustfit(D, s, options, "lad")


def totalvariation(s, lam, options):
    """solvers/totalvariation.m:62-167."""
    t0 = time.perf_counter()
    options = dict(options)
    if not (np.isscalar(lam) and np.real(lam) >= 0):                        # :190-194
        raise MatlabError("Given lambda parameter is not a nonnegative number!")
    if np.ndim(s) > 1 and 1 not in np.shape(s):                             # :197-199
        raise MatlabError("Argument s is not a vector!")
    s = _col(s)
    n = s.shape[0]
    # D = spdiags([ones(n,1) -ones(n,1)], 0:1, n, n)  (:127): D(i,i)=1, D(i,i+1)=-1
    D = sp.diags([np.ones(n), -np.ones(n - 1)], [0, 1], shape=(n, n), format="csr")
    Dt = D.T.tocsr()
    DtD = (Dt @ D).tocsr()
    objective = lambda x, z: 0.5 * float(np.sum((x - s) ** 2)) + lam * float(np.sum(np.abs(x[1:] - x[:-1])))
    args = dict(D=D, Dt=Dt, DtD=DtD, s=s)
    args["lambda"] = lam
    xmin, zmin, _ = getproxops("TotalVariation", args)                      # :148
    options.update(A=D, At=Dt, B=-1, mB=n, nB=n, c=0, m=n, obj=objective)   # :151-161
    results = admm(xmin, zmin, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results


def basispursuit(D, s, options):
    """solvers/basispursuit.m:52-146."""
    t0 = time.perf_counter()
    options = dict(options)
    D = np.asarray(D, dtype=np.float64)
    s = _col(s)
    mD, nD = D.shape
    ms = s.shape[0]
    if mD == nD and mD == ms:                                               # :192-203
        raise MatlabError("Square matrix problem Dx = s; don't need Basis Pursuit to solve this!")
    elif mD > nD and mD == ms:
        raise MatlabError("Overdetermined system Dx = s, as D has more rows thancolumns; use "
                          "Unwrapped ADMM solver for efficiency, instead.")
    elif mD != ms:
        raise MatlabError("The number of rows in matrix D must match the number of rows in "
                          "signal vector s!")
    n = nD
    DDt = D @ D.T                                                           # :116-120
    Dsol = np.linalg.solve(DDt, D)
    ssol = np.linalg.solve(DDt, s)
    P = np.eye(n) - D.T @ Dsol
    q = D.T @ ssol
    minx, minz, _ = getproxops("BasisPursuit", dict(P=P, q=q))              # :127
    options.update(A=1, B=-1, c=0, m=n, nA=n, nB=n, solver="basispursuit")  # :130-137
    options["obj"] = lambda x, z: float(np.sum(np.abs(x)))                  # :140
    results = admm(minx, minz, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results


def basispursuit_factored(D, s, options):
    """Test aid for n where the reference's dense n x n projector does not fit the host (n = 32768: two 8.6 GB
    temporaries): the SAME loop as basispursuit() with the x-update P*(z-u) + q (getProxOps.m:1031) applied as
    v - D'((DD') \\ (D v - s)), which is the algebra of P and q (basispursuit.m:116-120) without forming them.
    tests/test_oracle_known_answers.py pins it to the explicit-P path (1e-13) at sizes where both fit."""
    t0 = time.perf_counter()
    options = dict(options)
    D = np.asarray(D, dtype=np.float64)
    s = _col(s)
    mD, n = D.shape
    if not (mD < n and mD == s.shape[0]):
        raise MatlabError("oracle: basispursuit_factored needs an underdetermined system")
    cho = sla.cho_factor(D @ D.T, lower=True)
    _, minz, _ = getproxops("BasisPursuit", dict(P=np.zeros((1, 1)), q=np.zeros(1)))

    def minx(x, z, u, rho):
        v = z - u
        return v - D.T @ sla.cho_solve(cho, D @ v - s)
    options.update(A=1, B=-1, c=0, m=n, nA=n, nB=n, solver="basispursuit")
    options["obj"] = lambda x, z: float(np.sum(np.abs(x)))
    results = admm(minx, minz, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results


def quadraticprogram(P, q, r, cons1, cons2, options):
    """solvers/quadraticprogram.m:99-257, 'bounded' constraint branch only (:210-216; error checks
    :259-366).  The 'standard' branch (dense KKT solve per iteration) is out of scope."""
    t0 = time.perf_counter()
    options = dict(options)
    P = np.asarray(P, dtype=np.float64)
    q = _col(q)
    nP = P.shape[0]
    if q.shape[0] != nP:
        raise MatlabError("The dimensions of square matrix P and vector q do not match!")
    c1, c2 = np.asarray(cons1, dtype=np.float64), np.asarray(cons2, dtype=np.float64)
    isvec = lambda a: a.ndim <= 1 or 1 in a.shape
    if not (isvec(c1) and isvec(c2)):
        raise MatlabError("oracle: quadraticprogram 'standard' form is out of scope (SURVEY.md section 2)")
    c1, c2 = c1.reshape(-1), c2.reshape(-1)
    if c1.shape[0] != c2.shape[0]:
        raise MatlabError("Lengths of lower and upper bound constraints on solution x do not match!")
    if c1.shape[0] != nP:
        raise MatlabError("Bound vectors do not match predicted length of solution x!")
    if np.array_equal(np.maximum(c1, c2), c1):
        c1, c2 = c2, c1
    elif not np.array_equal(np.maximum(c1, c2), c2):
        raise MatlabError("Given constraint variables do not specify an upper and lower bound on solution x!")
    n = nP
    rho = float(options["rho"]) if "rho" in options else 1.0
    args = dict(P=P, q=q, lb=c1, ub=c2, rho=rho, n=n, constraint="bounded")
    minx, minz, _ = getproxops("quadraticprogram", args)
    options.update(A=1, B=-1, c=0, m=n, nA=n, nB=n)
    options["obj"] = lambda x, z: 0.5 * float(x @ (P @ x)) + float(q @ x) + r
    results = admm(minx, minz, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results


def model(P, Q, r, s, options):
    """solvers/model.m:99-146 (error checks :158-218): minimise 1/2||Px - r||^2 + 1/2||Qz - s||^2, x - z = 0."""
    t0 = time.perf_counter()
    P, Q = np.asarray(P, dtype=np.float64), np.asarray(Q, dtype=np.float64)
    if P.ndim != 2:
        raise MatlabError("Argument P is not a matrix!")
    if Q.ndim != 2:
        raise MatlabError("Argument Q is not a matrix!")
    r, s = np.asarray(r, dtype=np.float64), np.asarray(s, dtype=np.float64)
    isvec = lambda a: a.ndim <= 1 or (a.ndim == 2 and 1 in a.shape)
    if not isvec(r):
        raise MatlabError("Argument r is not a vector!")
    if not isvec(s):
        raise MatlabError("Argument s is not a vector!")
    r, s = _col(r), _col(s)
    if P.shape[0] != Q.shape[0]:
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
    args = dict(PtP=P.T @ P, Ptr=P.T @ r, QtQ=Q.T @ Q, Qts=Q.T @ s, n=n)    # :123-128
    minx, minz, _ = getproxops("Model", args)
    options.update(A=1, B=-1, c=0, m=n, nA=n, nB=n)                         # :133-138
    options["obj"] = lambda x, z: 0.5 * float(np.sum((P @ x - r) ** 2)) + 0.5 * float(np.sum((Q @ z - s) ** 2))
    results = admm(minx, minz, options)
    results["solverruntime"] = time.perf_counter() - t0
    return results
