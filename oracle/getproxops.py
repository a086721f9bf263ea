"""ORACLE (test infrastructure) -- restatement of the in-scope entries of the prox catalogue
``/root/reference/getProxOps.m`` (function ``getproxops``, getProxOps.m:13).  Parity unpinned
(see oracle/admm.py).  Each closure cites the reference lines it follows.

In scope: basispursuit, totalvariation, linearsvm (serial + sliced),
lasso (serial), lad, huberfit, plus the stand-alone z-prox kernels of linearprogram /
quadraticprogram (nonneg ``pos`` :1378-1382/:1422-1426 and box :1470-1474).  Out of scope:
model, linearprogram/quadraticprogram x-updates, covarianceselection, consensus (parfor)
lasso, whose reference implementation is not well defined.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

from .admm import MatlabError, slice_ranges


# --- shims for toolbox functions the reference calls ------------------
def subplus(v):
    """Curve Fitting Toolbox subplus: max(v, 0)."""
    return np.maximum(v, 0.0)


def pos(v):
    """CVX pos: max(v, 0)."""
    return np.maximum(v, 0.0)


def huber(v, M=1.0):
    """CVX huber(x, M): x^2 for |x|<=M, 2M|x|-M^2 otherwise."""
    a = np.abs(v)
    return np.where(a <= M, a * a, 2.0 * M * a - M * M)


def zminSoftThresholding(v, t):
    """getProxOps.m:933-938  sign(v).*subplus(abs(v) - t)."""
    return np.sign(v) * subplus(np.abs(v) - t)


def minz01(s, t):
    """getProxOps.m:1158-1180."""
    y = np.ones(len(s))
    inds = (s >= 1) | (s < (1 - math.sqrt(2 / t)))
    y[inds] = s[inds]
    return y


def zminNonNegative(x, _z, u, _rho):
    """getProxOps.m:1378-1382 / :1422-1426  pos(x + u)."""
    return pos(x + u)


def make_zminBox(lb, ub):
    """getProxOps.m:1470-1474  min(ub, max(lb, x + u))."""
    return lambda x, _z, u, _rho: np.minimum(ub, np.maximum(lb, x + u))


def _tri_lower_solve(L, y):
    return sla.solve_triangular(L, y, lower=True, check_finite=False)


def _tri_upper_solve(U, y):
    return sla.solve_triangular(U, y, lower=False, check_finite=False)


def getproxops(problem, args):
    """[minx, minz, extra] = getproxops(problem, args) -- getProxOps.m:13-917."""
    extra = {}
    if isinstance(problem, str):
        problem = problem.lower()                                           # :37
    else:
        raise MatlabError("Given problem argument is not a string specifying for which problem "
                          "proximal operators are needed!")
    if not isinstance(args, dict):
        raise MatlabError("Given struct args is not a struct containing arguments needed for "
                          "proximal operators for the given problem!")

    if problem == "basispursuit":                                           # :126-142
        P, q = args["P"], args["q"]
        minx = lambda _x, z, u, _rho: P @ (z - u) + q                       # :1027-1032
        minz = lambda x, _z, u, rho: zminSoftThresholding(u + x, 1 / rho)

    elif problem == "totalvariation":                                       # :172-199
        D, Dt, DtD = args["D"], args["Dt"], args["DtD"]
        s, lam = args["s"], args["lambda"]
        n = DtD.shape[0]
        dd = np.asarray(DtD.diagonal(0), dtype=np.float64)
        od = np.asarray(DtD.diagonal(1), dtype=np.float64)

        def minx(_x, z, u, rho):                                            # :1044-1048
            ab = np.zeros((2, n))
            ab[0, 1:] = rho * od          # upper diagonal of (Id + rho*DtD)
            ab[1, :] = 1.0 + rho * dd
            return sla.solveh_banded(ab, s + rho * (Dt @ (z - u)), lower=False,
                                     check_finite=False)
        minz = lambda x, _z, u, rho: zminSoftThresholding(u + D @ x, lam / rho)

    elif problem == "linearsvm":                                            # :256-309
        D, ell, C, loss = args["D"], args["ell"], args["C"], args["lossfunction"]
        if "slices" in args:
            ranges = slice_ranges(args["slices"])
            Di = [D[a:b, :] for a, b in ranges]
            minx = 0

            def minz(x, _z, u, rho, i):                                     # :1120-1143
                a, b = ranges[i - 1]
                Dxplusu = Di[i - 1] @ x + u[a:b]
                v = ell[a:b] * Dxplusu
                if loss != "01":
                    return Dxplusu + ell[a:b] * np.maximum(np.minimum(1 - v, C / rho), 0)
                return ell[a:b] * minz01(v, rho / C)
        else:
            Dplus = args["Dplus"]
            minx = lambda _x, z, u, _rho: Dplus @ (z - u)                   # :1064-1068

            def minz(x, _z, u, rho):                                        # :1084-1103
                Dxplusu = D @ x + u
                v = ell * Dxplusu
                if loss != "01":                # NB: '0-1' is NOT '01' -> hinge
                    return Dxplusu + ell * np.maximum(np.minimum(1 - v, C / rho), 0)
                return ell * minz01(v, rho / C)

    elif problem == "lasso":                                                # :311-456
        if args["parallel"]:
            raise MatlabError("oracle: consensus (parfor) LASSO is out of scope -- the reference "
                              "path returns an all-zero minz (getProxOps.m:1275-1276)")
        D, Dts, lam = args["D"], args["Dts"], args["lambda"]
        L, U, m, n = args["L"], args["U"], args["m"], args["n"]

        def minx(_x, z, u, rho):                                            # :1192-1206
            y = rho * (z - u) + Dts
            if m >= n:
                return _tri_upper_solve(U, _tri_lower_solve(L, y))
            return 1 / rho * y - 1 / rho ** 2 * (D.T @ _tri_upper_solve(U, _tri_lower_solve(L, D @ y)))
        minz = lambda x, _z, u, rho: zminSoftThresholding(u + x, lam / rho)

    elif problem in ("lad", "huberfit"):                                    # :780-912
        R, D, s = args["R"], args["D"], args["s"]
        Rt, Dt = R.T, D.T

        def minx(_x, z, u, _rho):                                           # :1511-1515
            return _tri_upper_solve(Rt, _tri_lower_solve(R, Dt @ (s + z - u)))
        userelax = bool(args.get("userelax", 0))
        if problem == "lad":
            if userelax:                                                    # :808
                minz = lambda x, _z, u, rho: zminSoftThresholding(x + u - s, 1 / rho)
            else:                                                           # :810
                minz = lambda x, _z, u, rho: zminSoftThresholding(D @ x + u - s, 1 / rho)
        else:
            def zminHuber(Ax, u, rho):                                      # :1529-1539
                v = Ax + u - s
                return 1 / (1 + rho) * (rho * v + zminSoftThresholding(v, 1 + 1 / rho))
            if userelax:                                                    # :906-907
                minz = lambda Dxhat, _z, u, rho: zminHuber(Dxhat, u, rho)
            else:                                                           # :910-911
                minz = lambda x, _z, u, rho: zminHuber(D @ x, u, rho)

    elif problem == "quadraticprogram" and args.get("constraint") == "bounded":      # :1441-1474
        P, q, lb, ub, n = args["P"], args["q"], args["lb"], args["ub"], args["n"]
        state = {"rhoprev": None, "R": None}

        def minx(_x, z, u, rho):                                            # xminQuadraticProgramBounded
            if rho != state["rhoprev"]:
                state["R"] = sla.cholesky(P + rho * np.eye(n), lower=False, check_finite=False)   # chol(Pnew): upper R
                state["rhoprev"] = rho
            R = state["R"]
            return _tri_upper_solve(R, _tri_lower_solve(R.T, rho * (z - u) - q))
        minz = make_zminBox(lb, ub)
    elif problem == "model":                                                # :55-110
        PtP, Ptr, QtQ, Qts, n = args["PtP"], args["Ptr"], args["QtQ"], args["Qts"], args["n"]

        # `rhoprev = 0` is never updated (:64), so rho is re-added to the diagonal of a fresh copy and
        # the dense system is re-solved with `\` (SPD -> Cholesky) in every iteration (:967-973, :1004-1011)
        def minx(_x, z, u, rho):                                            # xminModel
            return sla.solve(PtP + rho * np.eye(n), Ptr + rho * (z - u), assume_a="pos", check_finite=False)

        def minz(x, _z, u, rho):                                            # zminModel
            return sla.solve(QtQ + rho * np.eye(n), Qts + rho * (x + u), assume_a="pos", check_finite=False)
    elif problem in ("linearprogram", "quadraticprogram", "covarianceselection"):
        raise MatlabError("oracle: problem '%s' is out of scope (SURVEY.md section 2)" % problem)
    else:
        raise MatlabError("Invalid input for problem - given string is not a solver!")
    return minx, minz, extra
