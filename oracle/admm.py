"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the reference's
generic scaled-dual ADMM driver, ``/root/reference/admm.m``.

PARITY UNPINNED: the reference is MATLAB source, this image has neither MATLAB nor GNU
Octave, and the reference ships no golden vectors.  This file is a
line-by-line NumPy restatement of the algorithm text; every block cites the admm.m lines it
follows.  It is pinned only against hand-derived known answers (tests/test_oracle_*.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package.  The product (``admm_project_b200``) never does.

MATLAB -> Python conventions used throughout the oracle:
  * ``options`` / ``results`` structs are plain dicts; per-iteration arrays are Python lists
    while running and 1-D float64 arrays on return (entry ``[i-1]`` is MATLAB's ``(i)``).
  * vectors are 1-D float64 arrays (MATLAB column vectors).
  * ``error(...)`` becomes ``MatlabError`` carrying the same message text.
  * extension (documented deviation, SURVEY.md section 7 "History arrays"): ``options['history']``
    (default 1) switches the xvals/zvals/uvals/wvals records; the reference always records.
"""
from __future__ import annotations

import math
import time

import numpy as np

EPS = float(np.finfo(np.float64).eps)


class MatlabError(RuntimeError):
    """Stands in for MATLAB's error(...) (the reference has no error codes)."""


def _is_handle(f):
    return callable(f)


def _is_numeric_matrix(a):
    if _is_handle(a):
        return False
    if hasattr(a, "shape") and hasattr(a, "dot"):      # ndarray or scipy.sparse
        return len(a.shape) <= 2
    return isinstance(a, (int, float, np.integer, np.floating))


def _shape2(a):
    """MATLAB size() of a scalar / vector / matrix as (rows, cols)."""
    if isinstance(a, (int, float, np.integer, np.floating)):
        return 1, 1
    sh = a.shape
    if len(sh) == 0:
        return 1, 1
    if len(sh) == 1:
        return sh[0], 1
    return sh[0], sh[1]


def _as_operator(a):
    """``A = @(v) A*v`` (admm.m:120,164,204) for a scalar, dense or sparse matrix."""
    if isinstance(a, (int, float, np.integer, np.floating)) or getattr(a, "shape", None) in ((), (1, 1), (1,)):
        if hasattr(a, "toarray"):      # a 1 x 1 sparse matrix (total variation with n = 1) is a scalar to MATLAB
            a = a.toarray()
        s = float(np.asarray(a).reshape(-1)[0]) if hasattr(a, "shape") else float(a)
        return lambda v: s * v
    return lambda v: a @ v


def _transpose(a):
    if isinstance(a, (int, float, np.integer, np.floating)):
        return a
    return a.T


def _fro(v):
    """norm(v,'fro') for a scalar, vector or matrix."""
    v = np.asarray(v, dtype=np.float64)
    return float(math.sqrt(float(np.sum(v * v))))


def setopt(options, opttext, default):
    """admm.m:780-971.  Returns options[opttext] when the field exists, else the default.
    Quirk kept (admm.m:927-928): the 'Hnormtol' case reads the field ``Hreltol``, so a
    user-set ``options.Hnormtol`` without ``Hreltol`` raises (MATLAB: reference to
    non-existent field)."""
    if opttext in options:
        if opttext == "Hnormtol":
            if "Hreltol" not in options:
                raise MatlabError("Reference to non-existent field 'Hreltol'.")
            return options["Hreltol"]
        return options[opttext]
    return default


def slice_ranges(slices):
    """[runningsum+1, runningsum+slices(i)] bookkeeping used by every sliced prox
    (getProxOps.m:290-297, unwrappedadmm.m:105-110) as 0-based half-open ranges."""
    out, run = [], 0
    for s in slices:
        out.append((run, run + int(s)))
        run += int(s)
    return out


def admm(xminf, zming, options):
    """results = admm(xminf, zming, options) -- admm.m:24."""
    # admm.m:46-49
    if not isinstance(options, dict):
        raise MatlabError("Given options is not a struct! At least pass empty struct!")
    options = dict(options)            # MATLAB value semantics: the caller's struct is untouched
    results = {}

    # admm.m:51-76 -- defaults
    adaptive = setopt(options, "adaptive", 0)
    quiet = setopt(options, "quiet", 1)
    rho = float(setopt(options, "rho", 1.0))
    N = setopt(options, "maxiters", 1000)
    domaxiters = setopt(options, "domaxiters", 0)
    relax = float(setopt(options, "relax", 1))
    parallel = setopt(options, "parallel", "none")
    slices = setopt(options, "slices", 0)
    fast = setopt(options, "fast", 0)
    fasttype = setopt(options, "fasttype", "weak")
    obj = setopt(options, "obj", 0)
    objevals = setopt(options, "objevals", 0)
    convtest = setopt(options, "convtest", 0)
    convtol = setopt(options, "convtol", 1e-10)
    stopcond = setopt(options, "stopcond", "standard")
    nodualerror = setopt(options, "nodualerror", 0)
    ABSTOL = setopt(options, "abstol", 1e-5)
    RELTOL = setopt(options, "reltol", 1e-3)
    HNORMTOL = setopt(options, "Hnormtol", 1e-6)
    m = int(setopt(options, "m", 0))
    nA = int(setopt(options, "nA", 0))
    nB = int(setopt(options, "nB", 0))
    history = setopt(options, "history", 1)          # extension, see module docstring

    if adaptive:
        raise MatlabError("oracle: options.adaptive is out of scope (SURVEY.md section 2, "
                          "admm.m:724-741 is an unfinished experiment)")

    # admm.m:79-111 -- vector c
    if "c" in options:
        c = options["c"]
        carr = np.asarray(c, dtype=np.float64)
        if carr.ndim <= 1 or 1 in carr.shape:
            carr = carr.reshape(-1)
            if m == 0 and carr.size == 1:
                raise MatlabError("Given vector c is scalar and no length m has been provided; "
                                  "unable to infer m - please specify it in options struct.")
            elif carr.size != 1:
                m = carr.size
            c = carr if carr.size != 1 else float(carr[0])
        else:
            raise MatlabError("Given c in constraint Ax + Bz = c is not a vector!")
    else:
        if m > 0:
            if math.floor(m) == m:
                c = np.zeros(m)
            else:
                raise MatlabError("Noninteger size m of c in constraint Ax + Bz = c!")
        else:
            raise MatlabError("Must specify a vector c in constraint Ax + Bz = c!")

    # admm.m:113-161 -- matrix A
    if "A" in options:
        A = options["A"]
        if _is_numeric_matrix(A):
            mA, nAtemp = _shape2(A)
            options["At"] = _transpose(A)               # admm.m:119 (overrides a user At)
            A = _as_operator(A)
        elif _is_handle(A):
            if nA == 0:
                raise MatlabError("Matrix A is a function handle, but no number of columns nA "
                                  "specified for it; cannot infer nA - please specify it in "
                                  "options struct!")
            mA, nAtemp = _shape2(np.asarray(A(np.zeros(nA))))
        else:
            raise MatlabError("Given A in constraint Ax + Bz = c is neither a numeric matrix "
                              "nor function handle of single vector!")
        if mA != m and mA != 1:
            raise MatlabError("Number of rows in matrix A do not match length of column vector "
                              "c in constraint Ax + Bz = c")
        if nA == 0 and nAtemp == 1 and mA == 1:
            raise MatlabError("Given scalar as matrix A with no number of columnsnA specified in "
                              "options struct; cannot infer nA - please specify nA in options!")
        elif nAtemp != 1 and nA != nAtemp:
            nA = nAtemp
    else:
        raise MatlabError("Must specify a matrix A in constraint Ax + Bz = c!")

    # admm.m:163-201 -- transpose of A
    if "At" in options:
        At = options["At"]
        if _is_numeric_matrix(At):
            nAt, mAt = _shape2(At)
            At = _as_operator(At)
        elif _is_handle(At):
            nAt, mAt = _shape2(np.asarray(At(np.zeros(mA))))
        else:
            raise MatlabError("Given At (A transpose) in constraint Ax + Bz = c is neither a "
                              "numeric matrix nor function handle of single vector!")
        if mAt != mA:
            raise MatlabError("Number of columns in At (A transpose) does not match number of "
                              "rows in A, in constraint Ax + Bz = c")
        if nAt != nA and not (nAt == 1 and mAt == 1):
            raise MatlabError("Number of rows in At (A transpose) do not match number of "
                              "columns in A, in constraint Ax + Bz = c")
    else:
        raise MatlabError("Must specify a matrix A in constraint Ax + Bz = c!")

    # admm.m:203-245 -- matrix B
    if "B" in options:
        B = options["B"]
        if _is_numeric_matrix(B):
            mB, nBtemp = _shape2(B)
            B = _as_operator(B)
        elif _is_handle(B):
            if nB == 0:
                raise MatlabError("Matrix B is a function handle, but no number of columns nB "
                                  "specified for it; cannot infer nB - please specify it in "
                                  "options struct!")
            mB, nBtemp = _shape2(np.asarray(B(np.zeros(nB))))
        else:
            raise MatlabError("Given B in constraint Ax + Bz = c is neither a numeric matrix "
                              "nor function handle of single vector!")
        if mB != m and mB != 1:
            raise MatlabError("Number of rows in matrix B do not match length of column vector "
                              "c in constraint Ax + Bz = c")
        if nB == 0 and nBtemp == 1 and mB == 1:
            raise MatlabError("Given scalar as matrix B with no number of columns nBspecified in "
                              "options struct; cannot infer nB - please specify nB in options!")
        elif nBtemp != 1 and nB != nBtemp:
            nB = nBtemp
    else:
        raise MatlabError("Must specify a matrix B in constraint Ax + Bz = c!")

    # admm.m:248
    canEvalObj = bool(objevals) and _is_handle(obj)

    # admm.m:252-259 -- initial iterates
    x = np.array(setopt(options, "x0", np.zeros(nA)), dtype=np.float64).reshape(-1)
    z = np.array(setopt(options, "z0", np.zeros(nB)), dtype=np.float64).reshape(-1)
    u = np.array(setopt(options, "u0", np.zeros(m)), dtype=np.float64).reshape(-1)
    results["x0"], results["z0"], results["u0"] = x.copy(), z.copy(), u.copy()

    # admm.m:262-298 -- algorithm selection (0 vanilla, 1 fast, 2 accelerated)
    alg = 0
    if fast:
        v = z.copy()
        uhat = u.copy()
        acurr = 1.0
        aprev = 1.0
        if fasttype == "weak":
            d = math.inf
            dprev = math.inf
            nrestart = setopt(options, "restart", 0.999)
            if nrestart <= 0 or nrestart >= 1:
                nrestart = 0.999
            DVALTOL = setopt(options, "dvaltol", 1e-8)
            results["dvaltol"] = DVALTOL
            alg = 2
        else:
            alg = 1

    # admm.m:302-313 -- H-norm machinery
    use_hnorm = bool(convtest) or stopcond == "hnorm" or stopcond == "both"
    if use_hnorm:
        def H_norm_sq(wdiff):
            return (rho * _fro(B(wdiff[nA:nA + nB])) ** 2
                    + rho * _fro(wdiff[nA + nB:nA + nB + m]) ** 2)
        w = np.concatenate([x, z, rho * u])
        results["Hnormtol"] = HNORMTOL

    start = time.perf_counter()

    # admm.m:318-330 -- header
    if not quiet:
        if canEvalObj:
            print("%7s\t%20s\t%20s\t%20s\t%20s\t%20s" % (
                "Iteration", "Primal Residual Norm", "Primal Error", "Dual Residual Norm",
                "Dual Error", "Objective Value"))
        else:
            print("%7s\t%20s\t%20s\t%20s\t%20s" % (
                "Iteration", "Primal Residual Norm", "Primal Error", "Dual Residual Norm",
                "Dual Error"))

    # admm.m:334-339
    if N > 0:
        N = int(math.ceil(N))
    else:
        N = 1000

    # admm.m:343-408 -- parallel prox wrappers.  The reference asks the PCT pool for its
    # worker count (gcp, admm.m:346-347); here the count comes from options['workers'].
    if parallel in ("xminf", "zming", "both"):
        from .errorcheck import errorcheck
        workers = int(setopt(options, "workers", 0))
        if workers <= 0:
            raise MatlabError("There are no workers on this machine, cannot perform parallel ADMM!")
        is_cell = isinstance(slices, tuple)      # MATLAB cell {slicesx, slicesz} <-> Python tuple
        if not is_cell and parallel == "both":
            raise MatlabError("For parallelizing both proximal ops, please:\n"
                              "\tSpecify slices for f as a vector in options.slices.\n"
                              "\tSpecify slices for g as a vector in options.slices.")
        elif is_cell and parallel != "both":
            raise MatlabError("Trying to parallelize both proximal operators, but "
                              "options.slices is not a 2 element cell!")
        if is_cell:
            slicesx, slicesz = slices
        elif parallel == "xminf":
            slicesx, slicesz = slices, []
        else:
            slicesx, slicesz = [], slices
        if parallel in ("xminf", "both"):
            xminfi = xminf
            slicesx = errorcheck(slicesx, "slices", "x-slices",
                                 {"workers": workers, "slicelength": len(x)})

            def xminf(x, z, u, rho, _f=xminfi, _s=slicesx):            # parproxf, admm.m:416-436
                return np.concatenate([np.atleast_1d(_f(x, z, u, rho, k + 1))
                                       for k in range(len(_s))])
        if parallel in ("zming", "both"):
            zmingi = zming
            slicesz = errorcheck(slicesz, "slices", "z-slices",
                                 {"workers": workers, "slicelength": len(z)})

            def zming(x, z, u, rho, _g=zmingi, _s=slicesz):            # parproxg, admm.m:447-467
                return np.concatenate([np.atleast_1d(_g(x, z, u, rho, k + 1))
                                       for k in range(len(_s))])

    # admm.m:473-476
    if "preprocess" in options and _is_handle(options["preprocess"]):
        options["preprocess"]()

    pn, dn, pe, de, hn, ob = [], [], [], [], [], []
    xv, zv, uv, wv = [], [], [], []
    vv, uhv, av, dv, rs = [], [], [], [], []
    i = 0
    # admm.m:496-743 -- main loop
    for i in range(1, N + 1):
        zprev = z
        if alg == 0:
            x = np.asarray(xminf(x, z, u, rho), dtype=np.float64).reshape(-1)      # :502
        else:
            aprev = acurr
            uprev = u
            x = np.asarray(xminf(x, v, uhat, rho), dtype=np.float64).reshape(-1)   # :506
            if alg == 2:
                dprev = d

        if relax != 1:                                                            # :515-531
            Axhat = relax * A(x) - (1 - relax) * (B(zprev) - c)
            if alg == 0:
                z = zming(Axhat, z, u, rho)          # Axhat travels in x's slot (:521)
            else:
                z = zming(Axhat, z, uhat, rho)
        else:
            if alg == 0:
                z = zming(x, z, u, rho)
            else:
                z = zming(x, z, uhat, rho)
        z = np.asarray(z, dtype=np.float64).reshape(-1)

        Ax = A(x)                                                                 # :535-536
        Bz = B(z)

        if "altu" not in options:                                                 # :538-560
            if relax != 1:
                u = (u if alg == 0 else uhat) + (Axhat + Bz - c)
            else:
                u = (u if alg == 0 else uhat) + (Ax + Bz - c)
        else:
            if relax != 1:
                u = options["altu"](u, Axhat, Bz, c)
            else:
                u = options["altu"](u, Ax, Bz, c)

        if alg in (1, 2):                                                         # :562-600
            if alg == 1:
                acurr = 0.5 * (1 + math.sqrt(1 + 4 * aprev ** 2))
                v = z + (aprev - 1) / acurr * (z - zprev)
                uhat = u + (aprev - 1) / acurr * (u - uprev)
            else:
                d = 1 / rho * _fro(u - uhat) ** 2 + rho * _fro(B(z - v)) ** 2
                if d < nrestart * dprev:
                    acurr = 0.5 * (1 + math.sqrt(1 + 4 * aprev ** 2))
                    v = z + (aprev - 1) / acurr * (z - zprev)
                    uhat = u + (aprev - 1) / acurr * (u - uprev)
                    rs.append(0)
                else:
                    acurr = 1.0
                    v = zprev
                    uhat = uprev
                    d = dprev / nrestart
                    rs.append(1)
                dv.append(d)
            if history:
                vv.append(np.array(v))
                uhv.append(np.array(uhat))
            av.append(acurr)

        if canEvalObj:                                                            # :603-605
            ob.append(float(obj(x, z)))

        if history:                                                               # :608-610
            xv.append(x.copy()); zv.append(z.copy()); uv.append(np.array(u))

        if "specialnorms" in options and _is_handle(options["specialnorms"]):     # :612-616
            sn = options["specialnorms"](x, z, u, rho)
            pn.append(float(sn[0])); dn.append(float(sn[1]))
        else:                                                                     # :618-637
            if alg == 0:
                pn.append(_fro(Ax + Bz - c))
                dn.append(_fro(rho * At(B(z - zprev))) if not nodualerror else math.nan)
            elif alg == 1:
                pn.append(_fro(Ax + Bz - c))
                dn.append(rho * _fro(At(B(z - v))) if not nodualerror else math.nan)

        if alg in (0, 1):                                                         # :640-658
            M1 = np.size(Ax)
            M2 = np.size(Bz)
            pe.append(math.sqrt(M1) * ABSTOL
                      + RELTOL * max(max(_fro(Ax), _fro(Bz)), _fro(c)))
            de.append(math.sqrt(M2) * ABSTOL + RELTOL * _fro(rho * At(u))
                      if not nodualerror else math.nan)

        if not quiet and alg != 2:                                                # :661-673
            if canEvalObj:
                print("%3d\t%10.4f\t%10.4f\t%10.4f\t%10.4f\t%10.2f" %
                      (i, pn[-1], pe[-1], dn[-1], de[-1], ob[-1]))
            else:
                print("%3d\t%10.4f\t%10.4f\t%10.4f\t%10.4f" % (i, pn[-1], pe[-1], dn[-1], de[-1]))

        if use_hnorm:                                                             # :676-703
            wprev = w
            w = np.concatenate([x, z, rho * u])
            if history:
                wv.append(w)
            hn.append(H_norm_sq(wprev - w))
            if convtest and i >= 2:
                H2 = hn[i - 1]
                H1 = hn[i - 2]
                if alg == 0 and H1 > EPS and H2 > H1 and not ((H2 - H1) <= H1 * convtol):
                    print("Iteration %i: H norms not converging to given relative tolerance: "
                          "%g is not less or equal to tol. %g" % (i, (H2 - H1) / (H1 + EPS), convtol))
                    print("ADMM seems to not be converging! Please check that your proximal "
                          "operators are correct!")
                    # early `return` (admm.m:700): steps/xopt/zopt/uopt/runtime/options unset
                    _pack(results, pn, dn, pe, de, hn, ob, xv, zv, uv, wv, vv, uhv, av, dv, rs,
                          alg, use_hnorm, canEvalObj, history)
                    return results

        stop = False                                                              # :706-722
        if alg == 2 and i >= 2 and abs(d - dprev) <= DVALTOL * dprev:
            stop = True
        elif alg in (0, 1):
            if (stopcond in ("standard", "both")) and \
                    (not domaxiters and pn[-1] < pe[-1] and (nodualerror or dn[-1] < de[-1])):
                stop = True
        if stop:
            break
        if (stopcond in ("hnorm", "both")) and not domaxiters and i > 2 and hn[i - 1] <= HNORMTOL:
            break

    # admm.m:746-767 -- finalise
    _pack(results, pn, dn, pe, de, hn, ob, xv, zv, uv, wv, vv, uhv, av, dv, rs,
          alg, use_hnorm, canEvalObj, history)
    results["steps"] = i
    results["xopt"] = x
    results["zopt"] = z
    results["uopt"] = np.asarray(u, dtype=np.float64)
    if objevals and _is_handle(obj):
        results["objopt"] = float(obj(x, z))
    results["runtime"] = time.perf_counter() - start
    if not quiet:
        print("Elapsed time is %g seconds." % results["runtime"], end="")
        print("Number of steps to convergence: %d" % results["steps"], end="")
    results["options"] = options
    return results


def _pack(results, pn, dn, pe, de, hn, ob, xv, zv, uv, wv, vv, uhv, av, dv, rs,
          alg, use_hnorm, canEvalObj, history):
    f = lambda a: np.asarray(a, dtype=np.float64)
    results["pnorm"], results["dnorm"] = f(pn), f(dn)
    if alg in (0, 1):
        results["perr"], results["derr"] = f(pe), f(de)
    if use_hnorm:
        results["Hnormsq"] = f(hn)
    if canEvalObj:
        results["objevals"] = f(ob)
    if history:
        col = lambda lst: np.stack(lst, axis=1) if lst else np.zeros((0, 0))
        results["xvals"], results["zvals"], results["uvals"] = col(xv), col(zv), col(uv)
        if use_hnorm:
            results["wvals"] = col(wv)
        if alg in (1, 2):
            results["vvals"], results["uhatvals"] = col(vv), col(uhv)
    if alg in (1, 2):
        results["avals"] = f(av)
    if alg == 2:
        results["dvals"], results["restarted"] = f(dv), f(rs)
