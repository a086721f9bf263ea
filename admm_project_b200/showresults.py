"""showresults(results, test, options) -- the reference's reporting edge (showresults.m:13-411) over the results
struct the engine fills.

The reference's testers end with `showresults(results, test, options)` unless `quiet` (e.g. testers/lassotest.m:
164-167), so the results struct the gateway returns has to carry every field that function reads:
`steps, runtime, solverruntime, xopt, options, objevals, Hnormsq, Hnormtol, pnorm, perr, dnorm, derr, dvals,
dvaltol` (showresults.m:139-154, 169, 201-233, 297-298, 330-331, 362-363, 385-394).  This mirror

  * prints the same text report, line for line, with MATLAB's `num2str` formatting (:34-166), including the
    reference's quirks: `test.testobjx` prints `test.testobj` (:93-96), and `options.solver = 'unwrappedadmm'`
    leaves the header variable unset (:65-66), which MATLAB reports as an undefined variable;
  * builds the list of plots the reference would draw (:169-409) -- which series, on which of how many subplots,
    with which titles and axis labels -- and draws them only when asked to and matplotlib is importable (plotting
    itself is out of scope; the plan is what pins the results-struct contract);
  * `save_mat` writes results / test / options as MATLAB structs (column vectors, n x steps histories) so the
    reference's own showresults.m can load and render a run of this engine: the round trip through the gateway.
"""
from __future__ import annotations

import math
import sys

import numpy as np

from .admm import MatlabError

_HEADERS = {
    "model": "MODEL", "basispursuit": "BASIS PURSUIT", "covarianceselection": "SPARSE INVERSE COVARIANCE SELECTION",
    "huberfit": "HUBER FITTING", "lad": "LEAST ABSOLUTE DEVIATIONS", "lasso": "LASSO",
    "linearprogram": "LINEAR PROGRAMMING", "linearsvm": "LINEAR SUPPORT VECTOR MACHINE",
    "quadraticprogram": "QUADRATIC PROGRAMMING", "totalvariation": "TOTAL VARIATION MINIMIZATION",
}                                                                         # showresults.m:43-68


def num2str(x):
    """MATLAB num2str for a real scalar: integers print as integers, anything else with
    max(floor(log10(|x|)), 0) + 5 significant digits ('%.Ng', trailing zeros dropped)."""
    x = float(np.asarray(x).reshape(-1)[0]) if np.size(x) else float("nan")
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == math.floor(x) and abs(x) < 1e15:
        return "%d" % int(x)
    digits = max(int(math.floor(math.log10(abs(x)))), 0) + 5
    s = "%.*g" % (digits, x)
    if "e" in s:                                   # MATLAB prints a two-digit exponent with a sign: 1.2346e-05
        mant, exp = s.split("e")
        s = "%se%s%02d" % (mant, exp[0], int(exp[1:]))
    return s


def _has(d, k):
    return isinstance(d, dict) and k in d and d[k] is not None


def showresults(results=None, test=None, options=None, file=None, plot=False):
    """Text report + plot plan of one run.  Returns {'lines': [...], 'plots': [...], 'nplots': int}."""
    if results is None:
        raise MatlabError("No structs given to show results for! No arguments given.")      # showresults.m:25-26
    test = {} if test is None else test
    options = {} if options is None else dict(options)
    out = file if file is not None else sys.stdout
    lines = []

    def disp(s):
        lines.append(s)
        print(s, file=out)

    disp(" ")
    solver = "ADMM"
    if _has(options, "solver"):
        name = str(options["solver"]).lower()
        if name == "unwrappedadmm":                # :65-66 sets nothing; the next use of `solver` fails in MATLAB
            raise MatlabError("Undefined function or variable 'solver'.")
        solver = _HEADERS.get(name, "ADMM")
    if _has(options, "lossfunction"):
        disp("%s EXECUTION AND TEST RESULTS FOR %s ---" % (solver, str(options["lossfunction"]).upper()))
    else:
        disp("%s EXECUTION AND TEST RESULTS ---" % solver)
    for key, text, src in (("trueobjopt", "True optimal objective value: ", "trueobjopt"),
                           ("testobj", "Test's original objective value: ", "testobj"),
                           ("testobjx", "Test's original objective value for x: ", "testobj"),     # :93-96 prints testobj
                           ("objoptx", "ADMM's optimal objective value for x: ", "objoptx"),
                           ("admmopt", "ADMM's optimal objective value for (x, z): ", "admmopt"),
                           ("objerror", "Relative error in ADMM's objective: ", "objerror"),
                           ("constrainterror", "Average error in components of constraint D*x_opt = s: ", "constrainterror"),
                           ("constraintresidual", "Residual ||D*x_opt - s||: ", "constraintresidual"),
                           ("xerror", "Average error in components of x_admm: ", "xerror"),
                           ("xresidual", "Residual ||x* - x_admm||: ", "xresidual")):
        if _has(test, key):
            if not _has(test, src):
                raise MatlabError("Reference to non-existent field '%s'." % src)
            disp(text + num2str(test[src]))
    if _has(results, "steps"):
        disp("Number of iteration steps performed: " + num2str(results["steps"]))
    if _has(results, "runtime"):
        disp("Runtime of ADMM call: %s seconds." % num2str(results["runtime"]))
    if _has(results, "solverruntime"):
        disp("Overall runtime of solver: %s seconds." % num2str(results["solverruntime"]))
    if _has(test, "failed"):
        disp("TEST UNSUCCESSFUL!" if test["failed"] else "Test successful!")
    if _has(test, "failreason"):
        disp("   " + str(test["failreason"]))

    # ---- the plots the reference would draw (showresults.m:169-409) ----------------------------------------------
    if not _has(results, "steps"):
        raise MatlabError("Reference to non-existent field 'steps'.")          # :169  N = results.steps
    N = int(results["steps"])
    it = np.arange(1, N + 1)
    optline = None
    if _has(test, "trueobjopt"):
        optline = ("True optimal objective value", float(test["trueobjopt"]))
    elif _has(test, "testobj"):
        optline = ("Test's original objective value", float(test["testobj"]))
    if optline is not None and optline[1] == 0:    # `if optline ~= 0` (:245): an all-zero line counts as absent
        optline = None
    plots = []
    if _has(test, "D") and _has(test, "s") and _has(test, "testx"):           # :182-190
        tx = np.asarray(test["testx"]).reshape(-1)
        plots.append(dict(kind="signal", figure=1, title="Denoising Results For Noisy Signal",
                          xlabel="Signal component i", ylabel="Signal value at i",
                          series=[("Noisy signal x", np.arange(1, tx.size + 1), tx, "ko"),
                                  ("ADMM's denoised x_{opt}", np.arange(1, tx.size + 1),
                                   np.asarray(results["xopt"]).reshape(-1), "r")]))
    ro = results.get("options") or {}
    nplots = 3 if (ro.get("algorithm") == "fast" and ro.get("fasttype") == "weak") else 4    # :202-210
    for k in ("objevals", "Hnormsq", "pnorm", "dnorm"):                                       # :214-234
        if not _has(results, k):
            nplots -= 1
    sofar = 0

    def slot():
        nonlocal sofar
        if nplots != 1:
            sofar += 1
            if sofar > nplots:
                raise MatlabError("Index exceeds number of subplots.")
        return sofar

    lossy = _has(options, "tester") and options["tester"] == "linearsvm"
    loss = str(options.get("lossfunction", ""))
    if nplots > 0 and _has(results, "objevals"):
        obj = np.asarray(results["objevals"], dtype=float).reshape(-1)
        series = [("ADMM's objective value", np.arange(1, obj.size + 1), obj, "-k")]
        if optline is not None:
            series.insert(0, (optline[0], it, np.full(N, optline[1]), "b--"))
        idx = slot()
        plots.append(dict(kind="objective", subplot=idx, title="Plot of objective value for each iteration",
                          ylabel="Objective", xlabel="Iteration k" if nplots == 1 else None, series=series))
    if nplots > 0 and _has(results, "Hnormsq"):
        hn = np.asarray(results["Hnormsq"], dtype=float).reshape(-1)
        if not _has(results, "Hnormtol"):
            raise MatlabError("Reference to non-existent field 'Hnormtol'.")
        idx = slot()
        plots.append(dict(kind="hnorm", subplot=idx, logy=True,
                          title="Plot of H-Norm Squared Residuals (w = [x^T z^T u^T]^T)" +
                                (" for %s loss function" % loss if lossy else ""),
                          ylabel="||w^{k - 1} - w^k||_H^2", xlabel="Iteration k" if (nplots == 1 or idx == nplots) else None,
                          series=[("H-Norm", np.arange(1, hn.size + 1), np.maximum(1e-8, hn), "k"),
                                  ("Threshold", it, np.full(N, float(results["Hnormtol"])), "b--")]))
    for key, err, title, lab, ylabel in (
            ("pnorm", "perr", "Plot of Primal Residual Norm", ("Primal Norm", "Primal Error"), "||Ax^k - Bz^k - c||_2"),
            ("dnorm", "derr", "Plot of Dual Residual Norm", ("Dual Norm", "Dual Error"), "||\\rho*A^T*B*(z^k-z^{k-1})||_2")):
        if nplots > 0 and _has(results, key):
            v = np.asarray(results[key], dtype=float).reshape(-1)
            if not _has(results, err):
                raise MatlabError("Reference to non-existent field '%s'." % err)
            e = np.asarray(results[err], dtype=float).reshape(-1)
            if v.size != N or e.size != N:         # semilogy(1:N, results.pnorm, ...) needs N values (:297-298, :330-331)
                raise MatlabError("Vectors must be the same length.")
            idx = slot()
            plots.append(dict(kind=key, subplot=idx, logy=True, title=title + (" for %s loss function" % loss if lossy else ""),
                              ylabel=ylabel, xlabel="Iteration k" if (nplots == 1 or idx == nplots) else None,
                              series=[(lab[0], it, np.maximum(1e-8, v), "k"), (lab[1], it, e, "b--")]))
    if nplots > 0 and _has(results, "dvals"):
        dv = np.asarray(results["dvals"], dtype=float).reshape(-1)
        if not _has(results, "dvaltol"):
            raise MatlabError("Reference to non-existent field 'dvaltol'.")
        if dv.size != N:
            raise MatlabError("Vectors must be the same length.")
        idx = slot()
        plots.append(dict(kind="dvals", subplot=idx, logy=True,
                          title="Plot of Accelerated ADMM's Residual Norms" + (" for %s loss function" % loss if lossy else ""),
                          ylabel="1/\\rho||u^k - u_{hat}^k||^2 + \\rho||B(z^k - z_{hat}^k)||^2", xlabel="Iteration k",
                          series=[("Accelerated Residual Norm", it, np.maximum(1e-8, dv), "k"),
                                  ("Convergence point", it, np.full(N, float(results["dvaltol"])), "b--")]))
    if plot:
        _draw(plots, nplots)
    return {"lines": lines, "plots": plots, "nplots": nplots}


def _draw(plots, nplots):
    try:
        import matplotlib.pyplot as plt
    except ImportError as ex:                       # plotting is optional; the report and the plan are not
        raise RuntimeError("showresults(plot=True) needs matplotlib") from ex
    fig = None
    for p in plots:
        if p["kind"] == "signal":
            plt.figure()
            ax = plt.gca()
        else:
            if fig is None:
                fig = plt.figure()
            ax = fig.add_subplot(max(nplots, 1), 1, max(p["subplot"], 1))
        for label, x, y, style in p["series"]:
            (ax.semilogy if p.get("logy") else ax.plot)(x, y, {"ko": "ko", "r": "r", "b--": "b--", "-k": "-k", "k": "k"}[style],
                                                        label=label, linewidth=2)
        ax.set_title(p["title"])
        ax.set_ylabel(p["ylabel"])
        if p.get("xlabel"):
            ax.set_xlabel(p["xlabel"])
        ax.legend()


# ---- .mat round trip: what the reference's own showresults.m loads -------------------------------------------------
_COLUMN_FIELDS = ("xopt", "zopt", "uopt", "x0", "z0", "u0", "objevals", "Hnormsq", "pnorm", "dnorm", "perr", "derr", "dvals",
                  "avals", "restarted")


def _matlab_value(v):
    if isinstance(v, dict):
        return {k: _matlab_value(x) for k, x in v.items() if x is not None and not callable(x)}
    if isinstance(v, (bool, np.bool_)):
        return float(v)
    if isinstance(v, (int, float, np.integer, np.floating)):
        return float(v)
    if isinstance(v, str):
        return v
    a = np.asarray(v)
    if a.dtype == object:
        return None
    return a.astype(float) if a.dtype.kind in "iub" else a


def to_matlab_struct(d):
    """A dict as MATLAB would hold it: numeric scalars double, 1-D histories and iterates as COLUMN vectors
    (admm.m:746-767 returns columns), nested dicts as nested structs, callables and None dropped."""
    out = {}
    for k, v in d.items():
        if v is None or callable(v):
            continue
        mv = _matlab_value(v)
        if mv is None:
            continue
        if isinstance(mv, np.ndarray) and mv.ndim == 1 and (k in _COLUMN_FIELDS or mv.size > 1):
            mv = mv.reshape(-1, 1)
        out[k] = mv
    return out


def save_mat(path, results, test=None, options=None):
    """Write results / test / options as MATLAB structs: `load(path); showresults(results, test, options)` then runs
    the reference's reporting on a run of this engine."""
    from scipy.io import savemat
    savemat(path, {"results": to_matlab_struct(results), "test": to_matlab_struct(test or {}),
                   "options": to_matlab_struct(options or {})}, oned_as="column")


def load_mat(path):
    """The structs of `save_mat` back as dicts (1-D arrays flattened, scalars as floats)."""
    from scipy.io import loadmat
    raw = loadmat(path, squeeze_me=True, struct_as_record=False)

    def conv(o):
        if hasattr(o, "_fieldnames"):
            return {f: conv(getattr(o, f)) for f in o._fieldnames}
        if isinstance(o, np.ndarray) and o.ndim == 0:
            return conv(o.item())
        return o
    return {k: conv(raw[k]) for k in ("results", "test", "options") if k in raw}
