"""Python mirrors of the reference's tester scripts (testers/*.m): the same seeded instance recipe
(`generators.py`), the options each tester forces, and its pass criterion -- `[results, test] =
<name>test(seed, ..., errtol, quiet, options)`.  `solvers` is any namespace with the reference's solver
signatures: this package (the device engine, the default) or, in the CPU tests, the oracle -- the
criteria are evaluated here in NumPy either way.  `solvertester` repeats a tester over sizes like
testers/solvertester.m:100-240."""
from __future__ import annotations

import numpy as np

from . import generators as gen


def _huber1(v):
    a = np.abs(v)
    return np.where(a <= 1.0, a * a, 2.0 * a - 1.0)


def _solvers(solvers):
    if solvers is not None:
        return solvers, {}
    import admm_project_b200 as pkg
    return pkg, {}


def _finish(results, test, failed, reason, quiet=1, options=None):
    test.update(failed=int(bool(failed)), failreason=reason, steps=results.get("steps"))
    if not quiet:          # testers/*test.m end with `if ~quiet, showresults(results, test, options); end`
        from .showresults import showresults
        test["report"] = showresults(results, test, options or {})
    return results, test


def _finish_q(quiet, options, results, test, failed, reason):
    return _finish(results, test, failed, reason, quiet, options)


def lassotest(seed=0, rows=256, cols=64, errtol=1e-3, quiet=1, options=None, solvers=None, **kw):
    """testers/lassotest.m:31-170; demo size 2^8 x 2^6 (:94-95); pass: obj(xopt) < obj(testx) (:143-147)."""
    S, _ = _solvers(solvers)
    D, s, lam, testx = gen.lasso_problem(seed, rows, cols)
    o = dict(options or {}, objevals=1)                                     # :131
    results = S.lasso(D, s, lam, o, **kw)
    obj = lambda x: 0.5 * np.sum((D @ x - s) ** 2) + lam * np.sum(np.abs(x))
    test = dict(D=D, s=s, lam=lam, testx=testx, trueobjopt=obj(testx), objopt=obj(results["xopt"]),
                admmopt=results.get("objopt"), errtol=errtol)
    ok = test["objopt"] < test["trueobjopt"]
    return _finish_q(quiet, dict(o, solver="lasso"), results, test, not ok, "ADMM's objective %s the objective of the generating signal" %
                   ("is below" if ok else "is NOT below"))


def linearsvmtest(seed=0, mpos=128, mneg=128, sep=0.2, errtol=0.05, quiet=1, options=None, solvers=None, C=0.5, **kw):
    """testers/linearsvmtest.m:33-250; two runs, hinge and the string '0-1' -- which the prox treats as hinge
    (strcmp(loss,'01'), getProxOps.m:1094) while the objective handle becomes the sign variant
    (linearsvm.m:231-237).  pass: obj < trueobj and |1 + x2/x1| <= errtol (:180-192)."""
    S, _ = _solvers(solvers)
    D, ell = gen.svm_problem(seed, mpos, mneg, sep)
    o = dict(options or {}, objevals=1, convtest=1)                         # :148-149
    true = np.array([1.0, -1.0])
    trueobj = 0.5 * float(true @ true) + C * float(np.sum(np.maximum(np.sign(1 - ell * (D @ true)), 0)))   # :151-152
    rh = S.linearsvm(D, ell, C, o, **kw)
    r01 = S.linearsvm(D, ell, C, dict(o, lossfunction="0-1"), **kw)
    objh = lambda x: 0.5 * float(x @ x) + C * float(np.sum(np.maximum(1 - ell * (D @ x), 0)))
    obj01 = lambda x: 0.5 * float(x @ x) + C * float(np.sum(np.maximum(np.sign(1 - ell * (D @ x)), 0)))
    tests = []
    for r, obj in ((rh, objh), (r01, obj01)):
        x = r["xopt"]
        relerr = abs(1 - (-x[1] / x[0]))
        tests.append(dict(truexopt=true, trueobjopt=trueobj, xopt=x, objopt=obj(x), admmopt=r.get("objopt"),
                          relerror=relerr, failed=int(not (obj(x) < trueobj and relerr <= errtol)), steps=r["steps"]))
    test = dict(D=D, ell=ell, hinge=tests[0], zero_one=tests[1], errtol=errtol, objopt=tests[0]["objopt"],
                trueobjopt=trueobj)
    if not quiet:          # linearsvmtest.m:231-237: one report per loss function
        from .showresults import showresults
        test["reports"] = [showresults(rh, tests[0], dict(o, tester="linearsvm", lossfunction="hinge")),
                           showresults(r01, tests[1], dict(o, tester="linearsvm", lossfunction="01"))]
    return _finish(rh, test, tests[0]["failed"] or tests[1]["failed"], "hinge failed=%d, '0-1' failed=%d" %
                   (tests[0]["failed"], tests[1]["failed"]))


def huberfittest(seed=0, rows=2048, cols=128, errtol=1e-3, quiet=1, options=None, solvers=None, **kw):
    """testers/huberfittest.m:31-190; pass: 1/2*sum(huber(D*xopt - s)) below the generating x's (:154-158)."""
    S, _ = _solvers(solvers)
    D, s, testx = gen.huber_problem(seed, rows, cols)
    o = dict(options or {}, objevals=1, convtest=1)                         # :137-139
    results = S.huberfit(D, s, o, **kw)
    f = lambda x: 0.5 * float(np.sum(_huber1(D @ x - s)))
    test = dict(D=D, s=s, testx=testx, trueobjopt=f(testx), objopt=f(results["xopt"]), admmopt=results.get("objopt"))
    return _finish_q(quiet, dict(o, solver="huberfit"), results, test, not (test["objopt"] <= test["trueobjopt"]), "objective vs generating signal")


def ladtest(seed=0, rows=1024, cols=128, errtol=1e-3, quiet=1, options=None, solvers=None, **kw):
    """testers/ladtest.m:31-200; pass: ||xtrue - xopt|| < errtol and |obj - trueobj| <= errtol*trueobj (:149-168)."""
    S, _ = _solvers(solvers)
    D, s, xtrue = gen.lad_problem(seed, rows, cols)
    o = dict(options or {}, objevals=1, convtest=1)                         # :130-132
    results = S.lad(D, s, o, **kw)
    x = results["xopt"]
    trueobj, obj = float(np.sum(np.abs(D @ xtrue - s))), float(np.sum(np.abs(D @ x - s)))
    xres = float(np.linalg.norm(xtrue - x))
    test = dict(D=D, s=s, truexopt=xtrue, trueobjopt=trueobj, objopt=obj, xresidual=xres,
                xerror=float(np.sum(np.abs(xtrue - x)) / x.size), admmopt=results.get("objopt"))
    return _finish_q(quiet, dict(o, solver="lad"), results, test, not (xres < errtol and abs(obj - trueobj) <= errtol * trueobj),
                   "xresidual %.3g, objective %.6g vs %.6g" % (xres, obj, trueobj))


def totalvariationtest(seed=0, rows=128, errtol=1e-3, quiet=1, options=None, solvers=None, lam=1.0, **kw):
    """testers/totalvariationtest.m:30-190; pass: objective(xopt) < objective(truth) (:151-155)."""
    S, _ = _solvers(solvers)
    s, truth = gen.tv_problem(seed, rows)
    o = dict(options or {}, objevals=1, maxiters=10000)                     # :130-131
    results = S.totalvariation(s, lam, o, **kw)
    obj = lambda x: 0.5 * float(np.sum((x - s) ** 2)) + lam * float(np.sum(np.abs(np.diff(x))))
    test = dict(s=s, truth=truth, trueobjopt=obj(truth), objopt=obj(results["xopt"]), admmopt=results.get("objopt"))
    return _finish_q(quiet, dict(o, tester="totalvariation"), results, test, not (test["objopt"] < test["trueobjopt"]), "objective vs the clean signal")


def basispursuittest(seed=0, rows=64, cols=128, errtol=1e-3, quiet=1, options=None, solvers=None, density=1.0, **kw):
    """testers/basispursuittest.m:31-180; pass: ||testx||_1 >= ||xopt||_1 and mean relative constraint error
    <= errtol (:120-125).  The reference draws the truth with sprandn(cols, 1, 0.1*cols) (:109) -- a density far
    above 1, which MATLAB clamps, i.e. a DENSE truth: that is the default here (density = 1.0); pass 0.1 for the
    sparse-recovery instance SURVEY.md section 8d settles on for config C5b."""
    S, _ = _solvers(solvers)
    D, s, testx = gen.bp_problem(seed, rows, cols, density=density)
    o = dict(options or {}, objevals=1, maxiters=10000, convtest=0)         # :121-123
    results = S.basispursuit(D, s, o, **kw)
    x = results["xopt"]
    cerr = float(np.sum(np.abs((D @ x - s) / (D @ x))) / s.size)           # :119
    test = dict(D=D, s=s, testx=testx, trueobjopt=float(np.sum(np.abs(testx))), objopt=float(np.sum(np.abs(x))),
                constrainterror=cerr)
    return _finish_q(quiet, dict(o, solver="basispursuit"), results, test, not (test["trueobjopt"] >= test["objopt"] and cerr <= errtol),
                   "l1 norm %.6g vs %.6g, constraint error %.3g" % (test["objopt"], test["trueobjopt"], cerr))


def modeltest(seed=0, rows=128, cols=128, errtol=1e-3, quiet=1, options=None, solvers=None, **kw):
    """testers/modeltest.m:30-200; pass: |1 - obj/trueobj| <= errtol and ||truex - xopt|| <= errtol (:133-141)."""
    S, _ = _solvers(solvers)
    P, Q, r, s, truex = gen.model_problem(seed, rows, cols)
    o = dict(options or {}, objevals=1, maxiters=10000, convtest=1, stopcond="both")   # :124-129
    results = S.model(P, Q, r, s, o, **kw)
    x = results["xopt"]
    obj = lambda v: 0.5 * float(np.sum((P @ v - r) ** 2)) + 0.5 * float(np.sum((Q @ v - s) ** 2))
    objerr, xres = abs(1 - obj(x) / obj(truex)), float(np.linalg.norm(truex - x))
    test = dict(P=P, Q=Q, r=r, s=s, truexopt=truex, trueobjopt=obj(truex), objopt=obj(x), objerror=objerr, xresidual=xres,
                admmopt=results.get("objopt"))
    return _finish_q(quiet, dict(o, solver="model"), results, test, not (objerr <= errtol and xres <= errtol), "objerror %.3g, xresidual %.3g" % (objerr, xres))


TESTERS = {"lasso": lassotest, "linearsvm": linearsvmtest, "huberfit": huberfittest, "lad": ladtest,
           "totalvariation": totalvariationtest, "basispursuit": basispursuittest, "model": modeltest}


def solvertester(solver, minscale=4, maxscale=7, trials=1, seed=0, solvers=None, **kw):
    """testers/solvertester.m:100-240 for the in-scope solvers: sizes 2^scale (skinny 2^s x 2^(s-3) for lasso /
    lad / huber :593-596, ceil(n/5) x n for basis pursuit :407-408, 2^s + 2^s points for the SVM :537-538),
    default errtol 1e-3 (0.05 for the SVM) :115-123; any failure is reported (:233-240).  Per-trial seeds come
    from one RandomState(seed) instead of floor(rand*intmax) (:157)."""
    rs = np.random.RandomState(seed)
    out = []
    for scale in range(minscale, maxscale + 1):
        for _ in range(trials):
            sd = int(rs.randint(0, 2 ** 31 - 1))
            n = 2 ** scale
            if solver in ("lasso", "lad", "huberfit"):
                _, t = TESTERS[solver](sd, n, max(n // 8, 2), 1e-3, 1, {}, solvers, **kw)
            elif solver == "basispursuit":
                _, t = basispursuittest(sd, int(np.ceil(n / 5)), n, 1e-3, 1, {}, solvers, **kw)
            elif solver == "linearsvm":
                _, t = linearsvmtest(sd, n, n, 0.2, 0.05, 1, {}, solvers, **kw)
            elif solver == "totalvariation":
                _, t = totalvariationtest(sd, n, 1e-3, 1, {}, solvers, **kw)
            elif solver == "model":
                _, t = modeltest(sd, n, n, 1e-3, 1, {}, solvers, **kw)
            else:
                raise ValueError("solvertester: '%s' is not an in-scope solver" % solver)
            out.append(dict(scale=scale, seed=sd, failed=t["failed"], steps=t["steps"], failreason=t["failreason"]))
    return dict(solver=solver, trials=out, failures=sum(t["failed"] for t in out))
