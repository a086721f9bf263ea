"""One-pass x-update (csrc/symtri.cuh) against SciPy's substitutions, and its device time, at several factor sizes.
    python tools/check_symtri.py            (on a B200)"""
import os
import sys
import time

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_project_b200 import Engine, _lib as L  # noqa: E402
from admm_project_b200 import generators as gen  # noqa: E402


def main():
    import torch
    eng = Engine(0)
    ok = True
    for m, n in ((300, 64), (500, 129), (4000, 784), (6000, 1500), (9000, 4096), (16384, 8192)):
        D, s, lam, _ = gen.lasso_problem_big(0, m, n)
        eng.setup_lasso(D, s, 1.0)
        Lf = eng.get_factor()
        b = np.random.RandomState(1).randn(n)
        exact = sla.solve_triangular(Lf.T, sla.solve_triangular(Lf, b, lower=True), lower=False)
        x = eng.factor_solve(b, L.XSOLVE_INVFACTOR)
        err = float(np.linalg.norm(x - exact) / np.linalg.norm(exact))
        eng.set_lambda(lam)
        o = eng.default_options()
        eng.iterate_raw(o, 1, 10)
        eng.synchronize()
        t0 = time.perf_counter()
        eng.iterate_raw(o, 1, 200)
        eng.synchronize()
        us = (time.perf_counter() - t0) / 200 * 1e6
        eng.iterate_raw(o, 0, 10)
        eng.synchronize()
        t0 = time.perf_counter()
        eng.iterate_raw(o, 0, 200)
        eng.synchronize()
        us_it = (time.perf_counter() - t0) / 200 * 1e6
        gbs = n * (n + 1) * 8 / (us * 1e-6) / 1e9
        print("symtri n=%5d  rel.err=%.2e  x-update %.1f us (%.0f GB/s on the 2-read count)  iteration %.1f us" % (n, err, us, gbs, us_it), flush=True)
        ok = ok and err < 1e-12
    print("SYMTRI_OK" if ok else "SYMTRI_FAIL")
    eng.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
