"""Per-iteration measurements of the remaining BASELINE.json configs on one B200:
C1 lasso 1500 x 5000 (the reference's own CPU-runnable tester shape, with the oracle timed beside it),
C5a total variation n = 2^24, C5b basis pursuit 4096 x 32768.  (C2: bench.py; C3/C4: bench_unwrapped.py)"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (CPU baseline leg only)
from admm_project_b200 import Engine, basispursuit, lasso, totalvariation  # noqa: E402
from admm_project_b200 import generators as gen  # noqa: E402

HBM = 6547.2


def raw_us(eng, o, which, reps):
    import torch
    eng.iterate_raw(o, which, 5)
    eng.synchronize()
    t0 = time.perf_counter()
    eng.iterate_raw(o, which, reps)
    eng.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def main():
    which = sys.argv[1:] or ["c1", "c5a", "c5b"]
    eng = Engine(0)
    out = {}
    if "c1" in which:
        D, s, lam, testx = gen.lasso_problem(0, 1500, 5000)
        opts = {"objevals": 1, "history": 0}
        lasso(D, s, lam, opts, engine=eng)
        t0 = time.perf_counter()
        r = lasso(D, s, lam, opts, engine=eng)
        wall = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref = oracle.lasso(D, s, lam, opts)
        cpu = time.perf_counter() - t0
        o = eng.default_options()
        us = raw_us(eng, o, 0, 200)
        byt = 2 * 1500 * 5000 * 8 + 1500 * 1501 * 8
        out["c1_lasso_1500x5000"] = {"steps": r["steps"], "ref_steps": ref["steps"], "gpu_call_ms": wall * 1e3,
                                     "setup_ms": r["engine"]["setup_ms"], "loop_ms": r["engine"]["loop_ms"],
                                     "cpu_oracle_call_ms": cpu * 1e3, "cpu_cores": len(os.sched_getaffinity(0)),
                                     "us_per_iter": us, "GBs": byt / us / 1e3, "frac_hbm": byt / us / 1e3 / HBM,
                                     "rel_err_x": float(np.linalg.norm(r["xopt"] - ref["xopt"]) / np.linalg.norm(ref["xopt"]))}
    if "c5a" in which:
        n = 1 << 24
        s, truth = gen.tv_problem(0, n)
        opts = {"history": 0, "maxiters": 100, "domaxiters": 1}
        r = totalvariation(s, 1.0, opts, engine=eng)
        r5 = totalvariation(s, 1.0, dict(opts, maxiters=500, check_every=50), engine=eng)   # device-timed, warm
        o = eng.default_options()
        o.history = 0
        us = raw_us(eng, o, 0, 50)
        usx = raw_us(eng, o, 1, 50)
        usp = raw_us(eng, o, 2, 50)
        byt = 9 * n * 8                               # 9 vector passes
        out["c5a_tv_2^24"] = {"loop_us_per_iter_500_device": r5["engine"]["loop_ms"] * 1e3 / r5["steps"], "us_per_iter": us, "x_solve_us": usx, "prox_us": usp, "loop_ms_100": r["engine"]["loop_ms"],
                              "GBs_algorithmic": byt / us / 1e3, "frac_hbm": byt / us / 1e3 / HBM,
                              "GBs_actual_10_passes": 10 * n * 8 / us / 1e3}
    if "c5b" in which:
        m, n = 4096, 32768
        D, s, testx = gen.bp_problem(0, m, n, density=0.1)
        opts = {"history": 0, "maxiters": 100, "domaxiters": 1}
        r = basispursuit(D, s, opts, engine=eng)
        o = eng.default_options()
        o.history = 0
        us = raw_us(eng, o, 0, 50)
        byt = 2 * m * n * 8 + m * (m + 1) * 8
        out["c5b_bp_4096x32768"] = {"us_per_iter": us, "setup_ms": r["engine"]["setup_ms"], "setup": eng.setup_phases(),
                                    "GBs": byt / us / 1e3, "frac_hbm": byt / us / 1e3 / HBM,
                                    "constraint_rel_err_after_100": float(np.linalg.norm(D @ r["xopt"] - s) / np.linalg.norm(s))}
    print("CONFIGS " + json.dumps(out))


if __name__ == "__main__":
    main()
