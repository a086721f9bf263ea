"""The CPU path beside every GPU config: the oracle (NumPy/SciPy restatement of the reference -- there is
no MATLAB / Octave in the image) timed on the box's host cores, on a BOUNDED sample of each BASELINE.json
config, with the sample stated.  Reported: setup seconds, seconds per iteration, cores.  A baseline, not a
target.   python tools/bench_cpu_configs.py [--scale S]   (S < 1 shrinks every sample, for a quick check)"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (this tool IS the cpu baseline leg)
from admm_project_b200 import generators as gen  # noqa: E402


def timed(fn):
    t0 = time.perf_counter()
    r = fn()
    return time.perf_counter() - t0, r


def entry(sample, full, dt, r):
    steps = max(int(r["steps"]), 1)
    loop = float(r["runtime"])
    return {"sample": sample, "full_config": full, "steps": steps, "call_s": dt, "setup_s": dt - loop,
            "s_per_iter": loop / steps, "iters_per_s": steps / loop}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    sc = a.scale
    out = {"cores": len(os.sched_getaffinity(0)), "kind": "port (oracle/, OpenBLAS)"}
    fixed = {"history": 0, "domaxiters": 1}

    rows = int(15000 * sc)                                   # C3: a quarter of the rows, one of the ten classes
    D, ELL = gen.svm_mnist_like(0, rows, 784, nclass=10)
    D = D + 1e-3 * np.random.RandomState(3).randn(*D.shape)
    np.random.seed(1)
    dt, r = timed(lambda: oracle.linearsvm(D, ELL[:, 0], 0.5, {"history": 0}))     # unwrappedadmm.m forces maxiters = 1000
    out["c3_svm"] = entry("%d x 784, class 0, serial pinv path, run to its own stop test" % rows, "60000 x 784 x 10 classes", dt, r)

    rows = int(262144 * sc)                                  # C4: 1/16 of the rows
    D, s, _ = gen.huber_problem(0, rows, 1024)
    dt, r = timed(lambda: oracle.huberfit(D, s, dict(fixed, maxiters=10)))
    out["c4_huber"] = entry("%d x 1024, 10 iterations" % rows, "4194304 x 1024", dt, r)
    dt, r = timed(lambda: oracle.lad(D, s, dict(fixed, maxiters=10)))
    out["c4_lad"] = entry("%d x 1024, 10 iterations" % rows, "4194304 x 1024", dt, r)
    del D

    n = int((1 << 24) * sc)                                  # C5a: full size, 5 iterations
    s, _ = gen.tv_problem(0, n)
    dt, r = timed(lambda: oracle.totalvariation(s, 1.0, dict(fixed, maxiters=5)))
    out["c5a_tv"] = entry("n = %d, 5 iterations (banded Cholesky solve per iteration)" % n, "n = 2^24", dt, r)

    m, n = int(2048 * sc), int(16384 * sc)                   # C5b: a quarter of the elements
    D, s, _ = gen.bp_problem(0, m, n, density=0.1)
    dt, r = timed(lambda: oracle.basispursuit(D, s, dict(fixed, maxiters=10)))
    out["c5b_bp"] = entry("%d x %d, 10 iterations (dense n x n projector as in basispursuit.m:116-120)" % (m, n), "4096 x 32768", dt, r)
    print("CPUCONFIGS " + json.dumps(out))


if __name__ == "__main__":
    main()
