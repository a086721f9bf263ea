"""Eager loop vs CUDA-graph bursts (options.graph) on the launch-bound shapes: C1 lasso 1500 x 5000,
one C3 shard (7500 x 784, the 8-GPU row block of the MNIST-shaped SVM), a small Huber fit and a
small total-variation chain.  Device time of the loop (results.engine.loop_ms) per iteration."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_project_b200 import Engine, huberfit, lasso, linearsvm, totalvariation  # noqa: E402
from admm_project_b200 import generators as gen  # noqa: E402


def per_iter(run, iters):
    out = {}
    for g in (0, 1):
        best = 1e30
        for _ in range(3):
            r = run({"history": 0, "domaxiters": 1, "maxiters": iters, "check_every": 50, "graph": g})
            best = min(best, r["engine"]["loop_ms"] * 1e3 / r["steps"])   # unwrappedadmm.m:81 forces maxiters = 1000
        out["graph" if g else "eager"] = best
    out["speedup"] = out["eager"] / out["graph"]
    return out


def main():
    eng = Engine(0)
    res = {}
    D, s, lam, _ = gen.lasso_problem(0, 1500, 5000)
    res["c1_lasso_1500x5000_us"] = per_iter(lambda o: lasso(D, s, lam, o, engine=eng), 400)
    D, ELL = gen.svm_mnist_like(0, 7500, 784, nclass=2)
    D = D + 1e-3 * np.random.RandomState(3).randn(*D.shape)

    def svm(o):
        np.random.seed(1)
        return linearsvm(D, ELL[:, 0], 0.5, o, engine=eng)
    res["c3_shard_svm_7500x784_us"] = per_iter(svm, 400)
    Dh, sh, _ = gen.huber_problem(0, 65536, 256)
    res["huber_65536x256_us"] = per_iter(lambda o: huberfit(Dh, sh, o, engine=eng), 400)
    st, _ = gen.tv_problem(0, 1 << 16)
    res["tv_2^16_us"] = per_iter(lambda o: totalvariation(st, 1.0, o, engine=eng), 400)
    print("GRAPH " + json.dumps(res))


if __name__ == "__main__":
    main()
