"""Generates tests/golden/*.npz with the ORACLE (oracle/ -- the CPU restatement of the reference).

The reference is MATLAB and cannot run in this image, so these are NOT outputs of the reference
itself (parity stays "unpinned"): they freeze the oracle's answers on small seeded
problems so that (a) an accidental change of the oracle is caught on CPU and (b) the GPU tests have
fixed fixtures that do not depend on the oracle at run time.      python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from admm_project_b200 import generators as gen  # noqa: E402

KEEP = ("steps", "xopt", "zopt", "uopt", "pnorm", "dnorm", "perr", "derr", "objevals")


def pack(res, **inputs):
    out = {k: np.asarray(res[k]) for k in KEEP if k in res}
    out.update({"in_" + k: np.asarray(v) for k, v in inputs.items()})
    return out


def cases():
    D, s, lam, _ = gen.lasso_problem(0, 96, 40)
    yield "lasso_tall_96x40", pack(oracle.lasso(D, s, lam, {"objevals": 1, "relax": 1.5}), D=D, s=s, lam=lam, relax=1.5)
    D, s, lam, _ = gen.lasso_problem(1, 40, 96)
    yield "lasso_fat_40x96", pack(oracle.lasso(D, s, lam, {"objevals": 1}), D=D, s=s, lam=lam, relax=1.0)
    D, ell = gen.svm_problem(0, 48, 48)
    np.random.seed(21)
    yield "svm_hinge_96x2", pack(oracle.linearsvm(D, ell, 0.5, {"objevals": 1}), D=D, ell=ell, C=0.5, seed=21)
    D, s, _ = gen.huber_problem(0, 300, 10)
    yield "huber_300x10", pack(oracle.huberfit(D, s, {"objevals": 1, "convtest": 1}), D=D, s=s)
    D, s, _ = gen.lad_problem(0, 200, 8)
    yield "lad_200x8", pack(oracle.lad(D, s, {"objevals": 1, "convtest": 1, "relax": 1.4}), D=D, s=s, relax=1.4)
    s, _ = gen.tv_problem(0, 300)
    yield "tv_300", pack(oracle.totalvariation(s, 2.0, {"objevals": 1, "maxiters": 2000}), s=s, lam=2.0)
    D, s, _ = gen.bp_problem(0, 24, 60, density=0.08)
    yield "bp_24x60", pack(oracle.basispursuit(D, s, {"objevals": 1, "maxiters": 5000}), D=D, s=s)
    P, Q, r, s, _ = gen.model_problem(0, 60, 24)
    yield "model_60x24", pack(oracle.model(P, Q, r, s, {"objevals": 1, "relax": 1.3}), P=P, Q=Q, r=r, s=s, relax=1.3)


if __name__ == "__main__":
    for name, data in cases():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, int(data["steps"]), {k: v.shape for k, v in data.items() if k.startswith("in_")})
