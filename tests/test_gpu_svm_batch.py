"""GPU parity of the one-vs-all class batch (examples/mnistsvm.m:121-156): every class of the batched
run must reproduce the oracle's per-class linearsvm run -- same steps, iterates within 1e-9."""
import numpy as np
import pytest

import oracle
from admm_project_b200 import _lib as L
from admm_project_b200 import generators as gen

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("rows,cols,ncls", [(3000, 48, 10), (1201, 33, 3), (20000, 64, 5)])
def test_one_vs_all_batch_matches_per_class_oracle(engine, rows, cols, ncls):
    D, ell = gen.svm_mnist_like(0, rows, cols, nclass=ncls)
    D = D + 1e-3 * np.random.RandomState(1).randn(*D.shape)
    rs = np.random.RandomState(7)
    X0, Z0, U0 = rs.rand(cols, ncls), rs.rand(rows, ncls), rs.rand(rows, ncls)     # unwrappedadmm.m:87-89
    engine.setup_unwrapped(L.SVM_HINGE, D, ell[:, 0], 0.5)
    o = engine.default_options()
    o.nodualerror, o.stopcond, o.maxiters, o.objevals = 1, L.STOP_BOTH, 1000, 1     # unwrappedadmm.m:90-92
    out = engine.solve_unwrapped_batch(o, ell, X0, Z0, U0)
    for k in range(ncls):
        _, minz, _ = oracle.getproxops("LinearSVM", dict(D=D, Dt=D.T, ell=ell[:, k], C=0.5, lossfunction="hinge",
                                                          slices=[rows]))
        opts = dict(A=D, At=D.T, B=-1, nB=rows, c=0, m=rows, x0=X0[:, k], z0=Z0[:, k], u0=U0[:, k], maxiters=1000,
                    stopcond="both", nodualerror=1, history=0, objevals=1,
                    obj=lambda x, z, e=ell[:, k]: 0.5 * float(x @ x) + 0.5 * float(np.sum(np.maximum(1 - e * (D @ x), 0))))
        W = D.T @ D
        ref = oracle.admm(lambda x, z, u, rho: np.linalg.solve(W, D.T @ (z - u)), lambda x, z, u, rho: minz(x, z, u, rho, 1), opts)
        assert out["steps"][k] == ref["steps"], (k, out["steps"][k], ref["steps"])
        n = ref["steps"]
        assert rel(out["xopt"][:, k], ref["xopt"]) < 1e-9, k
        assert rel(out["zopt"][:, k], ref["zopt"]) < 1e-9 and rel(out["uopt"][:, k], ref["uopt"]) < 1e-9
        assert rel(out["pnorm"][:n, k], ref["pnorm"]) < 1e-9 and rel(out["perr"][:n, k], ref["perr"]) < 1e-9
        assert rel(out["objevals"][:n, k], ref["objevals"]) < 1e-9


def test_linearsvm_onevsall_equals_a_loop_of_linearsvm_calls(engine):
    from admm_project_b200 import linearsvm_onevsall
    D, ell = gen.svm_mnist_like(2, 2500, 40, nclass=4)
    D = D + 1e-3 * np.random.RandomState(3).randn(*D.shape)
    np.random.seed(12)
    refs = [oracle.linearsvm(D, ell[:, k], 0.5, {"objevals": 1, "history": 0}) for k in range(4)]   # mnistsvm.m:136-142
    np.random.seed(12)
    got = linearsvm_onevsall(D, ell, 0.5, {"objevals": 1}, engine=engine)
    for k in range(4):
        assert got[k]["steps"] == refs[k]["steps"]
        for key in ("xopt", "zopt", "uopt", "pnorm", "perr", "objevals"):
            assert rel(got[k][key], refs[k][key]) < 1e-9, (k, key)
        assert abs(got[k]["objopt"] - refs[k]["objopt"]) <= 1e-9 * abs(refs[k]["objopt"])
