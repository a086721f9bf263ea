import sys, time, math
sys.path.insert(0, "/root/repo")
import torch
import bench
from admm_project_b200 import DeviceMatrix, Engine, lasso
dev = torch.device("cuda", 0)
Dt, s, lam = bench.make_problem_device(torch, dev, bench.M, bench.N_COLS, seed=0)
eng = Engine(0)
Ddev = DeviceMatrix(Dt.data_ptr(), bench.M, bench.N_COLS, bench.M, keepalive=Dt)
base = {"rho": 1.0, "relax": 1.0, "abstol": 1e-5, "reltol": 1e-4, "history": 0}
for tag, o in (("iters200", dict(base, domaxiters=1, maxiters=200, check_every=50)),
               ("tol_graph", dict(base, maxiters=1000, check_every=8)),
               ("tol_graph", dict(base, maxiters=1000, check_every=8)),
               ("tol_eager", dict(base, maxiters=1000, check_every=8, graph=0)),
               ("tol_eager", dict(base, maxiters=1000, check_every=8, graph=0)),
               ("tol_graph", dict(base, maxiters=1000, check_every=8))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = lasso(Ddev, s.data_ptr(), lam, o, engine=eng)
    torch.cuda.synchronize()
    print(tag, "wall_ms %.1f" % ((time.perf_counter() - t0) * 1e3), "steps", r["steps"], "setup %.1f loop %.2f" % (r["engine"]["setup_ms"], r["engine"]["loop_ms"]))
