"""Per-iteration time of the 64-lambda lasso batch at C2 size (two triangular DMMA GEMMs + fused prox)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_project_b200 import DeviceMatrix, Engine  # noqa: E402
m, n, nb = 65536, 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
Dt = torch.randn(n, m, dtype=torch.float64, device=dev)
Dt /= Dt.norm(dim=1, keepdim=True)
s = torch.randn(m, dtype=torch.float64, device=dev)
lam_max = float((Dt @ s).abs().max())
torch.cuda.synchronize()
eng = Engine(0)
eng.setup_lasso(DeviceMatrix(Dt.data_ptr(), m, n, m, keepalive=Dt), s.data_ptr(), 1.0)
o = eng.default_options()
o.domaxiters, o.maxiters, o.check_every = 1, 40, 40
lams = lam_max * 10.0 ** (-np.arange(nb) / 21.0)
eng.solve_lasso_batch(o, lams, want_history=False)
r = eng.solve_lasso_batch(o, lams, want_history=False)
us = r["loop_ms"] / 40 * 1e3
print("BATCH " + json.dumps({"nb": nb, "us_per_iter": us, "rhs_iters_per_s": nb * 1e6 / us,
                             "tflops_triangular": 2.0 * n * n * nb / (us * 1e-6) / 1e12}))
