"""Short C2-shaped run for ncu: one setup (65536 x 8192) + a few raw iterations."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_project_b200 import DeviceMatrix, Engine  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = torch.device("cuda:0")
torch.manual_seed(0)
Dt = torch.randn(n, m, dtype=torch.float64, device=dev)
Dt /= Dt.norm(dim=1, keepdim=True)
s = torch.randn(m, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
eng = Engine(0)
eng.setup_lasso(DeviceMatrix(Dt.data_ptr(), m, n, m, keepalive=Dt), s.data_ptr(), 1.0)
eng.set_lambda(0.1)
o = eng.default_options()
eng.iterate_raw(o, 0, iters)
eng.synchronize()
print(eng.setup_phases(), eng.launch_count())
