"""Timing probe of the look-ahead Cholesky (+ inverse factor) at n = 8192: the full algorithm, without the riding
inverse, and the bare critical chain (bulk trailing updates left out -- wrong factor, timing only).  Development aid."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import torch
from admm_project_b200 import DeviceMatrix, Engine
n, m = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
torch.manual_seed(0)
Dt = torch.randn(n, m, dtype=torch.float64, device=dev)
Dt /= Dt.norm(dim=1, keepdim=True)
s = torch.randn(m, dtype=torch.float64, device=dev)
st = torch.cuda.Stream()
eng = Engine(0)
torch.cuda.synchronize()
eng.set_stream(st.cuda_stream)
D = DeviceMatrix(Dt.data_ptr(), m, n, m, keepalive=Dt)
best = None
for _ in range(4):
    try:
        eng.setup_lasso(D, s.data_ptr(), 1.0)
    except Exception as ex:
        print(json.dumps({"error": str(ex)[:100]})); sys.exit(0)
    ph = eng.setup_phases()
    if best is None or ph["chol_ms"] + ph["inverse_ms"] < best["chol_ms"] + best["inverse_ms"]:
        best = ph
print(json.dumps(best))
''' % ROOT


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    for name, env in (("full (default: no SM reserve)", {}), ("look-ahead depth 1", {"ADMM_B200_CHOL_DEPTH1": "1"}),
                      ("bulk in 2 K pieces", {"ADMM_B200_BULK_PIECES": "2"}), ("bulk in 4 K pieces", {"ADMM_B200_BULK_PIECES": "4"}),
                      ("bulk in 2 K pieces, inverse K chunk 512", {"ADMM_B200_BULK_PIECES": "2", "ADMM_B200_INV_KCHUNK": "512"}),
                      ("no riding inverse", {"ADMM_B200_NO_INV_OVERLAP": "1"}),
                      ("chain only (no bulk, no riding inverse)", {"ADMM_B200_NO_INV_OVERLAP": "1", "ADMM_B200_CHOL_PROBE_NOBULK": "1"}),
                      ("chain + riding inverse, no bulk", {"ADMM_B200_CHOL_PROBE_NOBULK": "1"}),
                      ("v1 (two streams)", {"ADMM_B200_CHOL_V1": "1"})):
        e = dict(os.environ, **env)
        p = subprocess.run([sys.executable, "-c", CHILD, str(n), str(m)], capture_output=True, text=True, env=e, timeout=300)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        print("CHOLPROBE n=%d %-42s %s" % (n, name, line[-1] if line else p.stderr[-300:]))


if __name__ == "__main__":
    main()
