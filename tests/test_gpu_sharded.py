"""Row-sharded solvers with 2 ranks: the sharded run must reproduce the SERIAL oracle -- same steps, iterates
within 1e-9.  On a box with 2 GPUs: one rank per GPU (NCCL for the Gram, peer mailboxes over
NVLink per iteration).  On a 1-GPU box the two ranks SHARE the device through the mailbox-only transport
(admm_b200_comm_ipc_export / _attach): same kernels, same in-kernel exchange, the driver time-slices the two
processes -- slow, but it is the whole sharded data path."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("problem,fast", [("svm", ""), ("huber", ""), ("lad", ""), ("huber", "weak"), ("lad", "strong"),
                                          ("huber", "onepass"), ("svm", "onepass"), ("lasso", ""), ("lassopath", ""),
                                          ("svmbatch", ""), ("svm", "persist"), ("svmbatch", "persist"), ("lasso", "wide")])
def test_two_rank_run_matches_serial_oracle(problem, fast):
    import torch
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "run_sharded.py"), "--check",
           "--problem", problem, "--rows", "5001", "--cols", "64"] + (["--fast", fast] if fast else [])
    env = dict(os.environ)
    if fast == "onepass":          # the single-pass tile kernel (csrc/onepass.cuh) on every rank's row block
        cmd = cmd[:-2]
        env["ADMM_B200_FORCE_ONEPASS"] = "1"
    if fast == "wide":             # a 208 x 200 Gram (41600 doubles) is larger than a mailbox slot: NCCL on 2 GPUs, two pieces
        cmd = cmd[:-2]             # through the mailboxes when the ranks share a device
        cmd[cmd.index("--rows") + 1], cmd[cmd.index("--cols") + 1] = "3001", "200"
    if fast == "persist":          # the persistent kernel with its in-kernel mailbox exchange (csrc/persist.cuh)
        cmd[cmd.index("--rows") + 1], cmd[cmd.index("--cols") + 1] = "20001", "160"
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("SHARDED ")]
    if not line and torch.cuda.device_count() < 2 and "timed out" in p.stderr:
        # ranks sharing ONE device are time-sliced by the driver; a bounded mailbox wait (2 s, p2p.cuh) can in principle
        # expire while the peer's process is descheduled on a loaded box: one more try, and say so
        print("shared-device run hit a mailbox timeout, retrying once:", p.stderr[-300:])
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("SHARDED ")]
    assert line, p.stdout[-2000:] + p.stderr[-2000:]
    out = json.loads(line[-1][8:])
    assert out["ok"], out
    assert out["zopt_len"] == {"persist": 20001, "wide": 3001}.get(fast, 5001)
