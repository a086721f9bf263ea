"""Row-sharded solvers with 2 ranks: the sharded run must reproduce the SERIAL oracle -- same steps, iterates
within 1e-9.  On a box with 2 GPUs: one rank per GPU (NCCL for the Gram, peer mailboxes over
NVLink per iteration).  On a 1-GPU box the two ranks SHARE the device through the mailbox-only transport
(admm_b200_comm_ipc_export / _attach): same kernels, same in-kernel exchange, the driver time-slices the two
processes -- slow, but it is the whole sharded data path."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _device_is_shareable():
    """False only when nvidia-smi positively reports an Exclusive / Prohibited compute mode (two ranks need two contexts)."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=compute_mode", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return True
    return not ("Exclusive" in out or "Prohibited" in out)


@pytest.mark.parametrize("problem,fast", [("svm", ""), ("huber", ""), ("lad", ""), ("huber", "weak"), ("lad", "strong"),
                                          ("huber", "onepass"), ("svm", "onepass"), ("lasso", ""), ("lassopath", ""),
                                          ("svmbatch", ""), ("svm", "persist"), ("svmbatch", "persist"), ("lasso", "wide")])
def test_two_rank_run_matches_serial_oracle(problem, fast):
    import torch
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    if torch.cuda.device_count() < 2 and not _device_is_shareable():
        pytest.skip("one GPU in an exclusive compute mode: two processes cannot share it")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "run_sharded.py"), "--check",
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
    if not line and torch.cuda.device_count() < 2:
        # ranks sharing ONE device are time-sliced by the driver; a bounded mailbox wait (2 s, p2p.cuh) can in principle
        # expire while the peer's process is descheduled on a loaded box: one more try, and say so
        print("shared-device run produced no result, retrying once:", p.stderr[-400:])
        cmd[cmd.index("--master-port") + 1] = str(_free_port())
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("SHARDED ")]
    assert line, p.stdout[-2000:] + p.stderr[-2000:]
    out = json.loads(line[-1][8:])
    assert out["ok"], out
    assert out["zopt_len"] == {"persist": 20001, "wide": 3001}.get(fast, 5001)


def test_mailbox_only_transport_argument_and_state_checks():
    """admm_b200_comm_ipc_export / _attach fail loudly on misuse and leave the handle usable for single-GPU work."""
    import numpy as np
    from admm_project_b200 import Engine, lasso
    from admm_project_b200._lib import EngineError
    eng = Engine(0)
    try:
        with pytest.raises(EngineError, match="comm_ipc_attach"):
            eng.comm_ipc_attach([b"\0" * 64])                        # nothing exported yet
        with pytest.raises(EngineError, match="bad arguments"):
            eng.comm_ipc_export(0, 9)                                 # more ranks than mailboxes (8)
        with pytest.raises(EngineError, match="bad arguments"):
            eng.comm_ipc_export(2, 2)
        hd = eng.comm_ipc_export(0, 2)
        assert len(hd) == 64 and any(hd)
        with pytest.raises(EngineError, match="could not be mapped"):
            eng.comm_ipc_attach([hd, b"\0" * 64])                     # a peer handle that is not one
        assert (eng.rank, eng.nranks) == (0, 1)                       # detached again, on both sides of the ABI
        rs = np.random.RandomState(0)
        D, s = rs.randn(40, 8), rs.randn(40)
        assert lasso(D, s, 0.1, {}, engine=eng)["steps"] > 0
    finally:
        eng.close()
