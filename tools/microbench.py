"""Per-kernel timings on one B200 (CUDA events on the handle's stream).  Not the bench contract --
a development aid whose output goes to gpurun_out/."""
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_project_b200 import DeviceMatrix, Engine  # noqa: E402


STREAM = None


def timed(stream_fn, reps, warm=3):
    # a real (non-legacy) stream: admm_b200_set_stream(NULL) means "the handle's own stream"
    with torch.cuda.stream(STREAM):
        for _ in range(warm):
            stream_fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            stream_fn()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    global STREAM
    STREAM = torch.cuda.Stream()
    out = {"m": m, "n": n}
    # fp64 dgemm peak probe (cuBLAS)
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    t = timed(lambda: torch.matmul(a, b), 5)
    out["cublas_dgemm_8192_tflops"] = 2 * 8192 ** 3 / t / 1e9
    del a, b
    Dt = torch.randn(n, m, dtype=torch.float64, device=dev)        # row-major n x m == column-major m x n
    Dt /= Dt.norm(dim=1, keepdim=True)
    s = torch.randn(m, dtype=torch.float64, device=dev)
    eng = Engine(0)
    torch.cuda.synchronize()
    eng.set_stream(STREAM.cuda_stream)
    D = DeviceMatrix(Dt.data_ptr(), m, n, m, keepalive=Dt)
    t0 = time.perf_counter()
    eng.setup_lasso(D, s.data_ptr(), 1.0)
    out["setup_wall_s"] = time.perf_counter() - t0
    ph = eng.setup_phases()
    out.update(ph)
    k = min(m, n)
    gram_flops = (m * n * (n + 1)) if m >= n else (n * m * (m + 1))
    out["gram_tflops_sym"] = gram_flops / ph["gram_ms"] / 1e9
    out["chol_tflops"] = (k ** 3 / 3) / ph["chol_ms"] / 1e9
    # cuBLAS syrk-equivalent reference for the Gram
    t = timed(lambda: torch.matmul(Dt, Dt.t()), 2, warm=1)
    out["cublas_gram_full_ms"] = t
    o = eng.default_options()
    eng.set_lambda(0.1)
    for which, name in ((0, "iter"), (1, "xupdate"), (2, "prox")):
        t = timed(lambda: eng.iterate_raw(o, which, 20), 5)
        out[name + "_us"] = t / 20 * 1e3
    tri_bytes = k * (k + 1) * 8
    if m >= n:
        out["xupdate_GBs"] = tri_bytes / (out["xupdate_us"] * 1e-6) / 1e9
    else:
        out["xupdate_GBs"] = (tri_bytes + 2 * m * n * 8) / (out["xupdate_us"] * 1e-6) / 1e9
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
