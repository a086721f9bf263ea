"""Per-iteration timing of the row-sharded A = D solvers at BASELINE.json config sizes
(C3: svm 60000 x 784, C4: huber / lad 4194304 x 1024), data generated on the device.
Single GPU: python tools/bench_unwrapped.py --problem svm
N GPUs:     python -m torch.distributed.run --nproc-per-node N ... tools/bench_unwrapped.py --problem svm"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_project_b200 import DeviceMatrix, Engine, _lib as L  # noqa: E402
from admm_project_b200.parallel import attach_comm, row_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="svm")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--dual", type=int, default=-1, help="1: with dual residual (3-vector pass), 0: nodualerror")
    ap.add_argument("--batch", type=int, default=0, help="one-vs-all class batch of this many label columns (svm)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = a.rows or (60000 if a.problem == "svm" else 4194304)
    n = a.cols or (784 if a.problem == "svm" else 1024)
    lo, hi = row_range(m, rank, world)
    ml = hi - lo
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    ld = ml + (ml & 1)
    Dt = torch.zeros(n, ld, dtype=torch.float64, device=dev)           # row-major n x ld == column-major ld x n
    if a.problem == "svm":   # MNIST-like: U(0,1) with 81% zeros
        Dt[:, :ml] = torch.rand(n, ml, dtype=torch.float64, device=dev, generator=g) * \
            (torch.rand(n, ml, dtype=torch.float64, device=dev, generator=g) < 0.19)
        aux = torch.where(torch.rand(ml, dtype=torch.float64, device=dev, generator=g) < 0.1, 1.0, -1.0).to(torch.float64)
        kind, nodual = L.SVM_HINGE, 1
    else:
        Dt[:, :ml] = torch.randn(n, ml, dtype=torch.float64, device=dev, generator=g)
        xt = torch.randn(n, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        aux = Dt[:, :ml].t() @ xt + 0.1 * torch.randn(ml, dtype=torch.float64, device=dev, generator=g)
        kind, nodual = (L.HUBERFIT if a.problem == "huber" else L.LAD), 0
    if a.dual >= 0:
        nodual = 1 - a.dual
    torch.cuda.synchronize()
    eng = Engine(local)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    attach_comm(eng)
    D = DeviceMatrix(Dt.data_ptr(), ml, n, ld, keepalive=Dt)
    eng.setup_unwrapped(kind, D, aux.data_ptr(), 0.5, m_total=m)
    ph = eng.setup_phases()
    o = eng.default_options()
    o.nodualerror = nodual
    o.history = 0
    out = {"problem": a.problem, "world": world, "rows": m, "cols": n, "setup_ms": ph}
    with torch.cuda.stream(stream):
        for which, name in ((0, "iter"), (1, "xsolve")):
            eng.iterate_raw(o, which, 10)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.iterate_raw(o, which, a.iters)
            e1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / a.iters * 1e3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name + "_us"] = float(t.item())
    # the whole loop through admm_b200_solve: eager bursts vs CUDA-graph bursts (options.graph), 400 iterations
    for gflag, name in ((0, "solve_eager_us"), (1, "solve_graph_us")):
        og = eng.default_options()
        og.nodualerror, og.history, og.domaxiters, og.maxiters, og.check_every, og.graph = nodual, 0, 1, 400, 50, gflag
        best = 1e30
        with torch.cuda.stream(stream):
            for _ in range(3):
                if world > 1:
                    dist.barrier()
                r = eng.solve(og, want_history=False)
                best = min(best, r["loop_ms"] * 1e3 / max(r["steps"], 1))
        t = torch.tensor([best], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = float(t.item())
    if a.batch:
        nbc = a.batch
        lab = torch.where(torch.rand(nbc, ld, dtype=torch.float64, device=dev, generator=g) < 0.1, 1.0, -1.0).to(torch.float64)
        ob = eng.default_options()
        ob.nodualerror, ob.domaxiters, ob.maxiters, ob.check_every, ob.stopcond = 1, 1, 200, 50, 2
        # column k of the m_local x nb label matrix = row k of `lab` (column stride ld)
        import ctypes as C
        res_keep = []
        def run_batch():
            n_, _, m_ = eng.dims()
            steps = np.zeros(nbc, dtype=np.int64); status = np.zeros(nbc, dtype=np.int32); X = np.zeros((n_, nbc), order="F")
            ms = C.c_double()
            L.check(eng._lib.admm_b200_solve_unwrapped_batch(eng._h, C.byref(ob), nbc, C.c_void_p(lab.data_ptr()), ld, None, None, None,
                                                             L.ptr(steps), L.ptr(status), L.ptr(X), None, None, None, None, None, C.byref(ms)))
            return ms.value
        with torch.cuda.stream(stream):
            run_batch()
            if world > 1:
                dist.barrier()
            tb = run_batch()
        t = torch.tensor([tb / 200 * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["batch"] = {"classes": nbc, "us_per_iter_all_classes": float(t.item()),
                        "class_iters_per_s": nbc * 1e6 / float(t.item()),
                        "speedup_vs_sequential": nbc * out["iter_us"] / float(t.item())}
    passes = 2
    out["bytes_per_iter_total"] = passes * m * n * 8
    out["GBs_total"] = out["bytes_per_iter_total"] / (out["iter_us"] * 1e-6) / 1e9
    out["GBs_per_gpu"] = out["GBs_total"] / world
    out["iters_per_s"] = 1e6 / out["iter_us"]
    if rank == 0:
        print("UNWRAPPED " + json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
