#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: ADMM iterations/s (and time-to-tolerance)
of lasso 65536 x 8192 FP64 (BASELINE.json configs[1]) through the engine's hot path.

A STEP is one solver call on one batch of synthetic input: the one-time setup of solvers/lasso.m
(Dts = D's, Gram D'D + rho*I, Cholesky, inverse factor) followed by ITERS ADMM iterations of
admm.m (x-update, fused relaxed z-update / soft threshold / u-update / residual norms / stop
bookkeeping; domaxiters = 1 so exactly ITERS run).
  value : ITERS * steps / device time, D and s already resident in HBM.
  e2e   : the same call through the public API admm_project_b200.lasso(D, s, lambda, options) with D
          and s in pinned HOST memory -- H2D of D and s, setup, loop, D2H of the results struct -- all
          inside the timed region.
Extra keys: loop_iters_per_s (steady-state loop only), time_to_tol (reltol 1e-4 from resident D),
setup phases in TFLOP/s, roofline of the dominant kernel (the DMMA Gram) and of the per-iteration
x-update (HBM), and a CPU baseline (oracle restatement of the reference, bounded sample).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
N > 1 (torchrun): the single lasso problem does not shard ("replicas only"): every rank
runs an independent replica (its own lambda of the regularisation path), no data-path collective;
value is the sum over ranks (weak scaling), time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, N_COLS = 65536, 8192          # BASELINE.json configs[1]
ITERS = 200                      # "iters/s also with domaxiters=1, maxiters=200"
RELTOL = 1e-4                    # north_star: "reaching reltol 1e-4"
CPU_SAMPLE_ROWS, CPU_SAMPLE_ITERS = 32768, 50   # ~10 s of CPU work on the box's 16 cores


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [c.strip() for c in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mhz = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# synthetic workload: testers/lassotest.m:109-122 recipe at 65536 x 8192, generated on the device
# --------------------------------------------------------------------------------------------
def make_problem_device(torch, dev, m, n, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    testx = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    testx *= (torch.rand(n, dtype=torch.float64, device=dev, generator=g) < 0.6)
    Dt = torch.randn(n, m, dtype=torch.float64, device=dev, generator=g)   # row-major n x m == column-major m x n
    Dt /= Dt.norm(dim=1, keepdim=True)                                      # unit-norm columns of D
    s = Dt.t() @ testx + math.sqrt(0.001) * torch.randn(m, dtype=torch.float64, device=dev, generator=g)
    lam = 0.1 * float((Dt @ s).abs().max())
    return Dt, s, lam


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from admm_project_b200 import DeviceMatrix, Engine, lasso
    from admm_project_b200 import _lib as L

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    Dt, s, lam_max = make_problem_device(torch, dev, M, N_COLS, seed=0)
    lam = lam_max * 10.0 ** (-(rank % 64) / 21.0)        # replica r solves lambda_r of the path
    eng = Engine(local)
    torch.cuda.synchronize()
    eng.set_stream(stream.cuda_stream)
    Ddev = DeviceMatrix(Dt.data_ptr(), M, N_COLS, M, keepalive=Dt)
    opts = {"rho": 1.0, "relax": 1.0, "abstol": 1e-5, "reltol": RELTOL, "history": 0,
            "domaxiters": 1, "maxiters": ITERS, "check_every": 50}

    def step_resident():
        return lasso(Ddev, s.data_ptr(), lam, opts, engine=eng)

    # ---- value: D resident -------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_resident()
        barrier()
        l0, g0 = eng.launch_count(), eng.graph_replays()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record(stream)
            for _ in range(args.steps):
                r = step_resident()
            e1.record(stream)
            barrier()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count() - l0          # kernels executed (a CUDA-graph replay counts the kernels it holds)
        graph_replays = eng.graph_replays() - g0
        phases = eng.setup_phases()
        loop_ms = r["engine"]["loop_ms"]

        # ---- per-kernel timing inside the same process (CUDA events on the launching stream) ---
        o = eng.default_options()
        eng.set_lambda(lam)

        def timed_raw(which, reps):
            eng.iterate_raw(o, which, 5)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            eng.iterate_raw(o, which, reps)
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps * 1e3       # us
        reps = 2 if args.light else 200
        xupd_us = timed_raw(1, reps)
        prox_us = timed_raw(2, reps)
        iter_us = timed_raw(0, reps)

        # ---- time to tolerance (reltol 1e-4) from resident D ----------------------------------
        tol_opts = dict(opts, domaxiters=0, maxiters=(3 if args.light else 1000), check_every=8)
        tol_wall = []
        for _ in range(1 if args.light else 2):     # the first call with a new maxiters re-allocates the histories
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rt = lasso(Ddev, s.data_ptr(), lam, tol_opts, engine=eng)
            torch.cuda.synchronize()
            tol_wall.append((time.perf_counter() - t0) * 1e3)

        # ---- the 64-lambda batch of configs[1]: x-update as two triangular DMMA GEMMs ---------------
        batch = None
        if not args.light:
            ob = eng.default_options()
            ob.reltol = RELTOL
            ob.domaxiters, ob.maxiters, ob.check_every = 1, 40, 40
            lams = (10.0 * lam_max) * 10.0 ** (-np.arange(64) / 21.0)   # lambda_max = max|D's| down to 1e-3 of it
            eng.solve_lasso_batch(ob, lams, want_history=False)               # warm-up
            rb = eng.solve_lasso_batch(ob, lams, want_history=False)
            us = rb["loop_ms"] / 40 * 1e3
            ob.domaxiters, ob.maxiters, ob.check_every = 0, 1000, 8
            rt64 = eng.solve_lasso_batch(ob, lams, want_history=False)
            batch = {"nb": 64, "us_per_iter": us, "rhs_iters_per_s": 64 * 1e6 / us,
                     "tflops_triangular": 2.0 * N_COLS * N_COLS * 64 / (us * 1e-6) / 1e12,
                     "to_tol_ms": rt64["loop_ms"], "steps_min": int(rt64["steps"].min()), "steps_max": int(rt64["steps"].max())}

    # max over ranks
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    value = ITERS * args.steps * world / (ms_max / 1e3)

    # ---- e2e: host buffers through the public API -------------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            Dh = torch.empty((N_COLS, M), dtype=torch.float64, pin_memory=True)
        except RuntimeError:            # N replicas x 4.3 GB of pinned host memory may not be available
            Dh = torch.empty((N_COLS, M), dtype=torch.float64)
        Dh.copy_(Dt)
        sh = torch.empty(M, dtype=torch.float64, pin_memory=Dh.is_pinned())
        sh.copy_(s)
        D_np = Dh.numpy().T                 # (M, N_COLS) Fortran-ordered view of the pinned buffer
        s_np = sh.numpy()
        eng.set_stream(None)
        lasso(D_np, s_np, lam, opts, engine=eng)          # warm-up (allocations)
        barrier()
        e2e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            re = lasso(D_np, s_np, lam, opts, engine=eng)
        eng.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        wmax = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
        e2e = {"value": ITERS * e2e_steps * world / float(wmax.item()), "unit": "iters/s",
               "h2d_bytes_per_step": int(M * N_COLS * 8 + M * 8),
               "d2h_bytes_per_step": int(3 * N_COLS * 8 + 4 * ITERS * 8),
               "steps": e2e_steps, "ms_per_step": float(wmax.item()) / e2e_steps * 1e3,
               "api": "admm_project_b200.lasso(D_host, s_host, lambda, options)", "host_memory": "pinned" if Dh.is_pinned() else "pageable"}
        assert re["steps"] == ITERS

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = measured_peaks()
    # FP64 tensor peak is not in MEASURED_PEAKS.json: measure cuBLAS DGEMM 8192^3 here
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(1 if args.light else 5):
        torch.matmul(a, b)
    c1.record()
    torch.cuda.synchronize()
    fp64_peak = 2 * 8192 ** 3 * (1 if args.light else 5) / (c0.elapsed_time(c1) / 1e3) / 1e12
    del a, b

    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    gram_flops = M * N_COLS * (N_COLS + 1)               # symmetric half, SURVEY.md section 8d
    gram_tflops = gram_flops / (phases["gram_ms"] / 1e3) / 1e12
    tri_bytes = N_COLS * (N_COLS + 1) * 8                # two reads of one triangle per iteration
    xupd_gbs = tri_bytes / (xupd_us * 1e-6) / 1e9
    out = {
        "metric": "admm_iters_per_s", "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "lasso_65536x8192_fp64 (BASELINE.json configs[1]); step = setup (D's, Gram+rho*I, "
                               "Cholesky, inverse factor) + %d ADMM iterations" % ITERS,
                   "rows": M, "cols": N_COLS, "iters_per_step": ITERS, "rho": 1.0, "relax": 1.0,
                   "lambda": "0.1*max|D's|", "l2": "inputs larger than L2 (D 4.3 GB, factor 2 x 0.27 GB per iteration)",
                   "parallelism": "replicas only" if world > 1 else "single GPU"},
        "e2e": e2e, "gpu_launches": int(launches), "graph_replays": int(graph_replays),
        "clocks": clk.summary(),
        "loop_iters_per_s": 1e6 / iter_us,
        "loop_us_per_iter": {"iteration": iter_us, "x_update": xupd_us, "fused_prox": prox_us},
        "setup_ms": phases,
        "lambda_batch": batch,
        "time_to_tol": {"reltol": RELTOL, "steps": int(rt["steps"]), "setup_ms": rt["engine"]["setup_ms"],
                        "loop_ms": rt["engine"]["loop_ms"], "wall_ms": tol_wall[-1], "first_call_wall_ms": tol_wall[0]},
        "roofline": {"kernel": "gemm_f64_dmma_kernel<T,N> (Gram D'D + rho*I, lower tiles)", "bound": "tensor",
                     "achieved": gram_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": gram_tflops / fp64_peak,
                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                     "algorithmic_flops": gram_flops,
                     "traffic": (traffic or {}).get("gram_dram_bytes")},
        "roofline_iter": {"kernel": "coldot_kernel<1> x2 (x = W'(W y): L\\ and L'\\ as products with the cached inverse factor)", "bound": "hbm",
                          "achieved": xupd_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": xupd_gbs / hbm_peak,
                          "frac_of_8TBs_nominal": xupd_gbs / 8000.0, "peak_source": peak_src,
                          "algorithmic_bytes": tri_bytes, "traffic": (traffic or {}).get("xupdate_dram_bytes")},
    }
    if not args.no_cpu and world == 1:          # rank 0 at N = 1 only
        out["cpu_baseline"] = cpu_baseline(CPU_SAMPLE_ROWS, CPU_SAMPLE_ITERS)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# CPU legs: the oracle (NumPy/SciPy restatement of admm.m / lasso.m; no MATLAB/Octave in the image)
# --------------------------------------------------------------------------------------------
def cpu_problem(rows, cols, seed=0):
    import numpy as np
    rs = np.random.RandomState(seed)
    testx = rs.randn(cols) * (rs.rand(cols) < 0.6)
    D = np.asfortranarray(rs.randn(rows, cols))
    D /= np.sqrt(np.einsum("ij,ij->j", D, D))[None, :]
    s = D @ testx + math.sqrt(0.001) * rs.randn(rows)
    lam = 0.1 * float(np.max(np.abs(D.T @ s)))
    return D, s, lam


def cpu_call(D, s, lam, iters):
    import oracle
    t0 = time.perf_counter()
    r = oracle.lasso(D, s, lam, {"rho": 1.0, "reltol": RELTOL, "history": 0, "domaxiters": 1, "maxiters": iters})
    dt = time.perf_counter() - t0
    assert r["steps"] == iters
    return dt, r


def cpu_baseline(rows, iters):
    cores = len(os.sched_getaffinity(0))
    D, s, lam = cpu_problem(rows, N_COLS)
    dt, r = cpu_call(D, s, lam, iters)
    return {"value": iters / dt, "unit": "iters/s", "cores": cores, "kind": "port",
            "sample": "oracle.lasso (NumPy/SciPy restatement of lasso.m + admm.m, OpenBLAS) on %d of %d rows x %d "
                      "cols, setup + %d iterations, one call" % (rows, M, N_COLS, iters),
            "seconds": dt, "loop_s_per_iter": r["runtime"] / iters, "setup_s": dt - r["runtime"],
            # the same call at the full workload (setup scales with the rows, the loop does not), for orientation
            "extrapolated_full_workload_iters_per_s": ITERS / ((dt - r["runtime"]) * M / rows + r["runtime"] / iters * ITERS)}


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use every host thread it can.
    Must run before NumPy / OpenBLAS are loaded (env) and again afterwards (threadpoolctl) to be sure."""
    cores = len(os.sched_getaffinity(0))
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(cores)
    return cores


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    world = env_int("WORLD_SIZE", 1)
    cores = use_all_cores()
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    rows, iters = CPU_SAMPLE_ROWS, CPU_SAMPLE_ITERS
    D, s, lam = cpu_problem(rows, N_COLS)
    for _ in range(min(args.warmup, 1)):
        cpu_call(D, s, lam, 2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_call(D, s, lam, iters)
    dt = time.perf_counter() - t0
    value = iters * args.steps / dt
    sample = ("oracle.lasso (CPU restatement of lasso.m + admm.m; the reference is MATLAB and neither MATLAB nor "
              "Octave exists in the image) on %d of %d rows x %d cols, step = setup + %d iterations" %
              (rows, M, N_COLS, iters))
    print(json.dumps({
        "impl": "reference", "metric": "admm_iters_per_s", "value": value, "unit": "iters/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "lasso_65536x8192_fp64 (BASELINE.json configs[1]), bounded sample", "rows": rows,
                   "cols": N_COLS, "iters_per_step": iters},
        "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    ap.add_argument("--light", action="store_true", help="timed steps only (the command profiled under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
