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
          inside the timed region; e2e_pageable is the same from ordinary (pageable) host memory.
Extra keys: loop_iters_per_s (steady-state loop only), time_to_tol (reltol 1e-4 from resident D),
setup phases, roofline of the dominant kernel (the DMMA Gram) and of the per-iteration x-update (HBM), the
64-lambda batch, the row-sharded linear SVM of configs[2] (svm_c3) and a CPU baseline (oracle restatement of
the reference on the SAME workload).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
N > 1 (torchrun): STRONG scaling of the same problem.  D is sharded by rows (errorcheck.m:249-259), every rank
forms D_g'D_g and D_g's_g, ONE allreduce sums the 8192 x 8192 Gram (the transpose reduction of
unwrappedadmm.m:114-122), every rank factors it and runs the n-sized iterations; the 64 lambda columns are
split over the ranks; svm_c3 is the row-sharded SVM iteration with its one peer-memory allreduce per iteration.
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
REF_MAX_STEPS = 2                # CPU arm: full workload per step (~10-15 s on 16 cores), at most this many timed steps
SVM_M, SVM_N, SVM_CLASSES = 60000, 784, 10       # BASELINE.json configs[2]


def workload_config():
    """The `config` object of BOTH arms (identical, so the driver's ratio compares like with like)."""
    return {"workload": "lasso_65536x8192_fp64 (BASELINE.json configs[1]); step = setup (D's, Gram+rho*I, "
                        "Cholesky) + %d ADMM iterations" % ITERS,
            "rows": M, "cols": N_COLS, "iters_per_step": ITERS, "rho": 1.0, "relax": 1.0,
            "lambda": "0.1*max|D's|", "reltol": RELTOL,
            "l2": "inputs larger than L2 (D 4.3 GB, factor 0.27-0.54 GB per iteration)"}


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

    def __init__(self, index, enabled=True):
        self.index, self.rows, self.stop, self.t, self.enabled = index, [], threading.Event(), None, enabled

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
        if self.enabled:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.t is not None:
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

    from admm_project_b200 import DeviceMatrix, Engine, RowShard, lasso
    from admm_project_b200 import _lib as L
    from admm_project_b200.parallel import attach_comm, row_range

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

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.only_svm:                   # development aid: just the configs[2] leg (short multi-GPU calls)
        eng = Engine(local)
        torch.cuda.synchronize()
        eng.set_stream(stream.cuda_stream)
        if world > 1:
            attach_comm(eng)
        svm = svm_leg(torch, dist, np, eng, stream, dev, world, rank, barrier, max_over_ranks)
        weak = svm_leg(torch, dist, np, eng, stream, dev, world, rank, barrier, max_over_ranks, SVM_M * world) if world > 1 else None
        if rank == 0:
            svm["n_gpus"], svm["p2p_mailboxes"] = world, bool(eng.info()["p2p_ready"])
            print(json.dumps({"svm_c3": svm, "svm_c3_weak": weak}))
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # The SAME problem at every N (same seed on every rank); rank r keeps rows [lo, hi) of it.
    Dt, s, lam = make_problem_device(torch, dev, M, N_COLS, seed=0)
    lam_max = lam / 0.1
    lo, hi = row_range(M, rank, world)
    ml = hi - lo
    if world > 1:
        ld = ml + (ml & 1)
        Dsh = torch.zeros(N_COLS, ld, dtype=torch.float64, device=dev)
        Dsh[:, :ml] = Dt[:, lo:hi]
        ssh = s[lo:hi].contiguous()
        del Dt, s
        torch.cuda.empty_cache()
    else:
        ld, Dsh, ssh = M, Dt, s
    eng = Engine(local)
    torch.cuda.synchronize()
    eng.set_stream(stream.cuda_stream)
    if world > 1:
        attach_comm(eng)           # NCCL communicator + peer mailboxes; its first collectives run here, not in a timed setup
    Ddev = DeviceMatrix(Dsh.data_ptr(), ml, N_COLS, ld, keepalive=Dsh)
    if world > 1:
        Ddev.m_total, Ddev.row_range = M, (lo, hi)
    opts = {"rho": 1.0, "relax": 1.0, "abstol": 1e-5, "reltol": RELTOL, "history": 0,
            "domaxiters": 1, "maxiters": ITERS, "check_every": 50}

    def step_resident():
        return lasso(Ddev, ssh.data_ptr(), lam, opts, engine=eng)

    # ---- value: D resident -------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_resident()
        barrier()
        l0, g0 = eng.launch_count(), eng.graph_replays()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local, enabled=(rank == 0)) as clk:   # one sampler per job: eight 10 Hz nvidia-smi loops starve the ranks' host threads
            e0.record(stream)
            for _ in range(args.steps):
                r = step_resident()
            e1.record(stream)
            barrier()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count() - l0          # kernels executed (a CUDA-graph replay counts the kernels it holds)
        graph_replays = eng.graph_replays() - g0
        phases = eng.setup_phases()
        assert r["steps"] == ITERS

        # ---- per-kernel timing inside the same process (CUDA events on the launching stream) ---
        o = eng.default_options()
        eng.set_lambda(lam)

        def timed_raw(which, reps, oo=None):
            oo = o if oo is None else oo
            eng.iterate_raw(oo, which, 5)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            eng.iterate_raw(oo, which, reps)
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps * 1e3       # us
        reps = 2 if args.light else 200
        xupd_us = max_over_ranks(timed_raw(1, reps))
        prox_us = max_over_ranks(timed_raw(2, reps))
        iter_us = max_over_ranks(timed_raw(0, reps))
        # the reference's own formulation, x = U \ (L \ y) by blocked substitution on L (options.xsolve = SUBST): what runs
        # when the conditioning guard fires; 16 dependent 512-wide steps per solve at n = 8192
        osub = eng.default_options()
        osub.xsolve = L.XSOLVE_SUBST
        xsub_us = max_over_ranks(timed_raw(1, max(2, reps // 10), osub))

        # ---- time to tolerance (reltol 1e-4) from resident D ----------------------------------
        tol_opts = dict(opts, domaxiters=0, maxiters=(3 if args.light else 1000), check_every=8)
        tol_wall = []
        for _ in range(1 if args.light else 2):     # the first call with a new maxiters re-allocates the histories
            barrier()
            t0 = time.perf_counter()
            rt = lasso(Ddev, ssh.data_ptr(), lam, tol_opts, engine=eng)
            torch.cuda.synchronize()
            tol_wall.append(max_over_ranks((time.perf_counter() - t0) * 1e3))

        # ---- the 64-lambda batch of configs[1]: x-update as two triangular DMMA GEMMs; the lambda columns are
        # split over the ranks (no communication) -------------------------------------------------------------
        batch = None
        if not args.light:
            ob = eng.default_options()
            ob.reltol = RELTOL
            ob.domaxiters, ob.maxiters, ob.check_every = 1, 40, 40
            lams = lam_max * 10.0 ** (-np.arange(64) / 21.0)              # lambda_max = max|D's| down to 1e-3 of it
            clo, chi = row_range(64, rank, world)
            mine = lams[clo:chi]
            eng.solve_lasso_batch(ob, mine, want_history=False)               # warm-up
            barrier()
            rb = eng.solve_lasso_batch(ob, mine, want_history=False)
            us = max_over_ranks(rb["loop_ms"] / 40 * 1e3)
            ob.domaxiters, ob.maxiters, ob.check_every = 0, 1000, 8
            rt64 = eng.solve_lasso_batch(ob, mine, want_history=False)
            batch = {"nb": 64, "columns_per_rank": int(chi - clo), "us_per_iter": us, "rhs_iters_per_s": 64 * 1e6 / us,
                     "tflops_triangular": 2.0 * N_COLS * N_COLS * 64 / (us * 1e-6) / 1e12,
                     "to_tol_ms": max_over_ranks(rt64["loop_ms"]), "steps_min_rank0": int(rt64["steps"].min()),
                     "steps_max_rank0": int(rt64["steps"].max())}

    ms_max = max_over_ranks(ms)
    value = ITERS * args.steps / (ms_max / 1e3)          # the ONE problem, however many GPUs share it (strong scaling)

    # ---- e2e: host buffers through the public API -------------------------------------------------
    def e2e_leg(pinned):
        Dh = torch.empty((N_COLS, ml), dtype=torch.float64, pin_memory=pinned)
        Dh.copy_(Dsh[:, :ml])
        sh = torch.empty(ml, dtype=torch.float64, pin_memory=pinned)
        sh.copy_(ssh)
        D_np = Dh.numpy().T                 # (ml, N_COLS) column-major view of the host buffer
        s_np = sh.numpy()
        D_arg = RowShard(D_np, M, (lo, hi)) if world > 1 else D_np
        eng.set_stream(None)
        lasso(D_arg, s_np, lam, opts, engine=eng)          # warm-up (allocations)
        barrier()
        e2e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            re = lasso(D_arg, s_np, lam, opts, engine=eng)
        eng.synchronize()
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        assert re["steps"] == ITERS
        return {"value": ITERS * e2e_steps / wall, "unit": "iters/s",
                "h2d_bytes_per_step": int(M * N_COLS * 8 + M * 8),            # all ranks together
                "d2h_bytes_per_step": int(world * (3 * N_COLS * 8 + 4 * ITERS * 8)),
                "steps": e2e_steps, "ms_per_step": wall / e2e_steps * 1e3,
                "api": "admm_project_b200.lasso(D_host, s_host, lambda, options)" if world == 1 else
                       "admm_project_b200.lasso(RowShard(D_host_rows, m_total), s_host_rows, lambda, options) on every rank",
                "host_memory": "pinned" if pinned else "pageable"}
    e2e = e2e_pageable = None
    if not args.no_e2e:
        try:
            e2e = e2e_leg(True)
        except RuntimeError as ex:          # pinned host memory may not be available
            e2e = {"unavailable": "pinned host allocation failed: %s" % str(ex)[:80]}
        if not args.light:
            e2e_pageable = e2e_leg(False)
        eng.set_stream(stream.cuda_stream)

    # ---- configs[2]: linear SVM by transpose reduction, 60000 x 784, rows sharded over the ranks ------------------
    svm = svm_weak = None
    if not args.light and not args.no_svm:
        svm = svm_leg(torch, dist, np, eng, stream, dev, world, rank, barrier, max_over_ranks)
        if world > 1:          # 60000 rows per GPU: the weak form of the same leg (at N = 1 it is the strong one)
            svm_weak = svm_leg(torch, dist, np, eng, stream, dev, world, rank, barrier, max_over_ranks, SVM_M * world)

    if rank != 0:
        eng.close()
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

    gram_flops = ml * N_COLS * (N_COLS + 1)              # symmetric half of this rank's rows, SURVEY.md section 8d
    gram_tflops = gram_flops / (phases["gram_ms"] / 1e3) / 1e12
    tri_bytes = N_COLS * (N_COLS + 1) * 8                # two reads of one triangle per iteration
    xupd_gbs = tri_bytes / (xupd_us * 1e-6) / 1e9
    info = eng.info()
    # DRAM bytes per launch of the two roofline kernels: counters of the committed ncu --set full captures of THIS command
    # (profiles/traffic.json names the .ncu-rep each came from); not re-measured in-run -- ncu cannot run inside a timed bench
    traffic, traffic_src = {}, None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic_src = "profiles/traffic.json (ncu --set full capture of this command, N = 1, full 65536 rows)"
    except Exception:
        pass
    out = {
        "metric": "admm_iters_per_s", "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "parallelism": "single GPU" if world == 1 else
                       "rows of D sharded over %d ranks; one ncclAllReduce of the 8192x8192 Gram + D's per step; Cholesky and "
                       "the n-sized iterations replicated; lambda batch split by columns" % world,
        "e2e": e2e, "e2e_pageable": e2e_pageable, "gpu_launches": int(launches), "graph_replays": int(graph_replays),
        "clocks": clk.summary(),
        "loop_iters_per_s": 1e6 / iter_us,
        "loop_us_per_iter": {"iteration": iter_us, "x_update": xupd_us, "fused_prox": prox_us,
                             "x_update_substitution": xsub_us},
        "setup_ms": phases,
        "last_step_ms": {"setup": r["engine"]["setup_ms"], "loop": r["engine"]["loop_ms"],
                         "loop_us_per_iter": r["engine"]["loop_ms"] / ITERS * 1e3},     # rank 0, the last timed step
        "lambda_batch": batch,
        "svm_c3": svm,
        "svm_c3_weak": svm_weak,
        "time_to_tol": {"reltol": RELTOL, "steps": int(rt["steps"]), "setup_ms": rt["engine"]["setup_ms"],
                        "loop_ms": rt["engine"]["loop_ms"], "wall_ms": tol_wall[-1], "first_call_wall_ms": tol_wall[0]},
        "roofline": {"kernel": "gemm_tma_kernel (Gram D_g'D_g: FP64 DMMA tiles fed by cp.async.bulk.tensor, lower tiles, this rank's %d rows)" % ml,
                     "bound": "tensor", "achieved": gram_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": gram_tflops / fp64_peak,
                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                     "algorithmic_flops": gram_flops,
                     "traffic": traffic.get("gram_dram_bytes") if world == 1 else None, "traffic_source": traffic_src},
        "roofline_iter": {"kernel": "x-update x = W'(W y) on the cached inverse factor", "bound": "hbm",
                          "achieved": xupd_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": xupd_gbs / hbm_peak,
                          "frac_of_8TBs_nominal": xupd_gbs / 8000.0, "peak_source": peak_src,
                          "note": "algorithmic bytes = the two triangular solves of getProxOps.m:1200 (two reads of one triangle, "
                                  "SURVEY 8d); symtri_kernel reads the triangle ONCE (x = W'(W y) from one pass), so the bytes really "
                                  "moved are half: see real_gbs",
                          "real_bytes": tri_bytes // 2, "real_gbs": xupd_gbs / 2, "real_frac": xupd_gbs / 2 / hbm_peak,
                          "algorithmic_bytes": tri_bytes,
                          "traffic": traffic.get("xupdate_dram_bytes") if world == 1 else None, "traffic_source": traffic_src},
        "comm": {"p2p_mailboxes": bool(info["p2p_ready"])} if world > 1 else None,
    }
    if not args.no_cpu and world == 1 and not args.light:          # rank 0 at N = 1 only
        out["cpu_baseline"] = cpu_baseline()
    print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def svm_leg(torch, dist, np, eng, stream, dev, world, rank, barrier, max_over_ranks, rows_total=SVM_M):
    """BASELINE.json configs[2]: hinge-loss linear SVM via unwrapped ADMM / transpose reduction on synthetic
    MNIST-shaped data (60000 x 784, U(0,1) with 81 % zeros, ten one-vs-all label columns), rows sharded over the
    ranks.  Reports the per-iteration time of one classifier and of the ten-class batch (device time, max over
    ranks); the driver's per-N runs give the strong-scaling efficiency.  rows_total = 60000 * world is the WEAK
    form (60000 rows per GPU, the same per-rank work at every N): per-iteration time should then stay flat."""
    SVM_M = rows_total
    from admm_project_b200 import DeviceMatrix
    from admm_project_b200 import _lib as L
    from admm_project_b200.parallel import row_range
    import ctypes as C
    g = torch.Generator(device=dev)
    g.manual_seed(1234)                                   # same stream on every rank: ONE global problem
    Dt = torch.rand(SVM_N, SVM_M, dtype=torch.float64, device=dev, generator=g) * \
        (torch.rand(SVM_N, SVM_M, dtype=torch.float64, device=dev, generator=g) < 0.19)
    digit = torch.randint(0, SVM_CLASSES, (SVM_M,), device=dev, generator=g)
    lo, hi = row_range(SVM_M, rank, world)
    ml = hi - lo
    ld = ml + (ml & 1)
    Dsh = torch.zeros(SVM_N, ld, dtype=torch.float64, device=dev)
    Dsh[:, :ml] = Dt[:, lo:hi]
    lab = -torch.ones(SVM_CLASSES, ld, dtype=torch.float64, device=dev)      # column k of the ml x 10 label matrix = row k
    lab[digit[lo:hi], torch.arange(ml, device=dev)] = 1.0
    del Dt
    torch.cuda.synchronize()
    D = DeviceMatrix(Dsh.data_ptr(), ml, SVM_N, ld, keepalive=Dsh)
    out = {"rows": SVM_M, "cols": SVM_N, "classes": SVM_CLASSES, "rows_per_rank": ml}
    with torch.cuda.stream(stream):
        eng.setup_unwrapped(L.SVM_HINGE, D, lab.data_ptr(), 0.5, m_total=SVM_M)
        out["setup_ms"] = eng.setup_phases()["total_ms"]
        og = eng.default_options()
        og.nodualerror, og.history, og.domaxiters, og.maxiters, og.check_every, og.stopcond = 1, 0, 1, 400, 50, 2
        best = 1e30
        for _ in range(5):
            barrier()
            r = eng.solve(og, want_history=False)
            best = min(best, max_over_ranks(r["loop_ms"] * 1e3 / max(r["steps"], 1)))
        out["single_us_per_iter"] = best
        out["single_hbm_gbs_algorithmic"] = 2.0 * SVM_M * SVM_N * 8 / (best * 1e-6) / 1e9      # 2 passes over D
        ob = eng.default_options()
        ob.nodualerror, ob.domaxiters, ob.maxiters, ob.check_every, ob.stopcond = 1, 1, 200, 50, 2
        n_, _, m_ = eng.dims()

        def run_batch():
            steps = np.zeros(SVM_CLASSES, dtype=np.int64)
            status = np.zeros(SVM_CLASSES, dtype=np.int32)
            X = np.zeros((n_, SVM_CLASSES), order="F")
            msb = C.c_double()
            L.check(eng._lib.admm_b200_solve_unwrapped_batch(eng._h, C.byref(ob), SVM_CLASSES, C.c_void_p(lab.data_ptr()), ld,
                                                             None, None, None, L.ptr(steps), L.ptr(status), L.ptr(X), None,
                                                             None, None, None, None, C.byref(msb)))
            return msb.value
        run_batch()
        bestb = 1e30
        for _ in range(5):          # 200 iterations each; the two-rank exchange makes single runs bimodal (105 / 115 / 138 us seen at N = 2)
            barrier()
            bestb = min(bestb, max_over_ranks(run_batch() / 200 * 1e3))
        out["batch10_us_per_iter"] = bestb
        out["batch10_class_iters_per_s"] = SVM_CLASSES * 1e6 / bestb
    return out


# --------------------------------------------------------------------------------------------
# CPU legs: the oracle (NumPy/SciPy restatement of admm.m / lasso.m; no MATLAB/Octave in the image) on the SAME
# workload as the GPU arm -- all 65536 rows, ITERS iterations per step
# --------------------------------------------------------------------------------------------
def cpu_problem(rows, cols, seed=0):
    """testers/lassotest.m:109-122 at rows x cols; column blocks are drawn by independent PCG64 streams on a
    thread pool (NumPy releases the GIL while it fills), so the 4.3 GB matrix takes seconds, not half a minute."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    ss = np.random.SeedSequence(seed)
    kids = ss.spawn((cols + 255) // 256 + 1)
    rs = np.random.default_rng(kids[-1])
    testx = rs.standard_normal(cols) * (rs.random(cols) < 0.6)
    D = np.empty((rows, cols), order="F")

    def fill(b):
        j0 = b * 256
        w = min(256, cols - j0)
        blk = np.random.default_rng(kids[b]).standard_normal((w, rows))
        blk /= np.sqrt(np.einsum("ij,ij->i", blk, blk))[:, None]      # unit-norm columns of D
        D[:, j0:j0 + w] = blk.T
    with ThreadPoolExecutor(max_workers=len(os.sched_getaffinity(0))) as ex:
        list(ex.map(fill, range((cols + 255) // 256)))
    s = D @ testx + math.sqrt(0.001) * rs.standard_normal(rows)
    lam = 0.1 * float(np.max(np.abs(D.T @ s)))
    return D, s, lam


def cpu_call(D, s, lam, iters):
    import oracle
    t0 = time.perf_counter()
    r = oracle.lasso(D, s, lam, {"rho": 1.0, "reltol": RELTOL, "history": 0, "domaxiters": 1, "maxiters": iters})
    dt = time.perf_counter() - t0
    assert r["steps"] == iters
    return dt, r


CPU_KIND = ("oracle.lasso: NumPy/SciPy (OpenBLAS) restatement of solvers/lasso.m + admm.m; the reference is MATLAB "
            "and neither MATLAB nor Octave exists in the image")


def cpu_baseline():
    cores = use_all_cores()
    D, s, lam = cpu_problem(M, N_COLS)
    dt, r = cpu_call(D, s, lam, ITERS)
    return {"value": ITERS / dt, "unit": "iters/s", "cores": cores, "kind": "port",
            "sample": "%s; the full workload once: %d x %d, setup + %d iterations" % (CPU_KIND, M, N_COLS, ITERS),
            "seconds": dt, "loop_s_per_iter": r["runtime"] / ITERS, "setup_s": dt - r["runtime"], "same_config": True}


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use every host thread it can.
    Set before NumPy / OpenBLAS are loaded (env) and again afterwards (threadpoolctl) to be sure."""
    cores = len(os.sched_getaffinity(0))
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(cores)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    world = env_int("WORLD_SIZE", 1)
    cores = use_all_cores()
    D, s, lam = cpu_problem(M, N_COLS)
    if args.warmup > 0:
        cpu_call(D[:4096], s[:4096], lam, 2)            # loads BLAS / SciPy, spins the thread pool up; not the workload
    steps = max(1, min(args.steps, REF_MAX_STEPS))      # every timed step IS the full workload; fewer of them than the GPU arm
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_call(D, s, lam, ITERS)
    dt = time.perf_counter() - t0
    value = ITERS * steps / dt
    sample = "%s; %d timed steps of the full workload (%d requested): %d x %d, setup + %d iterations each" % (
        CPU_KIND, steps, args.steps, M, N_COLS, ITERS)
    print(json.dumps({
        "impl": "reference", "metric": "admm_iters_per_s", "value": value, "unit": "iters/s", "n_gpus": world,
        "steps": steps, "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(), "same_config": True,
        "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs")
    ap.add_argument("--no-svm", action="store_true", help="skip the configs[2] SVM leg")
    ap.add_argument("--only-svm", action="store_true", help="only the configs[2] SVM leg")
    ap.add_argument("--light", action="store_true", help="timed steps only (the command profiled under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
