"""Row-sharded A = D solvers under torchrun (one process per GPU): parity against the SERIAL oracle
and timing.  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 \
    --master-port 29511 tools/run_sharded.py [--check] [--bench] [--problem svm|huber|lad]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from admm_project_b200 import Engine, huberfit, lad, lasso, lasso_path, linearsvm, linearsvm_onevsall  # noqa: E402
from admm_project_b200 import generators as gen  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="svm")
    ap.add_argument("--rows", type=int, default=6001)
    ap.add_argument("--cols", type=int, default=96)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--fast", default="", help="'' | weak | strong: fast / accelerated ADMM variant")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    ndev = torch.cuda.device_count()
    dev = local % ndev                 # more ranks than GPUs: ranks share a device (mailbox-only transport, parallel.attach_comm)
    torch.cuda.set_device(dev)
    if world > 1:
        if world > ndev:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    eng = Engine(dev)
    out = {"problem": args.problem, "world": world, "rows": args.rows, "cols": args.cols, "devices": min(world, ndev)}
    if args.problem == "lasso":
        # row-sharded setup (admm_b200_setup_lasso_sharded): every rank hands the FULL host matrix to lasso(), which
        # keeps its own rows; the result must be the serial oracle's on every rank
        D, s, lam, _ = gen.lasso_problem(0, args.rows, args.cols)
        opts = {"objevals": 1, "relax": 1.5, "history": 0}
        res = lasso(D, s, lam, opts, engine=eng)
        out["p2p"] = eng.info()["p2p_ready"]
        if args.check and rank == 0:
            ref = __import__("oracle").lasso(D, s, lam, opts)
    elif args.problem == "lassopath":
        D, s, lam, _ = gen.lasso_problem(0, args.rows, args.cols)
        lams = (lam / 0.1) * 10.0 ** (-np.arange(7) / 3.0)
        rb = lasso_path(D, s, lams, {"reltol": 1e-4}, engine=eng)
        oks = []
        if args.check and rank == 0:
            import oracle
            for j in range(7):
                ref = oracle.lasso(D, s, lams[j], {"reltol": 1e-4, "history": 0})
                oks.append(bool(rb["steps"][j] == ref["steps"] and rel(rb["xopt"][:, j], ref["xopt"]) < 1e-9 and
                                rel(rb["zopt"][:, j], ref["zopt"]) < 1e-9 and
                                rel(rb["pnorm"][:ref["steps"], j], ref["pnorm"]) < 1e-9))
        if rank == 0:
            print("SHARDED " + json.dumps({"problem": "lassopath", "world": world, "ok": bool(oks) and all(oks), "cols_ok": oks,
                                           "zopt_len": args.rows, "steps": [int(v) for v in rb["steps"]]}))
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return
    elif args.problem == "svmbatch":
        D, ELL = gen.svm_mnist_like(0, args.rows, args.cols, nclass=3)
        D = D + 1e-3 * np.random.RandomState(1).randn(*D.shape)
        np.random.seed(3)
        # --fast persist: no objective -> the batch runs as the persistent kernel (csrc/persist_batch.cuh)
        outs = linearsvm_onevsall(D, ELL, 0.5, {} if args.fast == "persist" else {"objevals": 1}, engine=eng)
        oks = []
        if args.check and rank == 0:
            import oracle
            np.random.seed(3)
            for k in range(3):
                ref = oracle.linearsvm(D, ELL[:, k], 0.5, {"history": 0} if args.fast == "persist" else {"objevals": 1, "history": 0})
                oks.append(bool(outs[k]["steps"] == ref["steps"] and rel(outs[k]["xopt"], ref["xopt"]) < 1e-9 and
                                rel(outs[k]["zopt"], ref["zopt"]) < 1e-9 and rel(outs[k]["uopt"], ref["uopt"]) < 1e-9))
        if rank == 0:
            print("SHARDED " + json.dumps({"problem": "svmbatch", "world": world, "ok": bool(oks) and all(oks), "cols_ok": oks,
                                           "zopt_len": int(outs[0]["zopt"].shape[0]), "p2p": eng.info()["p2p_ready"]}))
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return
    elif args.problem == "svm":
        D, ell = gen.svm_mnist_like(0, args.rows, args.cols, nclass=3)
        D = D + 1e-3 * np.random.RandomState(1).randn(*D.shape)
        aux = ell[:, 1]
        opts = {"objevals": 1, "history": 0}
        if args.fast == "persist":       # no objective: the loop runs as the persistent cooperative kernel (csrc/persist.cuh)
            opts = {"history": 0}
        elif args.fast:
            opts.update(fast=1, fasttype=args.fast, maxiters=30)
        np.random.seed(3)
        res = linearsvm(D, aux, 0.5, opts, engine=eng)
        out["launches"] = eng.launch_count()
        if args.check and rank == 0:
            np.random.seed(3)
            ref = __import__("oracle").linearsvm(D, aux, 0.5, opts)
    else:
        D, s, _ = (gen.huber_problem if args.problem == "huber" else gen.lad_problem)(0, args.rows, args.cols)
        opts = {"objevals": 1, "convtest": 1, "history": 0, "relax": 1.5}
        if args.fast:
            opts.update(fast=1, fasttype=args.fast, maxiters=30, restart=0.9, relax=1.0)
        fn = huberfit if args.problem == "huber" else lad
        res = fn(D, s, opts, engine=eng)
        if args.check and rank == 0:
            import oracle
            ref = (oracle.huberfit if args.problem == "huber" else oracle.lad)(D, s, opts)
    out["steps"] = res["steps"]
    out["loop_ms"] = res["engine"]["loop_ms"]
    out["zopt_len"] = int(res["zopt"].shape[0]) if args.problem != "lasso" else args.rows
    if args.check and rank == 0:
        out["ref_steps"] = ref["steps"]
        out["err_x"], out["err_z"], out["err_u"] = rel(res["xopt"], ref["xopt"]), rel(res["zopt"], ref["zopt"]), rel(res["uopt"], ref["uopt"])
        out["err_pnorm"] = rel(res["pnorm"], ref["pnorm"]) if len(ref["pnorm"]) else 0.0
        out["err_obj"] = rel(res["objevals"], ref["objevals"]) if "objevals" in ref else 0.0
        out["ok"] = bool(res["steps"] == ref["steps"] and max(out["err_x"], out["err_z"], out["err_u"], out["err_pnorm"], out["err_obj"]) < 1e-9)
    if rank == 0:
        print("SHARDED " + json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
