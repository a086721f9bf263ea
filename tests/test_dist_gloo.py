"""CPU, world_size 2, gloo: the host-side logic of the row-sharded path -- the reference's balanced
partition (errorcheck.m:249-259), shard gathering, the NCCL-id bootstrap payload, and the algebra the
device path relies on: one summed message [D_g'r ; D_g'dz ; D_g'u ; scalars] per iteration
reproduces the serial iteration (unwrappedadmm.m:96-141)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        import oracle
        from admm_project_b200 import generators as gen
        from admm_project_b200.parallel import dist_info, gather_rows, row_range
        assert dist_info() == (rank, world)
        m, n = 1001, 12
        lo, hi = row_range(m, rank, world)
        sizes = oracle.slicemaker(0, world, m)
        assert hi - lo == sizes[rank] and lo == sum(sizes[:rank])
        # shard gathering restores the reference's full-length vectors
        v = np.arange(m, dtype=np.float64)
        assert np.array_equal(gather_rows(v[lo:hi], m), v)
        H = np.arange(3 * m, dtype=np.float64).reshape(m, 3)
        assert np.array_equal(gather_rows(H[lo:hi], m), H)
        # bootstrap payload travels like the 128-byte NCCL id does
        payload = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(payload, src=0)
        assert payload[0] == bytes(range(128))
        # the sharded iteration (what the device does) equals the serial oracle iteration
        D, s, _ = gen.lad_problem(0, m, n)
        ref = oracle.lad(D, s, {"history": 1, "maxiters": 15, "domaxiters": 1})
        Dg, sg = D[lo:hi], s[lo:hi]
        W = torch.from_numpy(Dg.T @ Dg)
        dist.all_reduce(W)                                         # setup: allreduce of the Gram
        R = np.linalg.cholesky(W.numpy())
        z, u, rho = np.zeros(hi - lo), np.zeros(hi - lo), 1.0
        msg = torch.from_numpy(np.concatenate([Dg.T @ (sg + z - u), np.zeros(1)]))
        dist.all_reduce(msg)
        for it in range(15):
            d = msg.numpy()[:n]
            x = np.linalg.solve(R.T, np.linalg.solve(R, d))
            Ax = Dg @ x
            znew = oracle.zminSoftThresholding(Ax + u - sg, 1 / rho)
            u = u + (Ax - znew - sg)
            z = znew
            msg = torch.from_numpy(np.concatenate([Dg.T @ (sg + z - u), [np.sum((Ax - z - sg) ** 2)]]))
            dist.all_reduce(msg)                                   # ONE message per iteration
            assert np.allclose(x, ref["xvals"][:, it], rtol=1e-9, atol=1e-12)
            assert abs(np.sqrt(msg.numpy()[n]) - ref["pnorm"][it]) <= 1e-9 * ref["pnorm"][it] + 1e-10
        assert np.allclose(gather_rows(z, m), ref["zopt"], rtol=1e-9, atol=1e-12)
        # row-sharded lasso setup (admm_b200_setup_lasso_sharded): ONE allreduce of [D_g'D_g ; D_g's_g], + rho*I once,
        # every rank factors the same matrix -> the serial oracle's factor and Dts (lasso.m:160,168)
        Dl, sl, lam, _ = gen.lasso_problem(1, m, n)
        G = torch.from_numpy(np.concatenate([(Dl[lo:hi].T @ Dl[lo:hi]).reshape(-1), Dl[lo:hi].T @ sl[lo:hi]]))
        dist.all_reduce(G)
        Gm = G.numpy()[:n * n].reshape(n, n) + 1.0 * np.eye(n)
        assert np.allclose(np.linalg.cholesky(Gm), np.linalg.cholesky(Dl.T @ Dl + np.eye(n)), rtol=1e-12, atol=1e-14)
        assert np.allclose(G.numpy()[n * n:], Dl.T @ sl, rtol=1e-12, atol=1e-14)
        # the lambda columns of a regularisation path are split by the same balancing rule, no communication
        clo, chi = row_range(7, rank, world)
        assert (clo, chi) == ((0, 4) if rank == 0 else (4, 7))
        # rank 0 draws the random init of unwrappedadmm.m:87-89 for everybody
        from admm_project_b200.parallel import shared_draw
        np.random.seed(100 + rank)                                 # different streams on purpose
        x0 = shared_draw(lambda: np.random.rand(5))
        both = [None, None]
        dist.all_gather_object(both, x0)
        assert np.array_equal(both[0], both[1])
        # transport choice of attach_comm (recording stand-in for the engine, no GPU): ranks on DIFFERENT devices get the
        # NCCL id rank 0 made; ranks that SHARE a device exchange their mailbox handles in rank order instead
        from admm_project_b200 import parallel

        class FakeEngine:
            def __init__(self, key):
                self.rank, self.nranks, self.key, self.calls = 0, 1, key, []

            def unique_id(self):
                return bytes([7]) * 128

            def comm_init(self, r, w, uid):
                self.calls.append(("nccl", r, w, uid))
                self.rank, self.nranks = r, w

            def comm_ipc_export(self, r, w):
                self.calls.append(("export", r, w))
                self.rank, self.nranks = r, w
                return bytes([r + 1]) * 64

            def comm_ipc_attach(self, handles):
                self.calls.append(("attach", list(handles)))
        real_key = parallel._device_key
        try:
            parallel._device_key = lambda e: e.key
            apart = FakeEngine(("host", "GPU-%d" % rank))
            assert parallel.attach_comm(apart) == (rank, world)
            assert apart.calls == [("nccl", rank, world, bytes([7]) * 128)]
            assert parallel.attach_comm(apart) == (rank, world) and len(apart.calls) == 1      # already attached: no-op
            shared = FakeEngine(("host", "GPU-0"))
            assert parallel.attach_comm(shared) == (rank, world)
            assert shared.calls == [("export", rank, world), ("attach", [bytes([1]) * 64, bytes([2]) * 64])]
            os.environ["ADMM_B200_TRANSPORT"] = "ipc"                                            # forced, one rank per device
            forced = FakeEngine(("host", "GPU-%d" % rank))
            parallel.attach_comm(forced)
            assert [c[0] for c in forced.calls] == ["export", "attach"]
        finally:
            parallel._device_key = real_key
            os.environ.pop("ADMM_B200_TRANSPORT", None)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()[-600:]))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")], got
