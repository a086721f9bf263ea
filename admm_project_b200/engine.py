"""Thin object wrapper over the C-ABI handle (include/admm_b200.h).  One Engine = one GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class DeviceMatrix:
    """A column-major FP64 matrix already resident in HBM (raw device pointer + shape + ld), e.g.
    ``DeviceMatrix(t.data_ptr(), m, n, ld)`` for a torch tensor holding the column-major data.
    The solvers accept it wherever the reference takes the data matrix D."""

    def __init__(self, ptr, rows, cols, ld=None, keepalive=None):
        self.ptr, self.shape, self.ld = int(ptr), (int(rows), int(cols)), int(ld or rows)
        self.keepalive = keepalive


class RowShard:
    """THIS rank's rows of a row-sharded data matrix held in HOST memory (one process per GPU):
    ``RowShard(D_local, m_total)``.  The solvers accept it wherever the reference takes D; a DeviceMatrix with an
    ``m_total`` attribute is the device-resident equivalent."""

    def __init__(self, array, m_total, row_range=None):
        self.array, self.m_total, self.row_range = array, int(m_total), row_range
        self.shape = (int(array.shape[0]), int(array.shape[1]))


class Engine:
    def __init__(self, device=0):
        lib = L.load()
        h = C.c_void_p()
        L.check(lib.admm_b200_create(int(device), C.byref(h)))
        self._lib, self._h, self.device = lib, h, int(device)
        self._keep = []
        self.rank, self.nranks = 0, 1
        self.m_total = None          # global row count of a row-sharded A = D problem
        self.row_range = None        # (lo, hi) rows of this rank

    def close(self):
        if getattr(self, "_h", None):
            self._lib.admm_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing --------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        L.check(self._lib.admm_b200_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else None)))

    def synchronize(self):
        L.check(self._lib.admm_b200_synchronize(self._h))

    def launch_count(self):
        return int(self._lib.admm_b200_launch_count(self._h))

    def graph_replays(self):
        return int(self._lib.admm_b200_graph_replays(self._h))

    def setup_phases(self):
        out = (C.c_double * 4)()
        L.check(self._lib.admm_b200_get_setup_phases(self._h, out))
        return dict(gram_ms=out[0], chol_ms=out[1], inverse_ms=out[2], total_ms=out[3])

    def default_options(self):
        o = L.Options()
        self._lib.admm_b200_default_options(C.byref(o))
        return o

    @staticmethod
    def _matrix(D):
        if isinstance(D, DeviceMatrix):
            return D.ptr, D.shape[0], D.shape[1], D.ld, D
        if isinstance(D, RowShard):
            D = D.array
        a = np.asarray(D)
        # a column-major matrix or a ROW SLICE of one (strides (8, 8*ld)) is used in place with its leading
        # dimension: slicing the rows of a rank out of a pinned host matrix must not copy it
        if a.ndim == 2 and a.dtype == np.float64 and a.shape[0] > 0 and a.shape[1] > 0 and a.strides[0] == 8 and \
                a.strides[1] % 8 == 0 and a.strides[1] >= 8 * a.shape[0]:
            return a.ctypes.data, a.shape[0], a.shape[1], a.strides[1] // 8, a
        a = L.fmat(D)
        return a.ctypes.data, a.shape[0], a.shape[1], max(a.shape[0], 1), a

    # -- setup -----------------------------------------------------------------------------
    def setup_lasso(self, D, s, rho, xsolve=L.XSOLVE_INVFACTOR):
        p, m, n, ld, keep = self._matrix(D)
        sv = s if isinstance(s, (int, np.integer)) else L.fvec(s, m, "s")
        self._keep = [keep, sv]
        self.m_total = self.row_range = None
        L.check(self._lib.admm_b200_setup_lasso(self._h, m, n, C.c_void_p(p), ld, L.ptr(sv), float(rho), int(xsolve)))
        return m, n

    def setup_lasso_sharded(self, D_local, s_local, rho, m_total, xsolve=L.XSOLVE_INVFACTOR):
        """admm_b200_setup_lasso_sharded: D_local / s_local are THIS rank's rows of the tall problem."""
        p, m, n, ld, keep = self._matrix(D_local)
        sv = s_local if isinstance(s_local, (int, np.integer)) else L.fvec(s_local, m, "s")
        self._keep = [keep, sv]
        self.m_total, self.row_range = int(m_total), None
        L.check(self._lib.admm_b200_setup_lasso_sharded(self._h, m, int(m_total), n, C.c_void_p(p), ld, L.ptr(sv),
                                                        float(rho), int(xsolve)))
        return m, n

    def info(self):
        """admm_b200_get_info as a dict (generation, zero_cols, diag_ratio, xsolve_effective, p2p_ready, ...)."""
        out = L.Info()
        L.check(self._lib.admm_b200_get_info(self._h, C.byref(out)))
        return {k: getattr(out, k) for k, _ in L.Info._fields_}

    @property
    def generation(self):
        return self.info()["generation"]

    def setup_unwrapped(self, kind, D_local, aux_local, Cval=0.0, m_total=None):
        """admm_b200_setup_unwrapped: D_local / aux_local are THIS rank's rows."""
        p, m, n, ld, keep = self._matrix(D_local)
        av = aux_local if isinstance(aux_local, (int, np.integer)) else L.fvec(aux_local, m, "ell / s")
        self._keep = [keep, av]
        self.m_total = int(m_total) if m_total is not None else m
        self.row_range = None            # set by the caller that knows which rows these are
        L.check(self._lib.admm_b200_setup_unwrapped(self._h, int(kind), m, self.m_total, n, C.c_void_p(p), ld,
                                                    L.ptr(av), float(Cval)))
        return m, n

    def setup_basispursuit(self, D, s):
        p, m, n, ld, keep = self._matrix(D)
        sv = s if isinstance(s, (int, np.integer)) else L.fvec(s, m, "s")
        self._keep = [keep, sv]
        self.m_total = self.row_range = None
        L.check(self._lib.admm_b200_setup_basispursuit(self._h, m, n, C.c_void_p(p), ld, L.ptr(sv)))
        return m, n

    def setup_totalvariation(self, s, lam):
        sv = L.fvec(s)
        self._keep = [sv]
        self.m_total = self.row_range = None
        L.check(self._lib.admm_b200_setup_totalvariation(self._h, sv.size, L.ptr(sv), float(lam)))
        return sv.size

    def setup_quadratic(self, kind, P, q, r, rho, lb=None, ub=None):
        P = L.fmat(P)
        n = P.shape[0]
        q = L.fvec(q, n, "q")
        lb = None if lb is None else L.fvec(lb, n, "lb")
        ub = None if ub is None else L.fvec(ub, n, "ub")
        self._keep = [P, q, lb, ub]
        self.m_total = self.row_range = None
        L.check(self._lib.admm_b200_setup_quadratic(self._h, int(kind), n, L.ptr(P), n, L.ptr(q), float(r), float(rho),
                                                    L.ptr(lb), L.ptr(ub)))
        return n

    def setup_model(self, P, Q, r, s, rho):
        """admm_b200_setup_model: PtP, QtQ, Ptr, Qts and both Cholesky factors on the device."""
        pp, m, n, ldp, keepP = self._matrix(P)
        qp, mq, nq, ldq, keepQ = self._matrix(Q)
        rv, sv = L.fvec(r, m, "r"), L.fvec(s, m, "s")
        self._keep = [keepP, keepQ, rv, sv]
        self.m_total = self.row_range = None
        L.check(self._lib.admm_b200_setup_model(self._h, m, n, C.c_void_p(pp), ldp, C.c_void_p(qp), ldq, L.ptr(rv),
                                                L.ptr(sv), float(rho)))
        return m, n

    # -- row-sharded runs (one process per GPU) ---------------------------------------------------
    def comm_init(self, rank, nranks, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        L.check(self._lib.admm_b200_comm_init(self._h, int(rank), int(nranks), buf))
        self.rank, self.nranks = int(rank), int(nranks)

    def comm_ipc_export(self, rank, nranks):
        """Mailbox-only transport, step 1: allocate this rank's mailbox, return its 64-byte CUDA IPC handle."""
        buf = (C.c_char * 64)()
        L.check(self._lib.admm_b200_comm_ipc_export(self._h, int(rank), int(nranks), buf))
        self.rank, self.nranks = int(rank), int(nranks)
        return bytes(buf.raw)

    def comm_ipc_attach(self, handles):
        """Step 2: map every rank's mailbox (handles in rank order)."""
        blob = b"".join(bytes(hd) for hd in handles)
        if len(blob) != 64 * self.nranks:
            raise ValueError("comm_ipc_attach: need one 64-byte handle per rank")
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        try:
            L.check(self._lib.admm_b200_comm_ipc_attach(self._h, buf))
        except L.EngineError:
            self.rank, self.nranks = 0, 1       # the library detached the handle (a peer's mailbox could not be mapped)
            raise

    def comm_destroy(self):
        L.check(self._lib.admm_b200_comm_destroy(self._h))
        self.rank, self.nranks = 0, 1

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        L.check(L.load().admm_b200_get_unique_id(buf))
        return bytes(buf.raw)

    def allreduce(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        L.check(self._lib.admm_b200_allreduce(self._h, L.ptr(a), a.size))
        return a

    def set_lambda(self, lam):
        L.check(self._lib.admm_b200_set_lambda(self._h, float(lam)))

    def dims(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        L.check(self._lib.admm_b200_get_dims(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_init(self, x0=None, z0=None, u0=None):
        nA, nB, m = self.dims()
        x0 = None if x0 is None else L.fvec(x0, nA, "x0")
        z0 = None if z0 is None else L.fvec(z0, nB, "z0")
        u0 = None if u0 is None else L.fvec(u0, m, "u0")
        L.check(self._lib.admm_b200_set_init(self._h, L.ptr(x0), L.ptr(z0), L.ptr(u0)))

    def get_factor(self):
        k = C.c_int64()
        L.check(self._lib.admm_b200_get_factor(self._h, None, 0, C.byref(k)))
        out = np.zeros((k.value, k.value), order="F")
        L.check(self._lib.admm_b200_get_factor(self._h, L.ptr(out), k.value, C.byref(k)))
        return out

    # -- loop ------------------------------------------------------------------------------
    def solve(self, opts, want_history=True):
        """Runs admm.m's loop on the device; returns a dict with the reference's result fields."""
        nA, nB, m = self.dims()
        N = int(opts.maxiters) if opts.maxiters > 0 else 1000
        res = L.Result()
        bufs = {k: np.full(N, np.nan) for k in ("pnorm", "dnorm", "perr", "derr", "hnormsq", "objevals", "dvals", "avals",
                                                "restarted")}
        xo, zo, uo = np.zeros(nA), np.zeros(nB), np.zeros(m)
        res.xopt, res.zopt, res.uopt = (a.ctypes.data_as(L._dp) for a in (xo, zo, uo))
        for k, a in bufs.items():
            setattr(res, k, a.ctypes.data_as(L._dp))
        hist = None
        if want_history and opts.history:
            hist = [np.zeros((nA, N), order="F"), np.zeros((nB, N), order="F"), np.zeros((m, N), order="F")]
            res.xvals, res.zvals, res.uvals = (a.ctypes.data_as(L._dp) for a in hist)
        L.check(self._lib.admm_b200_solve(self._h, C.byref(opts), C.byref(res)))
        k = int(res.steps)
        out = dict(steps=k, status=int(res.status), xopt=xo, zopt=zo, uopt=uo, objopt=float(res.objopt),
                   setup_ms=float(res.setup_ms), loop_ms=float(res.loop_ms))
        for name, a in bufs.items():
            out[name] = a[:k].copy()
        if hist is not None:
            out["xvals"], out["zvals"], out["uvals"] = (np.ascontiguousarray(a[:, :k]) for a in hist)
        return out

    def solve_lasso_batch(self, opts, lambdas, want_history=True):
        """nb lasso problems (one per lambda) on the cached factor; returns per-column results."""
        nA, _, _ = self.dims()
        lam = L.fvec(lambdas)
        nb = lam.size
        N = int(opts.maxiters) if opts.maxiters > 0 else 1000
        steps = np.zeros(nb, dtype=np.int64)
        status = np.zeros(nb, dtype=np.int32)
        X, Z, U = (np.zeros((nA, nb), order="F") for _ in range(3))
        hist = [np.full((N, nb), np.nan, order="F") for _ in range(4)] if want_history else [None] * 4
        ms = C.c_double()
        L.check(self._lib.admm_b200_solve_lasso_batch(self._h, C.byref(opts), nb, L.ptr(lam), L.ptr(steps), L.ptr(status),
                                                      L.ptr(X), L.ptr(Z), L.ptr(U), *(L.ptr(a) for a in hist), C.byref(ms)))
        out = dict(steps=steps, status=status, xopt=X, zopt=Z, uopt=U, loop_ms=ms.value)
        if want_history:
            out.update(pnorm=hist[0], dnorm=hist[1], perr=hist[2], derr=hist[3])
        return out

    def solve_unwrapped_batch(self, opts, aux, X0=None, Z0=None, U0=None, want_history=True):
        """nb A = D problems sharing the D of the last setup_unwrapped (one-vs-all SVM); aux is
        m_local x nb (or a device pointer with ld = m_local)."""
        n, _, m = self.dims()
        if isinstance(aux, tuple):               # (device pointer, nb)
            aux_p, nb, keep = C.c_void_p(int(aux[0])), int(aux[1]), None
        else:
            keep = L.fmat(aux)
            nb = keep.shape[1]
            aux_p = L.ptr(keep)
        mats = []
        for a, rows in ((X0, n), (Z0, m), (U0, m)):
            mats.append(None if a is None else L.fmat(np.asarray(a, dtype=np.float64).reshape(rows, nb, order="F")))
        N = int(opts.maxiters) if opts.maxiters > 0 else 1000
        steps = np.zeros(nb, dtype=np.int64)
        status = np.zeros(nb, dtype=np.int32)
        X = np.zeros((n, nb), order="F")
        Z, U = (np.zeros((m, nb), order="F") for _ in range(2)) if want_history else (None, None)
        hist = [np.full((N, nb), np.nan, order="F") for _ in range(3)] if want_history else [None] * 3
        ms = C.c_double()
        L.check(self._lib.admm_b200_solve_unwrapped_batch(self._h, C.byref(opts), nb, aux_p, m, *(L.ptr(a) for a in mats),
                                                          L.ptr(steps), L.ptr(status), L.ptr(X), L.ptr(Z), L.ptr(U),
                                                          *(L.ptr(a) for a in hist), C.byref(ms)))
        out = dict(steps=steps, status=status, xopt=X, loop_ms=ms.value)
        if want_history:
            out.update(zopt=Z, uopt=U, pnorm=hist[0], perr=hist[1], objevals=hist[2])
        return out

    def iterate_raw(self, opts, which=0, reps=1):
        L.check(self._lib.admm_b200_iterate_raw(self._h, C.byref(opts), int(which), int(reps)))

    # -- building blocks -----------------------------------------------------------------------
    def dgemm(self, transa, transb, alpha, A, B, beta=0.0, Cmat=None, lower_only=False):
        A, B = L.fmat(A), L.fmat(B)
        M = A.shape[1] if transa else A.shape[0]
        K = A.shape[0] if transa else A.shape[1]
        N = B.shape[0] if transb else B.shape[1]
        Kb = B.shape[1] if transb else B.shape[0]
        if K != Kb:
            raise L.EngineError(L.ERR_INVALID, "dgemm: inner dimensions differ")
        out = np.zeros((M, N), order="F") if Cmat is None else np.array(Cmat, dtype=np.float64, order="F")
        L.check(self._lib.admm_b200_dgemm(self._h, int(bool(transa)), int(bool(transb)), M, N, K, float(alpha),
                                          L.ptr(A), max(A.shape[0], 1), L.ptr(B), max(B.shape[0], 1), float(beta),
                                          L.ptr(out), max(M, 1), int(bool(lower_only))))
        return out

    def gram(self, D, trans=True, scale=1.0, shift=0.0):
        p, m, n, ld, keep = self._matrix(D)
        k = n if trans else m
        G = np.zeros((k, k), order="F")
        L.check(self._lib.admm_b200_gram(self._h, int(bool(trans)), m, n, C.c_void_p(p), ld, float(scale),
                                         float(shift), L.ptr(G), k))
        return G

    def potrf(self, A, want_inverse=False):
        A = np.array(A, dtype=np.float64, order="F")
        k = A.shape[0]
        W = np.zeros((k, k), order="F") if want_inverse else None
        L.check(self._lib.admm_b200_potrf(self._h, k, L.ptr(A), k, L.ptr(W), k))
        return (A, W) if want_inverse else A

    def factor_solve(self, b, xsolve=L.XSOLVE_INVFACTOR):
        b = L.fvec(b)
        x = np.zeros_like(b)
        L.check(self._lib.admm_b200_factor_solve(self._h, L.ptr(b), L.ptr(x), int(xsolve)))
        return x


def acquire_engine(engine, options):
    """The engine a solver call runs on: the caller's (argument or options['engine']) or a fresh one on
    options['device'].  A fresh one is marked `_owned`; admm() closes it when the loop has returned, so a
    solver call without an explicit engine does not hold GPU buffers until garbage collection."""
    eng = engine or options.get("engine")
    if eng is not None:
        return eng
    eng = Engine(int(options.get("device", 0)))
    eng._owned = True
    return eng


def slicemaker(length, workers):
    """errorcheck.m:249-259 through the C-ABI (no GPU needed)."""
    out = (C.c_int64 * int(workers))()
    L.check(L.load().admm_b200_slicemaker(int(length), int(workers), out))
    return [int(v) for v in out]
