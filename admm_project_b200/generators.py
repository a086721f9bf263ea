"""Seeded synthetic problem generators mirroring the reference testers' recipes
(/root/reference/testers/*.m).  Pure NumPy, no engine import: tests feed the SAME arrays to
the oracle and to the CUDA engine, so MATLAB's RNG stream is irrelevant to parity
(SURVEY.md section 8c)."""
from __future__ import annotations

import math

import numpy as np


def lasso_problem(seed, rows, cols):
    """testers/lassotest.m:109-122.  Returns D (rows x cols, Fortran order), s, lambda, testx."""
    rs = np.random.RandomState(seed)
    d = 3 / 5
    testx = rs.randn(cols) * (rs.rand(cols) < d)              # sprandn(cols,1,d)
    D = np.asfortranarray(rs.randn(rows, cols))
    D /= np.sqrt(np.sum(D * D, axis=0))[None, :]              # unit-norm columns (:116)
    s = D @ testx + math.sqrt(0.001) * rs.randn(rows)         # :119
    lam = 0.1 * float(np.max(np.abs(D.T @ s)))                # :121-122
    return D, s, lam, testx


def svm_problem(seed, mpos, mneg, sep=0.2):
    """testers/linearsvmtest.m:133-144 (two clouds either side of x1 = x2)."""
    rs = np.random.RandomState(seed)
    base = np.linspace(0.0, 2.0, mpos)                        # (0:2/(mpos-1):2)'
    posp = np.stack([base + rs.rand(mpos) - sep * rs.rand(mpos),
                     base - rs.rand(mpos) + sep * rs.rand(mpos)], axis=1)
    basen = base[:mneg] if mneg <= mpos else np.linspace(0.0, 2.0, mneg)
    negp = np.stack([basen - rs.rand(mneg) + sep * rs.rand(mneg),
                     basen + rs.rand(mneg) - sep * rs.rand(mneg)], axis=1)
    D = np.asfortranarray(np.vstack([posp, negp]))
    ell = np.ones(mpos + mneg)
    ell[mpos:] = -1
    return D, ell


def svm_mnist_like(seed, rows, cols, nclass=10, labels=None):
    """BASELINE.json config 3: MNIST-shaped synthetic features (U(0,1) with 81% zeros, stored
    dense) and one-vs-all +-1 label columns.  ``labels`` (ints 0..9) may come from the real
    train-labels file (examples/mnistsvm.m:94,136-142); synthetic uniform digits otherwise."""
    rs = np.random.RandomState(seed)
    D = np.asfortranarray(rs.rand(rows, cols) * (rs.rand(rows, cols) < 0.19))
    if labels is None:
        labels = rs.randint(0, nclass, size=rows)
    ell = np.where(labels[:, None] == np.arange(nclass)[None, :], 1.0, -1.0)
    return D, np.asfortranarray(ell)


def huber_problem(seed, rows, cols):
    """testers/huberfittest.m:121-128."""
    rs = np.random.RandomState(seed)
    testx = rs.randn(cols)
    D = np.asfortranarray(rs.randn(rows, cols))
    D /= np.sqrt(np.sum(D * D, axis=0))[None, :]
    s = D @ testx + math.sqrt(0.01) * rs.randn(rows)
    mask = rs.rand(rows) < min(1.0, 200.0 / rows)             # sprand(rows,1,200/rows)
    s = s + 10 * rs.rand(rows) * mask
    return D, s, testx


def lad_problem(seed, rows, cols):
    """testers/ladtest.m:115-123."""
    rs = np.random.RandomState(seed)
    D = np.asfortranarray(rs.randn(rows, cols))
    xtrue = 10 * rs.randn(cols)
    s = D @ xtrue
    idx = rs.choice(rows, size=int(math.ceil(rows / 50)), replace=False)   # randsample
    s[idx] = s[idx] + 100 * rs.randn(idx.size)
    return D, s, xtrue


def tv_problem(seed, rows):
    """testers/totalvariationtest.m:109-127."""
    rs = np.random.RandomState(seed)
    truth = np.ones(rows)
    for _ in range(3):
        r = int(rs.randint(1, rows + 1))
        ri = int(rs.randint(1, 11))
        lo = int(math.ceil(r / 2))
        truth[lo - 1:r] = ri * truth[lo - 1:r]
    s = truth + rs.randn(rows)
    return s, truth


def bp_problem(seed, rows, cols, density=0.1):
    """testers/basispursuittest.m:114-117 (density 0.1, SURVEY.md section 8d note)."""
    rs = np.random.RandomState(seed)
    D = np.asfortranarray(rs.randn(rows, cols))
    testx = rs.randn(cols) * (rs.rand(cols) < density)
    s = D @ testx
    return D, s, testx


def model_problem(seed, rows, cols):
    """testers/modeltest.m:112-120: P, Q ~ N(0,1) (rows x cols), r, s ~ N(0,1); the true minimiser of
    1/2||Px-r||^2 + 1/2||Qx-s||^2 is (P'P + Q'Q) \\ (P'r + Q's)."""
    rs = np.random.RandomState(seed)
    P = np.asfortranarray(rs.randn(rows, cols))
    Q = np.asfortranarray(rs.randn(rows, cols))
    r, s = rs.randn(rows), rs.randn(rows)
    truex = np.linalg.solve(P.T @ P + Q.T @ Q, P.T @ r + Q.T @ s)
    return P, Q, r, s, truex


def lasso_problem_big(seed, rows, cols):
    """testers/lassotest.m:109-122 at BASELINE sizes (65536 x 8192 is 4.3 GB): the same recipe as lasso_problem,
    drawn in column blocks by independent PCG64 streams on a thread pool so it takes seconds."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    nblk = (cols + 255) // 256
    kids = np.random.SeedSequence(seed).spawn(nblk + 1)
    rs = np.random.default_rng(kids[-1])
    testx = rs.standard_normal(cols) * (rs.random(cols) < 0.6)
    D = np.empty((rows, cols), order="F")

    def fill(b):
        j0 = b * 256
        w = min(256, cols - j0)
        blk = np.random.default_rng(kids[b]).standard_normal((w, rows))
        blk /= np.sqrt(np.einsum("ij,ij->i", blk, blk))[:, None]
        D[:, j0:j0 + w] = blk.T
    with ThreadPoolExecutor(max_workers=len(os.sched_getaffinity(0))) as ex:
        list(ex.map(fill, range(nblk)))
    s = D @ testx + math.sqrt(0.001) * rs.standard_normal(rows)
    lam = 0.1 * float(np.max(np.abs(D.T @ s)))
    return D, s, lam, testx


def randn_big(seed, rows, cols, colnorm=False):
    """rows x cols N(0,1) in Fortran order, drawn in parallel column blocks (huber / lad / bp at BASELINE sizes)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    nblk = (cols + 127) // 128
    kids = np.random.SeedSequence(seed).spawn(nblk)
    D = np.empty((rows, cols), order="F")

    def fill(b):
        j0 = b * 128
        w = min(128, cols - j0)
        blk = np.random.default_rng(kids[b]).standard_normal((w, rows))
        if colnorm:
            blk /= np.sqrt(np.einsum("ij,ij->i", blk, blk))[:, None]
        D[:, j0:j0 + w] = blk.T
    with ThreadPoolExecutor(max_workers=len(os.sched_getaffinity(0))) as ex:
        list(ex.map(fill, range(nblk)))
    return D
