"""Host-side mirror of errorcheck.m's 'slices' case (errorcheck.m:123-134, slicemaker :216-267).
The balanced partition (slices == 0) is the row partition of the multi-GPU solvers and is computed
by the C-ABI (admm_b200_slicemaker); the two user-specified forms are plain host logic."""
from __future__ import annotations

import math

import numpy as np

from .engine import slicemaker as _balanced


class MatlabError(RuntimeError):
    """MATLAB error(...) with the reference's message text."""


def slicemaker(slices, workers, length):
    arr = np.atleast_1d(np.asarray(slices))
    if arr.ndim != 1 or not np.issubdtype(arr.dtype, np.number):
        raise MatlabError("Argument slices is not a numeric vector or integer!")
    arr = np.floor(np.real(arr)).astype(np.int64)                      # errorcheck.m:231
    workers, length = int(workers), int(length)
    if arr.size == 1 and arr[0] > 0:                                   # :239-243 (off-by-one kept)
        k = int(arr[0])
        nfull, nceil = length // k, -(-length // k)
        out = [0] * max(nfull, nceil, 1)
        out[0] = k
        for i in range(nfull):
            out[i] = k
        if nceil >= 1:
            out[nceil - 1] = length % k
        return out
    if arr.size == 1 and arr[0] == 0:                                  # :249-259
        return _balanced(length, workers)
    if int(arr.sum()) != length:                                       # :263-265
        raise MatlabError("The number of parallel slices does not match length of x!")
    return [int(v) for v in arr]


def errorcheck(arg, check, name, options=None):
    """errorcheck.m:17 -- the 'slices' case; the other validators stay in each solver."""
    options = options or {}
    if check == "slices":
        if "slicelength" in options and "workers" in options:
            return slicemaker(arg, int(math.floor(options["workers"])), int(math.floor(options["slicelength"])))
        if "slicelength" not in options:
            raise MatlabError("Did not provide slicelength in options struct!")
        raise MatlabError("Did not provide workers in options struct!")
    raise MatlabError("errorcheck: check '%s' is not part of the engine's host mirror" % check)
