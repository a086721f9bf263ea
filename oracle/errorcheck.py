"""ORACLE (test infrastructure) -- restatement of the slice partitioner in
``/root/reference/errorcheck.m`` (``'slices'`` case :123-134, ``slicemaker`` :216-267).
Parity unpinned (no MATLAB/Octave here; see oracle/admm.py).  The remaining validators of
errorcheck.m are host-side argument checks and out of scope."""
from __future__ import annotations

import math

import numpy as np

from .admm import MatlabError


def slicemaker(slices, workers, length):
    """errorcheck.m:216-267.  Returns a list of ints (MATLAB row vector)."""
    arr = np.atleast_1d(np.asarray(slices))
    if arr.ndim != 1 or not np.issubdtype(arr.dtype, np.number):
        raise MatlabError("Argument slices is not a numeric vector or integer!")
    arr = np.floor(np.real(arr)).astype(np.int64)                      # :231
    workers, length = int(workers), int(length)
    if arr.size == 1 and arr[0] > 0:                                   # :239-243
        slicesize = int(arr[0])
        nfull, nceil = length // slicesize, math.ceil(length / slicesize)
        out = [0] * max(nfull, nceil, 1)
        out[0] = slicesize                                             # slices(1) was slicesize
        for k in range(nfull):
            out[k] = slicesize
        # slices(ceil(len/slicesize)) = mod(len, slicesize): when slicesize divides len this
        # overwrites the last full block with 0 -- the reference's off-by-one, kept (:242).
        if nceil >= 1:
            out[nceil - 1] = length % slicesize
        return out
    elif arr.size == 1 and arr[0] == 0:                                # :249-259
        if length % workers != 0:
            rem = length % workers
            slicesize = length // workers
            return [slicesize + 1] * rem + [slicesize] * (workers - rem)
        return [length // workers] * workers
    elif int(arr.sum()) != length:                                     # :263-265
        raise MatlabError("The number of parallel slices does not match length of x!")
    return [int(v) for v in arr]


def errorcheck(arg, check, name, options=None):
    """errorcheck.m:17 -- only the 'slices' case (errorcheck.m:123-134) is restated."""
    options = options or {}
    if check == "slices":
        if "slicelength" in options and "workers" in options:
            return slicemaker(arg, int(math.floor(options["workers"])),
                              int(math.floor(options["slicelength"])))
        elif "slicelength" not in options:
            raise MatlabError("Did not provide slicelength in options struct!")
        raise MatlabError("Did not provide workers in options struct!")
    raise MatlabError("oracle errorcheck: check '%s' is out of scope" % check)
