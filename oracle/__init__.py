"""ORACLE -- CPU (NumPy/SciPy, FP64) restatement of the hot path of PeterSutor/ADMM-Project:
admm.m, the in-scope getProxOps.m operators, errorcheck.m's slicemaker and the one-time setup
of solvers/{lasso,unwrappedadmm,linearsvm,huberfit,lad,totalvariation,basispursuit,model}.m.

THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  PARITY UNPINNED: the reference is MATLAB,
no MATLAB/Octave exists in this image and the reference ships no golden vectors, so the oracle
is pinned only to the algorithm text (file:line cited per function) and to hand-derived
known answers in tests/test_oracle_known_answers.py.
"""
from .admm import MatlabError, admm, setopt, slice_ranges          # noqa: F401
from .errorcheck import errorcheck, slicemaker                     # noqa: F401
from .getproxops import (getproxops, zminSoftThresholding, minz01, zminNonNegative,   # noqa: F401
                         make_zminBox, subplus, pos, huber)
from .solvers import (lasso, unwrappedadmm, linearsvm, huberfit, lad,                 # noqa: F401
                      totalvariation, basispursuit, basispursuit_factored, quadraticprogram, model)
