"""admm_project_b200 -- Blackwell (B200, sm_100a) FP64 engine for the hot path of
PeterSutor/ADMM-Project, behind the reference's own call signatures:

    results = admm(xminf, zming, options)                    (admm.m:24)
    [minx, minz, extra] = getproxops(problem, args)          (getProxOps.m:13)
    results = lasso(D, s, lambda, options)  ...              (solvers/*.m)

The arithmetic runs in libadmm_b200.so (hand-written CUDA, C-ABI in include/admm_b200.h).  There
is no CPU fallback: without the built library and a B200 every compute entry point raises.
"""
from ._lib import EngineError                                    # noqa: F401
from .engine import Engine, DeviceMatrix, RowShard, slicemaker             # noqa: F401
from .admm import admm, setopt, MatlabError                      # noqa: F401
from .getproxops import getproxops, EngineProx                   # noqa: F401
from .errorcheck import errorcheck                               # noqa: F401
from . import solvers                                            # noqa: F401
from . import mnist                                              # noqa: F401
from . import testers                                            # noqa: F401
from .showresults import showresults                             # noqa: F401
from .solvers import linearsvm_onevsall, lasso_path                       # noqa: F401
from .solvers import lasso, unwrappedadmm, linearsvm, huberfit, lad, basispursuit, totalvariation, quadraticprogram, model   # noqa: F401
