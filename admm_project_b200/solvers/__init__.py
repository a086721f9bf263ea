"""Host-side mirrors of solvers/<name>.m: argument checks as in the reference, one-time setup on the
device (Gram / Cholesky / inverse factor), then ``admm``."""
from .lasso import lasso, lasso_path           # noqa: F401
from .unwrappedadmm import unwrappedadmm       # noqa: F401
from .linearsvm import linearsvm, linearsvm_onevsall   # noqa: F401
from .robustfit import huberfit, lad           # noqa: F401
from .basispursuit import basispursuit        # noqa: F401
from .totalvariation import totalvariation    # noqa: F401
from .quadraticprogram import quadraticprogram  # noqa: F401
from .model import model                        # noqa: F401
