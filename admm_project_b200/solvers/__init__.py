"""Host-side mirrors of solvers/*.m: argument checks as in the reference, one-time setup on the
device (Gram / Cholesky / inverse factor), then ``admm``."""
from .lasso import lasso                     # noqa: F401
