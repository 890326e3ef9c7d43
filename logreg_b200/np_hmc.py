"""Import switch for Python/fit-np-hmc.py: the same names with that script's signatures.

    from logreg_b200.np_hmc import *       # ll, lprior, lpost, glp, mhKernel, hmcKernel, mcmc
    bind_data(X, y, pscale)

Differs from the package top level only in `mhKernel`: fit-np-hmc.py:56-63 defines the
two-argument form whose kernel takes and returns the state alone (kernel(x) -> x, lpost
re-evaluated at x every step); fit-numpy.py:53-62 defines the form that threads the old
log-density (kernel(x, ll) -> (x, ll)), which is what `logreg_b200.mhKernel` returns.
"""
from .api import (RandomWalk, bind_data, glp, hmcKernel, ll, lpost, lprior, mcmc,  # noqa: F401
                  mhKernel as _mhKernel)


def mhKernel(lpost, rprop):
    """fit-np-hmc.py:56-63: kernel(x) -> x with a = lpost(prop) - lpost(x)."""
    return _mhKernel(lpost, rprop, threaded=False)


__all__ = ["ll", "lprior", "lpost", "glp", "mhKernel", "hmcKernel", "mcmc", "bind_data", "RandomWalk"]
