"""logreg_b200 -- B200-native backend for the hot path of darrenjw/logreg.

Exports the names the reference's NumPy scripts define (ll, lprior, lpost, glp,
mhKernel, ulKernel, malaKernel, hmcKernel, mcmc) plus bind_data() for the data
those scripts keep in module globals. See logreg_b200/api.py.
"""
from ._native import LogregB200Error, device_count, library_path  # noqa: F401
from .api import (DeviceKernel, Problem, RandomWalk, bind_data, current, glp, hmcKernel, ll,  # noqa: F401
                  lpost, lprior, malaKernel, mcmc, mhKernel, ulKernel, use)

__all__ = ["ll", "lprior", "lpost", "glp", "mhKernel", "ulKernel", "malaKernel", "hmcKernel", "mcmc",
           "bind_data", "RandomWalk", "Problem", "DeviceKernel", "LogregB200Error", "use", "current"]
