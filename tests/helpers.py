import numpy as np


def ungrad_scale(X, y, beta, pscale):
    """Un-cancelled magnitude of glp: |X|'|y-p| + |beta/pscale^2|. Relative
    gradient errors are measured against its max (hard part 3:
    glp -> 0 at the MAP, so a component-wise relative error is meaningless there)."""
    X = np.asarray(X, dtype=np.float64)
    pr = 1 / (1 + np.exp(-X.dot(beta)))
    return float(np.max(np.abs(X).T.dot(np.abs(y - pr)) + np.abs(beta / pscale ** 2)))
