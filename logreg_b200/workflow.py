"""The steps either side of the hot path, as thin host helpers around the
device kernels: data ingest (f2), MAP initialisation (f1), output + summaries (f3).

Reference lines mirrored (paths relative to the reference root):
    load_pima       Python/fit-numpy.py:12-19 (parquet -> X with a ones column, y float32)
                    C/fit-bayes.c:46-68 reads the same table from the text file pima.data
    map_estimate    method="newton": Python/fit-jax.py:62-79, Newton with the exact Hessian and step
                    halving, on the device (lrb_map: X'WX block kernel, Cholesky solve, decisions);
                    method="BFGS": Python/fit-np-ul.py:54  minimize(-lpost, init, jac=-glp, method='BFGS'),
                    SciPy on the host around the fused evaluation (one pass per iteration)
    describe_device running mean / covariance accumulated on the device during the run (lrb_run_moments;
                    Dex/djwutils.dx:97-103 meanAndCovariance)
    save_samples    Python/fit-numpy.py:89-90  DataFrame(out, columns=b0..).to_parquet(...)
    describe        Python/fit-numpy.py:92-96  scipy.stats.describe(out): mean / variance
"""
from __future__ import annotations

import numpy as np


def load_pima(path):
    """Pima.tr as the scripts build it: X (n x 8, float64, column-major, leading ones), y float32.
    Accepts pima.parquet (pandas/pyarrow), the space-separated pima.data, or an .npz with X, y."""
    path = str(path)
    if path.endswith(".npz"):
        g = np.load(path)
        return np.asfortranarray(g["X"]), g["y"].astype(np.float32)
    if path.endswith(".parquet"):
        import pandas as pd
        df = pd.read_parquet(path)
        n = df.shape[0]
        y = pd.get_dummies(df["type"])["Yes"].to_numpy(dtype="float32")
        X = df.drop(columns="type").to_numpy()
        return np.asfortranarray(np.hstack((np.ones((n, 1)), X))), y
    rows = [line.split() for line in open(path) if line.strip()]
    y = np.array([1.0 if r[-1].strip('"') == "Yes" else 0.0 for r in rows], dtype=np.float32)
    X = np.array([[float(v) for v in r[:-1]] for r in rows])
    return np.asfortranarray(np.hstack((np.ones((len(rows), 1)), X))), y


class MapResult(dict):
    """Result of map_estimate: attribute access like scipy's OptimizeResult (x, fun, nit, success)."""
    __getattr__ = dict.__getitem__


def map_estimate(problem, init, method="newton", tol=None, maxit=None, **kw):
    """MAP estimate, the `init` of the samplers.

    method="newton" (default): the reference's Newton loop (fit-jax.py:62-79) entirely on the
        device -- gradient and lpost from the fused kernel, the Hessian X'WX + diag(pscale^-2) from
        a float64 block kernel, Cholesky solve and step halving in device kernels (lrb_map).
        tol (default 0.01, the reference's) bounds ||glp||; maxit defaults to 500.
    method="BFGS" (or any scipy.optimize.minimize method): quasi-Newton with the hand-coded gradient
        as fit-np-ul.py:54; lpost and glp at the same point come from ONE fused pass over X (the
        reference spends three per iteration), the iteration itself runs in SciPy on the host."""
    init = np.asarray(init, dtype=np.float64)
    if method.lower() == "newton":
        x, info = problem.map(init, tol=0.01 if tol is None else tol, maxit=500 if maxit is None else maxit)
        return MapResult(x=x, fun=-info["lpost"], nit=info["iterations"], success=bool(info["converged"]),
                         nfev=info["evals"], halvings=info["halvings"], grad_norm=info["grad_norm"], method="newton")
    from scipy.optimize import minimize
    if tol is not None:
        kw.setdefault("tol", tol)
    if maxit is not None:
        kw.setdefault("options", {}).setdefault("maxiter", maxit)

    def fun(b):
        lp, _, g = problem.eval(b, want_grad=True)
        return -lp, -g

    res = minimize(fun, init, jac=True, method=method, **kw)
    return res


def describe_device(problem, pooled=True):
    """The same summary from the moments the device accumulated during run(..., moments=True):
    nothing but O(p^2) numbers crosses the bus (what config 4's 4096 chains need)."""
    n, mean, cov = problem.moments(pooled=pooled, cov=True)
    var = np.diagonal(cov, axis1=-2, axis2=-1)
    return {"nobs": n, "mean": mean, "variance": var, "covariance": cov}


def describe(mat):
    """mean / variance per coefficient (ddof=1), the numbers fit-numpy.py:95-96 prints."""
    mat = np.asarray(mat, dtype=np.float64)
    return {"nobs": mat.shape[0], "mean": mat.mean(axis=0), "variance": mat.var(axis=0, ddof=1),
            "min": mat.min(axis=0), "max": mat.max(axis=0)}


def effective_sample_size(mat, max_lag=200):
    """Per-coefficient ESS from the initial positive sequence of autocorrelations (what
    smfsb::mcmcSummary reports in Python/analyse.R:16)."""
    mat = np.asarray(mat, dtype=np.float64)
    n, p = mat.shape
    out = np.empty(p)
    for j in range(p):
        x = mat[:, j] - mat[:, j].mean()
        v = x.dot(x) / n
        if v == 0:
            out[j] = n
            continue
        s = 0.0
        for k in range(1, min(max_lag, n - 1)):
            r = x[:-k].dot(x[k:]) / (n * v)
            if r <= 0:
                break
            s += r
        out[j] = n / (1 + 2 * s)
    return out


def save_samples(mat, path):
    """Samples -> parquet (columns b0..b{p-1}, fit-numpy.py:89-90) or .npy / .tsv by extension."""
    path = str(path)
    cols = [f"b{j}" for j in range(mat.shape[1])]
    if path.endswith(".parquet"):
        import pandas as pd
        pd.DataFrame(mat, columns=cols).to_parquet(path)
    elif path.endswith(".tsv"):
        np.savetxt(path, mat, delimiter="\t", header="\t".join(cols), comments="")
    else:
        np.save(path, mat)
    return path
