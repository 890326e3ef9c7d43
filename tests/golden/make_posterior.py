#!/usr/bin/env python3
"""Posterior summaries of Pima from the REFERENCE samplers (AST-lifted, unmodified, run in the
build container): the acceptance bar of BASELINE.json config 1 ("posterior means and standard
deviations must lie within Monte Carlo error of the reference samplers").

    python tests/golden/make_posterior.py        # ~3 minutes of CPU

Stores, per sampler, mean / sd / effective sample size of a 2000-row thinned chain started at the
MAP with the reference scripts' tuning constants (thinning reduced so it finishes).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import make_golden as mg  # noqa: E402
from logreg_b200.workflow import effective_sample_size  # noqa: E402


def main():
    X, y, n, p = mg.load_pima()
    g = dict(np.load(os.path.join(HERE, "pima.npz")))
    init = g["map"].copy()
    pre = np.array([100., 1., 1., 1., 1., 1., 25., 1.])
    out = {}
    rw = mg.namespace("fit-numpy.py", {"ll", "lprior", "lpost", "mhKernel", "mcmc", "pre", "rprop"}, X=X, y=y, n=n, p=p, init=init)
    ul = mg.namespace("fit-np-ul.py", {"ll", "pscale", "lprior", "lpost", "glp", "ulKernel", "mcmc"}, X=X, y=y, n=n, p=p, init=init)
    ma = mg.namespace("fit-np-mala.py", {"ll", "pscale", "lprior", "lpost", "glp", "mhKernel", "malaKernel", "mcmc"}, X=X, y=y, n=n, p=p, init=init)
    hm = mg.namespace("fit-np-hmc.py", {"ll", "pscale", "lprior", "lpost", "glp", "mhKernel", "hmcKernel", "mcmc"}, X=X, y=y, n=n, p=p, init=init)
    runs = {
        "rwmh": (rw["mcmc"], lambda: rw["mhKernel"](rw["lpost"], rw["rprop"]), 200),
        "ul": (ul["mcmc"], lambda: ul["ulKernel"](ul["glp"], dt=1e-6, pre=pre), 400),
        "mala": (ma["mcmc"], lambda: ma["malaKernel"](ma["lpost"], ma["glp"], dt=1e-5, pre=pre), 200),
        "hmc": (hm["mcmc"], lambda: hm["hmcKernel"](hm["lpost"], hm["glp"], eps=1e-3, l=50, dmm=1 / pre), 10),
    }
    for k, (mcmc, kern, thin) in runs.items():
        np.random.seed(2024)
        mat = mcmc(init, kern(), thin=thin, iters=2000, verb=False)
        out[k + "_mean"] = mat.mean(0)
        out[k + "_sd"] = mat.std(0, ddof=1)
        out[k + "_ess"] = effective_sample_size(mat)
        out[k + "_thin"] = np.array(thin)
        print(k, "mean", np.round(out[k + "_mean"], 3), "ess", np.round(out[k + "_ess"]))
    np.savez_compressed(os.path.join(HERE, "pima_posterior.npz"), **out)


if __name__ == "__main__":
    main()
