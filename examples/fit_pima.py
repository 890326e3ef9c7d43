#!/usr/bin/env python3
"""The reference's four NumPy scripts (Python/fit-numpy.py, fit-np-ul.py, fit-np-mala.py,
fit-np-hmc.py) with the logreg_b200 backend: same flow (data -> MAP -> mcmc -> save -> describe),
same tuning constants, the density / gradient / sampler loop on the GPU.

    python examples/fit_pima.py --sampler mala [--data ../pima.parquet] [--iters 10000] [--out fit.parquet]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from logreg_b200 import (RandomWalk, bind_data, glp, hmcKernel, ll, lpost, malaKernel, mcmc, mhKernel,  # noqa: E402
                         ulKernel)
from logreg_b200.workflow import describe, effective_sample_size, load_pima, map_estimate, save_samples  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sampler", default="mala", choices=["rwmh", "ul", "mala", "hmc"])
ap.add_argument("--data", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "pima.npz"))
ap.add_argument("--iters", type=int, default=10000)
ap.add_argument("--thin", type=int, default=0, help="0 = the reference script's value")
ap.add_argument("--out", default="")
ap.add_argument("--rng", default="philox", choices=["philox", "numpy"])
args = ap.parse_args()

X, y = load_pima(args.data)                                   # fit-numpy.py:12-19
n, p = X.shape
pscale = np.array([10.] + [1.] * (p - 1))                     # fit-np-ul.py:31
prob = bind_data(X, y, pscale)                                # the one extra line

init = np.random.randn(p) * 0.1                               # fit-numpy.py:26
print("MAP:")
res = map_estimate(prob, init, method="BFGS")                 # fit-np-ul.py:54 (method="newton": fit-jax.py:62-79 on the device)
print(res.x, ll(res.x), glp(res.x))

pre = np.array([100., 1., 1., 1., 1., 1., 25., 1.])            # fit-np-mala.py:97
if args.sampler == "rwmh":                                    # fit-numpy.py:81-86
    kernel, thin = mhKernel(lpost, RandomWalk(0.02 * np.array([10., 1., 1., 1., 1., 1., 5., 1.]))), 1000
elif args.sampler == "ul":                                    # fit-np-ul.py:88
    kernel, thin = ulKernel(glp, dt=1e-6, pre=pre), 2000
elif args.sampler == "mala":                                  # fit-np-mala.py:99
    kernel, thin = malaKernel(lpost, glp, dt=1e-5, pre=pre), 1000
else:                                                         # fit-np-hmc.py:107-108
    kernel, thin = hmcKernel(lpost, glp, eps=1e-3, l=50, dmm=1 / pre), 20
out = mcmc(res.x, kernel, thin=args.thin or thin, iters=args.iters, verb=False, rng=args.rng)
print(out)
if args.out:
    save_samples(out, args.out)
summ = describe(out)                                          # fit-numpy.py:92-96
print("Posterior summaries:")
print("Mean: " + str(summ["mean"]))
print("Variance: " + str(summ["variance"]))
print("ESS: " + str(effective_sample_size(out)))
print("accept rate:", prob.last_accept_rate)
