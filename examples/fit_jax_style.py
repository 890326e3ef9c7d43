#!/usr/bin/env python3
"""Python/fit-jax2.py (RWMH) / fit-jax-ul.py / fit-jax-mala.py / fit-jax-hmc.py with the backend switched
to logreg_b200.jaxlike: same flow (data -> Newton MAP -> keyed mcmc -> parquet -> summary), same keyed
kernel signatures, no jax.

    python examples/fit_jax_style.py --sampler hmc --iters 1000
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from logreg_b200.jaxlike import (RandomWalk, bind_data, glp, hmcKernel, ll, lpost, malaKernel, mcmc,  # noqa: E402
                                 mhKernel, random, ulKernel)
from logreg_b200.workflow import describe, load_pima, map_estimate, save_samples  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--data", default=os.path.join(ROOT, "tests", "golden", "pima.npz"))
ap.add_argument("--sampler", default="rwmh", choices=["rwmh", "ul", "mala", "hmc"])
ap.add_argument("--iters", type=int, default=1000)
ap.add_argument("--thin", type=int, default=0)
ap.add_argument("--out", default="")
args = ap.parse_args()

X, y = load_pima(args.data)                                    # fit-jax2.py:17-27
n, p = X.shape
pscale = np.array([10.] + [1.] * (p - 1))
prob = bind_data(X.astype(np.float32), y, pscale)              # fit-jax2.py:30-31: float32 throughout

np.random.seed(41)                                             # fit-jax2.py:37-38
init = (np.random.randn(p) * 0.1).astype(np.float32)
print("MAP:")
res = map_estimate(prob, init)                                 # fit-jax2.py:62-79, Newton with step halving, on the device
beta = res.x
print(beta, ll(beta), np.linalg.norm(glp(beta)))

pre = np.array([100., 1., 1., 1., 1., 1., 25., 1.], dtype=np.float32)
if args.sampler == "rwmh":                                     # fit-jax2.py:118-125
    kernel, thin = mhKernel(lpost, RandomWalk(0.02 * np.array([10., 1., 1., 1., 1., 1., 5., 1.]))), 1000
elif args.sampler == "ul":                                     # fit-jax-ul.py:111
    kernel, thin = ulKernel(lpost, dt=1e-6, pre=pre), 4000
elif args.sampler == "mala":                                   # fit-jax-mala.py:132
    kernel, thin = malaKernel(lpost, dt=1e-6, pre=pre), 2000
else:                                                          # fit-jax-hmc.py:148
    kernel, thin = hmcKernel(lpost, glp, eps=1e-3, l=50, dmm=1 / pre), 20
out = mcmc(beta, kernel, thin=args.thin or thin, iters=args.iters)   # root key PRNGKey(42), as in the scripts
print(out)
# one more step through the keyed callable, the way a user would drive the kernel by hand
key = random.split(random.PRNGKey(7), 2)[0]
step = kernel(key, out[-1], lpost(out[-1])) if kernel.threaded else kernel(key, out[-1])
print("one keyed step:", step[0] if kernel.threaded else step)
if args.out:
    save_samples(out.astype(np.float64), args.out)
s = describe(out)
print("Posterior summaries:\n", "\nMean: " + str(s["mean"]), "\nVariance: " + str(s["variance"]))
