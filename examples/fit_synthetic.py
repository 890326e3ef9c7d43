#!/usr/bin/env python3
"""Large-n workflow on one GPU: synthetic data generated in HBM, MAP by BFGS over the fused
evaluation, then HMC with a diagonal mass matrix -- the shape of BASELINE config 3 at a size you choose.

    python examples/fit_synthetic.py --n 10000000 --p 64 --iters 200
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import logreg_b200 as lr  # noqa: E402
from logreg_b200.workflow import describe, effective_sample_size, map_estimate  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--p", type=int, default=64)
ap.add_argument("--mode", default="fp32", choices=["fp32", "fp64"])
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--L", type=int, default=20)
args = ap.parse_args()

prob = lr.Problem()
t0 = time.perf_counter()
beta_true = prob.gen_synthetic(args.n, args.p, mode=args.mode, seed=42)
print(f"data: n={args.n} p={args.p} {args.mode} generated in {time.perf_counter() - t0:.2f} s "
      f"({prob.info()['bytes_per_eval'] / 1e9:.2f} GB per fused evaluation)")
lr.use(prob)

t0 = time.perf_counter()
res = map_estimate(prob, np.zeros(args.p), method="newton", tol=1e-6 * args.n)   # device Newton (lrb_map)
print(f"MAP: {res.nit} Newton iterations, {res.nfev} fused evaluations, {time.perf_counter() - t0:.2f} s; "
      f"|MAP - beta_true|_max = {np.max(np.abs(res.x - beta_true)):.2e}")

sd = 2.2 / np.sqrt(args.n)                      # posterior sd scale for unit-variance covariates
kernel = lr.hmcKernel(lr.lpost, lr.glp, eps=5.0 * sd / args.L, l=args.L, dmm=1.0)
t0 = time.perf_counter()
out = lr.mcmc(res.x, kernel, thin=1, iters=args.iters, verb=False)
dt = time.perf_counter() - t0
print(f"HMC: {args.iters} iterations x L={args.L} in {dt:.2f} s = {args.iters / dt:.1f} iters/s, "
      f"{args.iters * args.L / dt:.0f} fused evals/s, accept {prob.last_accept_rate:.2f}")
s = describe(out)
print("posterior mean - beta_true (first 6):", np.round((s["mean"] - beta_true)[:6] / sd, 2), "(in units of the sd scale)")
print("posterior sd / sd scale (first 6):  ", np.round(np.sqrt(s["variance"])[:6] / sd, 2))
print("ESS (first 6):", np.round(effective_sample_size(out)[:6]))
