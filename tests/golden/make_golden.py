#!/usr/bin/env python3
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN FUNCTIONS.

Run in the build container only (needs /root/reference, pandas, pyarrow, scipy):

    python tests/golden/make_golden.py

The reference scripts cannot be imported (hyphenated names; their top level
runs 1e7 MCMC iterations and imports matplotlib), so their function
definitions are lifted out by AST -- unmodified -- and exec'd in a namespace in
which the script globals they close over (X, y, n, p, init, pscale, pre) are
injected.  Nothing from the reference is copied into this repository: only the
numbers those functions return are stored.

The random stream is pinned by `np.random.seed(seed)`; the (Z, U) draws each
chain consumes are re-drawn from the same seed in the reference's consumption
order (randn(p) then rand() per kernel call; UL: randn(p) only) and stored so
the device sampler can replay them.
"""
import ast
import os
import sys

import numpy as np
import pandas as pd
import scipy as sp
import scipy.stats
from scipy.optimize import minimize

REF = os.environ.get("LOGREG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def lift(script, names):
    """Return {name: ast node} for top-level defs/assignments in a reference script."""
    src = open(os.path.join(REF, "Python", script)).read()
    tree = ast.parse(src)
    out = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out.append(node)
        elif isinstance(node, ast.Assign) and len(node.targets) == 1 \
                and isinstance(node.targets[0], ast.Name) and node.targets[0].id in names:
            out.append(node)
    mod = ast.Module(body=out, type_ignores=[])
    return compile(mod, script, "exec")


def namespace(script, names, **inject):
    ns = {"np": np, "sp": sp, "scipy": scipy, "pd": pd, "os": os}
    ns.update(inject)
    exec(lift(script, names), ns)
    return ns


def load_pima():
    # fit-numpy.py:12-19, executed as written (minus the prints)
    df = pd.read_parquet(os.path.join(REF, "pima.parquet"))
    n, p = df.shape
    y = pd.get_dummies(df["type"])["Yes"].to_numpy(dtype='float32')
    X = df.drop(columns="type").to_numpy()
    X = np.hstack((np.ones((n, 1)), X))
    return X, y, n, p


def predraw(seed, steps, p, uniforms=True):
    np.random.seed(seed)
    Z = np.empty((steps, p))
    U = np.empty(steps) if uniforms else np.empty(0)
    for i in range(steps):
        Z[i] = np.random.randn(p)
        if uniforms:
            U[i] = np.random.rand()
    return Z, U


def main():
    X, y, n, p = load_pima()
    assert (n, p) == (200, 8)
    g = {}
    g["X"] = np.ascontiguousarray(X)
    g["X_is_fortran"] = np.array(X.flags["F_CONTIGUOUS"])
    g["y"] = y

    init0 = np.zeros(p)
    rw = namespace("fit-numpy.py", {"ll", "lprior", "lpost", "mhKernel", "mcmc", "pre", "rprop"},
                   X=X, y=y, n=n, p=p, init=init0)
    ul = namespace("fit-np-ul.py", {"ll", "pscale", "lprior", "lpost", "glp", "ulKernel", "mcmc"},
                   X=X, y=y, n=n, p=p, init=init0)
    ma = namespace("fit-np-mala.py", {"ll", "pscale", "lprior", "lpost", "glp", "mhKernel", "malaKernel", "mcmc"},
                   X=X, y=y, n=n, p=p, init=init0)
    hm = namespace("fit-np-hmc.py", {"ll", "pscale", "lprior", "lpost", "glp", "mhKernel", "hmcKernel", "mcmc"},
                   X=X, y=y, n=n, p=p, init=init0)
    g["pscale"] = ma["pscale"]
    g["pre_rw"] = rw["pre"]

    # ---- point evaluations (SURVEY appendix B + extra points)
    b0 = np.array([-9.8, 0.1, 0.03, -0.005, 0.0, 0.08, 1.8, 0.04])
    res = minimize(lambda x: -ma["lpost"](x), b0, jac=lambda x: -ma["glp"](x), method='BFGS')
    bmap = res.x
    rs = np.random.RandomState(2024)
    pts = [np.zeros(p), b0, bmap]
    for k in range(5):
        pts.append(bmap + rs.randn(p) * np.array([1., .02, .005, .005, .005, .01, .3, .01]))
    B = np.array(pts)
    g["B"] = B
    g["ll"] = np.array([ma["ll"](b) for b in B])
    g["lprior"] = np.array([ma["lprior"](b) for b in B])
    g["lprior_fitnumpy"] = np.array([rw["lprior"](b) for b in B])
    g["lpost"] = np.array([ma["lpost"](b) for b in B])
    g["glp"] = np.array([ma["glp"](b) for b in B])
    g["map"] = bmap

    # one MALA log-alpha in the reference's own arithmetic (appendix B row)
    pre = np.array([100., 1., 1., 1., 1., 1., 25., 1.])
    g["pre"] = pre
    dt = 1e-5
    z = np.linspace(-1, 1, 8)
    adv = lambda x: x + 0.5 * pre * ma["glp"](x) * dt
    prop = adv(b0) + z * np.sqrt(pre) * np.sqrt(dt)
    dprop = lambda new, old: np.sum(sp.stats.norm.logpdf(new, loc=adv(old), scale=np.sqrt(pre) * np.sqrt(dt)))
    g["mala_prop"] = prop
    g["mala_log_alpha"] = np.array(ma["lpost"](prop) - ma["lpost"](b0) + dprop(b0, prop) - dprop(prop, b0))

    # ---- replayed chains.  thin=1 chains expose every state (=> the accept
    #      sequence); thinned chains pin the thinning semantics.
    def run(tag, seed, mcmc, kernel_factory, init, thin, iters, uniforms=True):
        np.random.seed(seed)
        mat = mcmc(init, kernel_factory(), thin=thin, iters=iters, verb=False)
        Z, U = predraw(seed, thin * iters, p, uniforms)
        g[tag + "_mat"] = mat
        g[tag + "_Z"] = Z
        g[tag + "_U"] = U
        g[tag + "_cfg"] = np.array([seed, thin, iters])

    init = bmap.copy()
    g["chain_init"] = init
    # RWMH, fit-numpy.py:86 tuning (proposal sd 0.02*pre)
    run("rwmh_t1", 11, rw["mcmc"], lambda: rw["mhKernel"](rw["lpost"], rw["rprop"]), init, 1, 1500)
    run("rwmh_t50", 12, rw["mcmc"], lambda: rw["mhKernel"](rw["lpost"], rw["rprop"]), init, 50, 40)
    # UL, fit-np-ul.py:88 tuning
    ul["init"] = init
    run("ul_t1", 21, ul["mcmc"], lambda: ul["ulKernel"](ul["glp"], dt=1e-6, pre=pre), init, 1, 600, uniforms=False)
    run("ul_t40", 22, ul["mcmc"], lambda: ul["ulKernel"](ul["glp"], dt=1e-6, pre=pre), init, 40, 25, uniforms=False)
    # MALA, fit-np-mala.py:99 tuning
    ma["init"] = init
    run("mala_t1", 31, ma["mcmc"], lambda: ma["malaKernel"](ma["lpost"], ma["glp"], dt=1e-5, pre=pre), init, 1, 800)
    run("mala_t25", 32, ma["mcmc"], lambda: ma["malaKernel"](ma["lpost"], ma["glp"], dt=1e-5, pre=pre), init, 25, 30)
    # scalar pre (the reference default pre=1)
    run("mala_scalar_t1", 33, ma["mcmc"], lambda: ma["malaKernel"](ma["lpost"], ma["glp"], dt=1e-6), init, 1, 300)
    # HMC, fit-np-hmc.py:105-108 tuning (l=50) and a short-trajectory variant
    run("hmc_t1", 41, hm["mcmc"], lambda: hm["hmcKernel"](hm["lpost"], hm["glp"], eps=1e-3, l=50, dmm=1 / pre), init, 1, 120)
    run("hmc_t5", 42, hm["mcmc"], lambda: hm["hmcKernel"](hm["lpost"], hm["glp"], eps=1e-3, l=50, dmm=1 / pre), init, 5, 20)
    run("hmc_l7_t1", 43, hm["mcmc"], lambda: hm["hmcKernel"](hm["lpost"], hm["glp"], eps=2e-3, l=7, dmm=1 / pre), init, 1, 200)

    np.savez_compressed(os.path.join(HERE, "pima.npz"), **g)
    print("pima.npz:", {k: getattr(v, "shape", None) for k, v in g.items()})

    # ---- seeded synthetic problem (n=2000, p=32): X stored as float32 so the
    #      fp32-mode and fp64-mode device paths see identical inputs.
    n2, p2 = 2000, 32
    rs = np.random.RandomState(42)
    X32 = np.hstack((np.ones((n2, 1)), rs.randn(n2, p2 - 1))).astype(np.float32)
    bt = np.random.RandomState(41).randn(p2) / np.sqrt(p2)
    Xd = X32.astype(np.float64)
    pr = 1 / (1 + np.exp(-Xd.dot(bt)))
    y2 = (rs.rand(n2) < pr).astype(np.float32)
    ps2 = np.ones(p2); ps2[0] = 10.
    sy = namespace("fit-np-mala.py", {"ll", "lprior", "lpost", "glp", "mhKernel", "malaKernel", "mcmc"},
                   X=Xd, y=y2, n=n2, p=p2, init=np.zeros(p2), pscale=ps2)
    r3 = np.random.RandomState(43)
    B2 = np.array([bt + s * r3.randn(p2) for s in (0.0, 0.01, 0.1, 0.5, 1.0)] + [np.zeros(p2)])
    s = {"X32": X32, "y": y2, "pscale": ps2, "beta_true": bt, "B": B2,
         "ll": np.array([sy["ll"](b) for b in B2]),
         "lpost": np.array([sy["lpost"](b) for b in B2]),
         "glp": np.array([sy["glp"](b) for b in B2])}
    # a MALA chain on it (scalar pre), thin 1
    init2 = bt.copy()
    sy["init"] = init2
    np.random.seed(51)
    s["mala_mat"] = sy["mcmc"](init2, sy["malaKernel"](sy["lpost"], sy["glp"], dt=2e-3), thin=1, iters=200, verb=False)
    s["mala_Z"], s["mala_U"] = predraw(51, 200, p2)
    np.savez_compressed(os.path.join(HERE, "synth2000x32.npz"), **s)
    print("synth2000x32.npz written")


if __name__ == "__main__":
    sys.exit(main())
