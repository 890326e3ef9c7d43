"""CPU oracle for the logreg hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a NumPy restatement of the algorithm in the reference's four
NumPy scripts (paths relative to /root/reference):

    Python/fit-numpy.py     RWMH          ll :23-24  lprior :37-39  lpost :43-44
                                          mhKernel :53-62  mcmc :64-79  rprop :81-84
    Python/fit-np-ul.py     Langevin      lprior :33-34  glp :45-48  ulKernel :61-68
                                          mcmc :70-84
    Python/fit-np-mala.py   MALA          mhKernel :61-70  malaKernel :72-78  mcmc :80-95
    Python/fit-np-hmc.py    HMC           mhKernel :56-63  hmcKernel :65-87  mcmc :89-103

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it, and only as the checker / timed CPU
baseline.  Nothing under `logreg_b200/` imports it; the product path has no
CPU fallback.

Parity pinning: the reference ships NO golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against *outputs of the
reference itself*: `tests/golden/make_golden.py` lifts the reference's own
function definitions out of the scripts (by AST, unmodified) in the build
container, runs them on Pima and on a seeded synthetic problem, and commits the
inputs/outputs as `tests/golden/*.npz`.  `tests/test_oracle.py` checks this
module against those fixtures (bit-for-bit for the replayed chains) and, when
/root/reference is present, against the live lifted functions as well.

Deliberate differences from the scripts, none of which change a number:
  * the scripts close over module globals (X, y, p, init, pscale, pre); here the
    data live in a `Target` object and `p` is taken from the data;
  * random draws come from an explicit `rng` object with `randn(p)` / `rand()`
    (default: the legacy global `np.random`, which is what the scripts use), so a
    pre-drawn stream can be replayed (`ReplayRNG`);
  * `norm_logpdf` restates scipy.stats.norm.logpdf's arithmetic
    (-z^2/2 - log(sqrt(2 pi)) - log(scale)) so the oracle runs without SciPy.
"""
from __future__ import annotations

import numpy as np

_LOG_SQRT_2PI = np.log(np.sqrt(2 * np.pi))


def norm_logpdf(x, loc, scale):
    """scipy.stats.norm.logpdf(x, loc, scale) as SciPy evaluates it
    (fit-numpy.py:38-39 and fit-np-mala.py:34,78 call it)."""
    z = (np.asarray(x, dtype=np.float64) - loc) / scale
    return -(z ** 2) / 2.0 - _LOG_SQRT_2PI - np.log(scale)


class Target:
    """The script globals X (n x p, float64), y (n, float32 0/1) and pscale in
    one place (fit-np-mala.py:12-21,31)."""

    def __init__(self, X, y, pscale):
        self.X = X
        self.y = y
        self.n, self.p = X.shape
        self.pscale = np.asarray(pscale, dtype=np.float64)

    # fit-numpy.py:23-24 (identical in the other three scripts). Naive form:
    # overflows to -inf when -(2y-1)*eta > ~709, exactly like the reference.
    def ll(self, beta):
        eta = self.X.dot(beta)
        return np.sum(-np.log(1 + np.exp(-(2 * self.y - 1) * eta)))

    # fit-np-ul.py:33-34 (vector pscale); fit-numpy.py:37-39 is the same number
    # for pscale = [10,1,...,1].
    def lprior(self, beta):
        return np.sum(norm_logpdf(beta, 0, self.pscale))

    # fit-numpy.py:43-44
    def lpost(self, beta):
        return self.ll(beta) + self.lprior(beta)

    # fit-np-ul.py:45-48
    def glp(self, beta):
        glpr = -beta / (self.pscale * self.pscale)
        gll = (self.X.T).dot(self.y - 1 / (1 + np.exp(-self.X.dot(beta))))
        return glpr + gll

    # -- row-chunked evaluation (ll and X'(y-p) are row sums): used for sizes
    #    whose temporaries would not fit, and for row-shard tests.
    def ll_gll_rows(self, beta, lo, hi):
        Xc = self.X[lo:hi]
        yc = self.y[lo:hi]
        eta = Xc.dot(beta)
        llc = np.sum(-np.log(1 + np.exp(-(2 * yc - 1) * eta)))
        gllc = (Xc.T).dot(yc - 1 / (1 + np.exp(-eta)))
        return llc, gllc

    def lpost_glp_chunked(self, beta, chunk=1 << 20):
        llt = 0.0
        g = np.zeros(self.p)
        for lo in range(0, self.n, chunk):
            a, b = self.ll_gll_rows(beta, lo, min(self.n, lo + chunk))
            llt += a
            g += b
        return llt + self.lprior(beta), g - beta / (self.pscale * self.pscale)


class ReplayRNG:
    """Supplies pre-drawn N(0,1) rows and U(0,1) scalars in the order the
    reference kernels consume them: `randn(p)` takes the next
    row of Z, `rand()` the next entry of U."""

    def __init__(self, Z, U=None):
        self.Z = np.asarray(Z, dtype=np.float64)
        self.U = None if U is None else np.asarray(U, dtype=np.float64)
        self.iz = 0
        self.iu = 0

    def randn(self, p):
        z = self.Z[self.iz]
        assert z.shape[0] == p
        self.iz += 1
        return z

    def rand(self):
        u = self.U[self.iu]
        self.iu += 1
        return u


def predraw(rng, steps, p, uniforms=True):
    """Draw the (Z, U) stream `steps` kernel applications would consume from
    `rng`, in the reference's order: randn(p) then rand() per step
    (fit-numpy.py:84 then :58; fit-np-hmc.py:85 then :60); UL draws no
    uniforms (fit-np-ul.py:67)."""
    Z = np.empty((steps, p))
    U = np.empty(steps) if uniforms else None
    for i in range(steps):
        Z[i] = rng.randn(p)
        if uniforms:
            U[i] = rng.rand()
    return Z, U


# ---------------------------------------------------------------- kernels

def mh_kernel(lpost, rprop, dprop=lambda new, old: 1., rng=np.random, trace=None):
    """fit-numpy.py:53-62 / fit-np-mala.py:61-70: Metropolis-Hastings step that
    threads the current log-density through. `trace`, if a list, receives
    (log_alpha, log_u, accepted) per step (test instrumentation only)."""
    def kernel(x, ll):
        prop = rprop(x)
        lp = lpost(prop)
        a = lp - ll + dprop(x, prop) - dprop(prop, x)
        lu = np.log(rng.rand())
        acc = bool(lu < a)
        if trace is not None:
            trace.append((a, lu, acc))
        if acc:
            x = prop
            ll = lp
        return x, ll
    return kernel


def mh_kernel_recompute(lpost, rprop, rng=np.random, trace=None):
    """fit-np-hmc.py:56-63: the HMC script's MH step; evaluates lpost at both
    the proposal and the current state every time."""
    def kernel(x):
        prop = rprop(x)
        a = lpost(prop) - lpost(x)
        lu = np.log(rng.rand())
        acc = bool(lu < a)
        if trace is not None:
            trace.append((a, lu, acc))
        if acc:
            x = prop
        return x
    return kernel


def rw_proposal(scale, rng=np.random):
    """fit-numpy.py:81-84: beta + 0.02*pre*randn(p), with scale = 0.02*pre."""
    scale = np.asarray(scale, dtype=np.float64)
    return lambda beta: beta + scale * rng.randn(len(beta))


def ul_kernel(glpi, p, dt=1e-4, pre=1, rng=np.random):
    """fit-np-ul.py:61-68."""
    sdt = np.sqrt(dt)
    spre = np.sqrt(pre)
    advance = lambda x: x + 0.5 * pre * glpi(x) * dt
    def kernel(x):
        return advance(x) + rng.randn(p) * spre * sdt
    return kernel


def mala_kernel(lpi, glpi, p, dt=1e-4, pre=1, rng=np.random, trace=None):
    """fit-np-mala.py:72-78."""
    sdt = np.sqrt(dt)
    spre = np.sqrt(pre)
    advance = lambda x: x + 0.5 * pre * glpi(x) * dt
    return mh_kernel(
        lpi,
        lambda x: advance(x) + rng.randn(p) * spre * sdt,
        lambda new, old: np.sum(norm_logpdf(new, advance(old), spre * sdt)),
        rng=rng, trace=trace)


def hmc_kernel(lpi, glpi, eps=1e-4, l=10, dmm=1, rng=np.random, trace=None):
    """fit-np-hmc.py:65-87."""
    sdmm = np.sqrt(dmm)

    def leapf(q, p):
        p = p + 0.5 * eps * glpi(q)
        for i in range(l):
            q = q + eps * p / dmm
            if i < l - 1:
                p = p + eps * glpi(q)
            else:
                p = p + 0.5 * eps * glpi(q)
        return (q, -p)

    def alpi(x):
        (q, p) = x
        return lpi(q) - 0.5 * np.sum((p ** 2) / dmm)

    mhk = mh_kernel_recompute(alpi, lambda x: leapf(*x), rng=rng, trace=trace)

    def kern(q):
        d = len(q)
        p = rng.randn(d) * sdmm
        return mhk((q, p))[0]
    return kern


# ---------------------------------------------------------------- chain runners

def mcmc_threaded(init, kernel, thin=10, iters=10000):
    """fit-numpy.py:64-79 / fit-np-mala.py:80-95 (kernel(x, ll) -> (x, ll);
    ll starts at -inf so the first proposal is always accepted)."""
    p = len(init)
    ll = -np.inf
    mat = np.zeros((iters, p))
    x = init
    for i in range(iters):
        for j in range(thin):
            x, ll = kernel(x, ll)
        mat[i, :] = x
    return mat


def mcmc_plain(init, kernel, thin=10, iters=10000):
    """fit-np-ul.py:70-84 / fit-np-hmc.py:89-103 (kernel(x) -> x)."""
    p = len(init)
    mat = np.zeros((iters, p))
    x = init
    for i in range(iters):
        for j in range(thin):
            x = kernel(x)
        mat[i, :] = x
    return mat


# ---------------------------------------------------------------- closed forms

def mala_log_alpha(tgt, x, prop, dt, pre):
    """log acceptance ratio of one MALA move x -> prop with the normalisers
    cancelled (the reference's own closed form is
    Dex/fit-mala.dx:65-68). Equals fit-np-mala.py:64-65's `a` with ll = lpost(x)."""
    pre = np.asarray(pre, dtype=np.float64) * np.ones(tgt.p)
    adv = lambda v: v + 0.5 * pre * tgt.glp(v) * dt
    fwd = np.sum((prop - adv(x)) ** 2 / (pre * dt))
    bwd = np.sum((x - adv(prop)) ** 2 / (pre * dt))
    return tgt.lpost(prop) - tgt.lpost(x) - 0.5 * bwd + 0.5 * fwd


def stable_ll(X, y, beta):
    """ll with the overflow-free softplus the CUDA kernel uses; identical to
    Target.ll wherever the reference does not overflow."""
    z = (2 * np.asarray(y, dtype=np.float64) - 1) * X.dot(beta)
    return np.sum(np.minimum(z, 0.0) - np.log1p(np.exp(-np.abs(z))))
