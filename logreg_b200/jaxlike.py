"""A second front-end over the same kernels: the JAX-style, explicitly keyed API of
Python/fit-jax2.py, fit-jax-ul.py, fit-jax-mala.py and fit-jax-hmc.py.

    from logreg_b200.jaxlike import *      # instead of the jax imports and the defs
    bind_data(X, y, pscale)                # float32 storage mode by default, as the JAX scripts

Reference definitions mirrored (paths relative to the reference root):
    mhKernel(lpost, rprop, dprop)  -> kernel(key, x, ll)     Python/fit-jax2.py:87-96
    ulKernel(lpi, dt, pre)         -> kernel(key, x)         Python/fit-jax-ul.py:80-89
    malaKernel(lpi, dt, pre)       -> kernel(key, x, ll)     Python/fit-jax-mala.py:98-107
    hmcKernel(lpi, glpi, eps, l, dmm) -> kern(key, q)        Python/fit-jax-hmc.py:99-130
    mcmc(init, kernel, thin, iters)                          Python/fit-jax2.py:98-116 (no `verb`;
        the root key is PRNGKey(42) as in the scripts, or `key=`)

Keys.  A key is one unsigned 64-bit integer.  `split(key, n)[i]` is words (x, y) of
Philox4x32-10 at counter (i_lo, i_hi, 0, 4) under the parent key (`lrb_key_child` in the C ABI:
integer arithmetic, identical on host and device).  `normal(key, [p])[j]` is the device's
Box-Muller normal at Philox counter (0, 0, j, 0) under `key` and `uniform(key)` its 53-bit uniform
at counter (0, 0, 0, 1): exactly what the sampler kernels draw in LRB_RNG_KEYED mode, where
kernel application t of a run uses the key split(split(root, iters)[t // thin], thin)[t % thin]
-- the tree fit-jax2.py:100,109 builds with jax.random.split.  The streams are NOT jax's threefry
streams (the reference's own JAX and NumPy variants do not share streams either); what is kept
is the structure: same keys => same chain, independent sub-keys per step.

There is no jax here (BASELINE.json forbids JAX/XLA): all numerics run in the CUDA library,
float32 storage / float64 accumulation; returned arrays are float32 like the reference's.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from . import api as _api
from .api import DeviceKernel, Problem, RandomWalk  # noqa: F401

_M64 = (1 << 64) - 1


def PRNGKey(seed: int) -> int:
    """jax.random.PRNGKey: the root key of a run (fit-jax2.py:99 uses 42)."""
    return int(seed) & _M64


def split(key: int, num: int = 2):
    """jax.random.split: `num` independent child keys (see the module docstring)."""
    lib = N.load()
    return [int(lib.lrb_key_child(C.c_uint64(int(key) & _M64), C.c_uint64(i))) for i in range(int(num))]


def _draws(key, p):
    """(normals[p], uniform) the device generates under `key` (lrb_rng_dump at counter t=0)."""
    return _api.current().rng_dump_p(int(key) & _M64, 0, 1, p)


def normal(key, shape):
    """jax.random.normal(key, [p]) -> float32 vector: the device's draws under `key`."""
    p = int(shape[0]) if hasattr(shape, "__len__") else int(shape)
    z, _ = _draws(key, p)
    return z[0].astype(np.float32)


def uniform(key):
    """jax.random.uniform(key) -> float in (0, 1)."""
    _, u = _draws(key, 1)
    return float(u[0])


class random:  # namespace so that `jax.random.split` -> `random.split` keeps reading the same
    PRNGKey = staticmethod(PRNGKey)
    split = staticmethod(split)
    normal = staticmethod(normal)
    uniform = staticmethod(uniform)


def bind_data(X, y, pscale=None, mode="fp32", device=0):
    """The scripts' globals X, y (cast to float32 at fit-jax2.py:30-31) and the prior scales."""
    return _api.bind_data(X, y, pscale, mode=mode, device=device)


def ll(beta):
    return np.float32(_api.ll(np.asarray(beta, dtype=np.float64)))


def lprior(beta):
    return np.float32(_api.lprior(np.asarray(beta, dtype=np.float64)))


def lpost(beta):
    return np.float32(_api.lpost(np.asarray(beta, dtype=np.float64)))


def glp(beta):
    """jit(grad(lpost)) of the scripts: here the hand-coded gradient of the same fused pass."""
    return _api.glp(np.asarray(beta, dtype=np.float64)).astype(np.float32)


class KeyedKernel:
    """A transition kernel with the JAX scripts' calling convention: kernel(key, x, ll) ->
    (x, ll) for the samplers that thread the log-density (RWMH, MALA), kernel(key, x) -> x for UL
    and HMC.  mcmc() recognises it and runs the whole loop on the device in keyed mode."""

    def __init__(self, dev: DeviceKernel):
        self.dev = dev
        self.threaded = dev.threaded

    def __call__(self, key, x, ll=None):
        d = self.dev
        prob = d.problem
        x = np.asarray(x, dtype=np.float64)
        if self.threaded and ll is None:
            raise TypeError("this kernel threads the log-density: call kernel(key, x, ll)")
        # one application whose step key is `key`: the device splits it exactly as the reference
        # kernel does, so draw on the host side of the same tree and replay
        if d.sampler == N.UL:
            z, _ = prob.rng_dump_p(int(key) & _M64, 0, 1, prob.p)
            Z, U = z, None
        else:
            k0, k1 = split(key)
            if d.sampler == N.HMC:
                k1 = split(k1)[1]
            Z, _ = prob.rng_dump_p(k0, 0, 1, prob.p)
            _, U = prob.rng_dump_p(k1, 0, 1, 1)
        mat, _ = prob.run(d, x, 1, 1, Z=Z, U=U, init_lpost=(float(ll) if self.threaded else -np.inf))
        out = mat[0].astype(np.float32)
        if self.threaded:
            _, lp, _ = prob.chain_state()
            return out, np.float32(lp)
        return out


def _scale(v, p):
    return _api._vec(np.asarray(v, dtype=np.float64), p, "scale")


def mhKernel(lpost, rprop, dprop=None):
    """fit-jax2.py:87-96.  rprop is a RandomWalk(scale) descriptor (the script's
    `beta + 0.02*pre*jax.random.normal(key, [p])` with scale = 0.02*pre) for the device loop; any
    other rprop(key, x) / dprop(new, old) callables give the reference's host-side closure around
    the device lpost."""
    prob = _api._owner(lpost, "lpost") or (_api.current() if lpost is globals().get("lpost") else None)
    if prob is not None and isinstance(rprop, RandomWalk) and dprop is None:
        return KeyedKernel(DeviceKernel(prob, N.RWMH, rprop.scale))
    if dprop is None:
        dprop = lambda new, old: 1.

    def kernel(key, x, ll):
        key0, key1 = split(key)
        prop = rprop(key0, x)
        lp = lpost(prop)
        a = lp - ll + dprop(x, prop) - dprop(prop, x)
        if np.log(uniform(key1)) < a:
            return prop, lp
        return x, ll
    return kernel


def _problem_of(fn, name):
    prob = _api._owner(fn, name)
    if prob is None and fn is globals().get(name):
        prob = _api.current()
    return prob


def ulKernel(lpi, dt=1e-4, pre=1):
    """fit-jax-ul.py:80-89 (takes the log-density; its gradient is the library's glp)."""
    prob = _problem_of(lpi, "lpost")
    if prob is None:
        raise TypeError("jaxlike.ulKernel needs this module's lpost (the gradient comes from the device kernel)")
    return KeyedKernel(DeviceKernel(prob, N.UL, _scale(pre, prob.p), step=float(dt)))


def malaKernel(lpi, dt=1e-4, pre=1):
    """fit-jax-mala.py:98-107."""
    prob = _problem_of(lpi, "lpost")
    if prob is None:
        raise TypeError("jaxlike.malaKernel needs this module's lpost (the gradient comes from the device kernel)")
    return KeyedKernel(DeviceKernel(prob, N.MALA, _scale(pre, prob.p), step=float(dt)))


def hmcKernel(lpi, glpi, eps=1e-4, l=10, dmm=1):
    """fit-jax-hmc.py:99-130."""
    prob = _problem_of(lpi, "lpost")
    if prob is None or _problem_of(glpi, "glp") is not prob:
        raise TypeError("jaxlike.hmcKernel needs this module's lpost and glp")
    return KeyedKernel(DeviceKernel(prob, N.HMC, _scale(dmm, prob.p), step=float(eps), l=int(l)))


def mcmc(init, kernel, thin=10, iters=10000, key=None):
    """fit-jax2.py:98-116 / fit-jax-ul.py:91-107: (iters, p) float32 matrix of thinned states.
    The root key is PRNGKey(42), as hard-coded in the scripts, unless `key` is given."""
    root = PRNGKey(42) if key is None else int(key) & _M64
    init = np.asarray(init, dtype=np.float64)
    if isinstance(kernel, KeyedKernel):
        prob = kernel.dev.problem
        mat, acc = prob.run(kernel.dev, init, int(thin), int(iters), seed=root, keyed=True)
        prob.last_accepted = acc
        prob.last_accept_rate = acc / max(1, thin * iters)
        return mat.astype(np.float32)
    # user-supplied Python kernel: the reference's scan, stepping the callable
    import inspect
    threaded = len(inspect.signature(kernel).parameters) >= 3
    keys = split(root, iters)
    x, llv = init, -np.inf
    mat = np.zeros((iters, len(init)), dtype=np.float32)
    for i in range(iters):
        for k in split(keys[i], thin):
            if threaded:
                x, llv = kernel(k, x, llv)
            else:
                x = kernel(k, x)
        mat[i, :] = x
    return mat


__all__ = ["PRNGKey", "split", "normal", "uniform", "random", "bind_data", "ll", "lprior", "lpost", "glp",
           "mhKernel", "ulKernel", "malaKernel", "hmcKernel", "mcmc", "RandomWalk", "KeyedKernel"]
