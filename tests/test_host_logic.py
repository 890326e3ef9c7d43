"""Host-side logic that needs no GPU: row sharding, argument normalisation, the
Philox specification, RNG pre-draw order, kernel construction."""
import numpy as np
import pytest

import logreg_b200 as lr
from logreg_b200 import _native as N
from logreg_b200 import api, dist


def test_shard_rows_partition():
    for n in (0, 1, 7, 200, 10**8, 10**8 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [dist.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dist.shard_rows(10, 3, 3)


def test_vec_broadcast_like_reference_defaults():
    # the reference's defaults are scalars (pre=1, dmm=1): fit-np-mala.py:72, fit-np-hmc.py:65
    np.testing.assert_array_equal(api._vec(1, 4, "pre"), np.ones(4))
    np.testing.assert_array_equal(api._vec([1, 2, 3], 3, "pre"), [1., 2., 3.])
    with pytest.raises(ValueError):
        api._vec([1, 2], 3, "pre")


def test_random_walk_descriptor_matches_reference_rprop():
    # fit-numpy.py:83-84: beta + 0.02*pre*np.random.randn(p)
    pre = np.array([10., 1., 1., 1., 1., 1., 5., 1.])
    rw = lr.RandomWalk(0.02 * pre)
    beta = np.arange(8.0)
    np.random.seed(3)
    a = rw(beta)
    np.random.seed(3)
    b = beta + 0.02 * pre * np.random.randn(8)
    np.testing.assert_array_equal(a, b)


def test_python_fallback_kernels_are_the_reference_closures():
    """With user callables (not the device-bound ones) the kernel constructors return the
    reference's host closures; checked against the oracle's restatement on a toy density."""
    from oracle import logreg_oracle as O
    lpi = lambda b: -0.5 * np.sum(b * b)
    glpi = lambda b: -b
    x0 = np.array([0.3, -0.2, 0.1])
    for mine, ref in (
        (lambda: lr.malaKernel(lpi, glpi, dt=0.1, pre=np.array([1., 2., 3.])),
         lambda: O.mala_kernel(lpi, glpi, 3, dt=0.1, pre=np.array([1., 2., 3.]))),
    ):
        np.random.seed(1); k = mine(); x, l = x0, lpi(x0); a = []
        for _ in range(25):
            x, l = k(x, l); a.append(x)
        np.random.seed(1); k = ref(); x, l = x0, lpi(x0); b = []
        for _ in range(25):
            x, l = k(x, l); b.append(x)
        np.testing.assert_allclose(np.array(a), np.array(b), rtol=1e-12, atol=1e-14)
    np.random.seed(2); k = lr.hmcKernel(lpi, glpi, eps=0.1, l=5, dmm=np.array([1., 2., 1.])); x = x0; a = []
    for _ in range(25):
        x = k(x); a.append(x)
    np.random.seed(2); k = O.hmc_kernel(lpi, glpi, eps=0.1, l=5, dmm=np.array([1., 2., 1.])); x = x0; b = []
    for _ in range(25):
        x = k(x); b.append(x)
    np.testing.assert_allclose(np.array(a), np.array(b), rtol=1e-12, atol=1e-14)
    np.random.seed(4); k = lr.ulKernel(glpi, dt=0.1, pre=2.0); x = x0; a = []
    for _ in range(10):
        x = k(x); a.append(x)
    np.random.seed(4); k = O.ul_kernel(glpi, 3, dt=0.1, pre=2.0); x = x0; b = []
    for _ in range(10):
        x = k(x); b.append(x)
    np.testing.assert_allclose(np.array(a), np.array(b), rtol=1e-12)
    # mcmc drives a plain Python kernel exactly like the reference loop (both signatures)
    np.random.seed(8)
    m1 = lr.mcmc(x0, lr.malaKernel(lpi, glpi, dt=0.1), thin=3, iters=7, verb=False)
    np.random.seed(8)
    m2 = O.mcmc_threaded(x0, O.mala_kernel(lpi, glpi, 3, dt=0.1), 3, 7)
    np.testing.assert_allclose(m1, m2, rtol=1e-12)
    np.random.seed(9)
    m1 = lr.mcmc(x0, lr.ulKernel(glpi, dt=0.1), thin=2, iters=5, verb=False)
    np.random.seed(9)
    m2 = O.mcmc_plain(x0, O.ul_kernel(glpi, 3, dt=0.1), 2, 5)
    np.testing.assert_allclose(m1, m2, rtol=1e-12)


def test_mcmc_verb_output_matches_reference_format(capsys):
    lpi = lambda b: -0.5 * np.sum(b * b)
    glpi = lambda b: -b
    lr.mcmc(np.zeros(2), lr.ulKernel(glpi, dt=0.1), thin=1, iters=3, verb=True)
    out = capsys.readouterr().out
    assert out == "3 iterations\n0 1 2 \nDone.\n"   # fit-numpy.py:69-78


def philox_ref(c, k):
    """Philox4x32-10 in plain Python (Salmon et al. 2011)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = list(c); k = list(k)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert philox_ref([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox_ref([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox_ref([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_front_end_namespaces_export_the_reference_names():
    """The three import switches (fit-np-*.py, fit-np-hmc.py, fit-jax*.py) expose the names those
    scripts define; importing them needs no GPU."""
    import logreg_b200 as top
    import logreg_b200.jaxlike as J
    import logreg_b200.np_hmc as H
    for name in ("ll", "lprior", "lpost", "glp", "mhKernel", "ulKernel", "malaKernel", "hmcKernel", "mcmc", "bind_data"):
        assert callable(getattr(top, name))
    for name in ("ll", "lprior", "lpost", "glp", "mhKernel", "hmcKernel", "mcmc", "bind_data"):
        assert callable(getattr(H, name))
    for name in ("ll", "lprior", "lpost", "glp", "mhKernel", "ulKernel", "malaKernel", "hmcKernel", "mcmc", "bind_data",
                 "PRNGKey", "split", "normal", "uniform"):
        assert callable(getattr(J, name))
    import inspect
    assert list(inspect.signature(J.mcmc).parameters)[:4] == ["init", "kernel", "thin", "iters"]      # fit-jax2.py:98
    assert list(inspect.signature(H.mhKernel).parameters) == ["lpost", "rprop"]                         # fit-np-hmc.py:56
    assert list(inspect.signature(top.mhKernel).parameters)[:3] == ["lpost", "rprop", "dprop"]          # fit-numpy.py:53
    # fit-np-hmc.py's one-argument kernel on plain Python callables (no device involved)
    np.random.seed(0)
    k = H.mhKernel(lambda x: -0.5 * float(np.sum(x * x)), lambda x: x + 0.5 * np.random.randn(len(x)))
    x = np.zeros(3)
    for _ in range(20):
        x = k(x)
    assert x.shape == (3,) and np.all(np.isfinite(x))
    # keys are plain integers; split is deterministic host arithmetic
    assert J.PRNGKey(42) == 42 and J.random.split(5, 3) == J.split(5, 3)
