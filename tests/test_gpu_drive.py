"""Drive mode (eval_persist_kernel.cuh): one cooperative launch per run, dynamic row-batch
scheduling, atomic accumulation.  It must agree with the static fixed-order kernel to rounding,
with the oracle to the stated tolerances, and must never leave its counters dirty."""
import numpy as np
import pytest

from tests.helpers import ungrad_scale
from oracle import logreg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lr():
    import logreg_b200
    assert logreg_b200.device_count() >= 1
    return logreg_b200


@pytest.mark.parametrize("mode,n,p", [("fp32", 300_007, 64), ("fp64", 150_001, 32), ("fp32", 70_003, 13),
                                      ("fp64", 4_001, 200), ("fp32", 257, 8), ("fp32", 1, 8)])
def test_drive_matches_static_kernel_and_oracle(lr, mode, n, p):
    d = lr.Problem()
    bt = d.gen_synthetic(n, p, mode=mode, seed=5)
    s = lr.Problem(deterministic=True)
    s.gen_synthetic(n, p, mode=mode, seed=5)
    rs = np.random.RandomState(1)
    X, y = d.copy_rows(0, n)
    tgt = O.Target(X, y, d.pscale)
    for _ in range(3):
        b = bt + 0.05 * rs.randn(p)
        lp_d, ll_d, g_d = d.eval(b)
        lp_s, ll_s, g_s = s.eval(b)
        scale = ungrad_scale(X, y, b, d.pscale)
        # same per-batch arithmetic, different order of the float64 additions
        assert abs(lp_d - lp_s) <= 1e-12 * max(1.0, abs(lp_s))
        assert np.max(np.abs(g_d - g_s)) <= 1e-12 * scale
        tol = 1e-10 if mode == "fp64" else 1e-5
        assert abs(lp_d - tgt.lpost(b)) <= tol * max(1.0, abs(lp_d))
        assert np.max(np.abs(g_d - tgt.glp(b))) <= tol * scale
        assert d.ll(b) == pytest.approx(ll_s, rel=1e-12)      # the no-gradient variant


@pytest.mark.parametrize("kind", ["rwmh", "ul", "mala", "hmc"])
def test_drive_chain_equals_static_chain(lr, kind):
    """Whole sampler runs: one cooperative launch vs one launch per evaluation, Philox draws."""
    n, p = 120_011, 32
    probs = []
    for det in (False, True):
        q = lr.Problem(deterministic=det)
        bt = q.gen_synthetic(n, p, mode="fp64", seed=9)
        probs.append(q)
    sd = 2.2 / np.sqrt(n)
    outs = []
    for q in probs:
        k = {"rwmh": lambda: lr.mhKernel(q.lpost, lr.RandomWalk(0.3 * sd * np.ones(p))),
             "ul": lambda: lr.ulKernel(q.glp, dt=(0.3 * sd) ** 2, pre=1.0),
             "mala": lambda: lr.malaKernel(q.lpost, q.glp, dt=(0.6 * sd) ** 2, pre=1.0),
             "hmc": lambda: lr.hmcKernel(q.lpost, q.glp, eps=0.25 * sd, l=6, dmm=1.0)}[kind]()
        mat, acc = q.run(k, bt, 3, 40, seed=123)
        mat2, acc2 = q.run(k, None, 3, 25, seed=123)       # continuation
        outs.append((mat, acc, mat2, acc2, q.chain_state()))
    (m_d, a_d, m2_d, a2_d, st_d), (m_s, a_s, m2_s, a2_s, st_s) = outs
    assert a_d == a_s and a2_d == a2_s
    np.testing.assert_allclose(m_d, m_s, rtol=0, atol=1e-9 * sd * 1e3)
    np.testing.assert_allclose(m2_d, m2_s, rtol=0, atol=1e-9 * sd * 1e3)
    assert st_d[2] == st_s[2] == 3 * 65
    assert 0 < a_d <= 120


def test_drive_counters_clean_and_launch_count(lr):
    """A run is ONE kernel launch; surplus or repeated calls find the counters at rest."""
    q = lr.Problem()
    bt = q.gen_synthetic(50_000, 64, mode="fp32", seed=3)
    k = lr.hmcKernel(q.lpost, q.glp, eps=1e-3, l=5, dmm=1.0)
    i0 = q.info()
    mat, _ = q.run(k, bt, 1, 10, seed=1)
    i1 = q.info()
    assert i1["eval_launches"] - i0["eval_launches"] == 10 * 5 + 1
    assert i1["kernel_launches"] - i0["kernel_launches"] == 2          # sampler_begin + one drive launch
    # interleave single evaluations and runs: same answers every time
    lp0 = q.lpost(bt)
    for _ in range(3):
        q.run(k, bt, 1, 3, seed=2)
        assert q.lpost(bt) == pytest.approx(lp0, rel=1e-13)
    m2, _ = q.run(k, bt, 1, 10, seed=1)
    np.testing.assert_allclose(m2, mat, rtol=0, atol=1e-12)
