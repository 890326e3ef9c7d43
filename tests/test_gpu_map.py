"""f1: the MAP optimiser on the device -- Newton with the exact Hessian X'WX and
step halving (Python/fit-jax.py:62-79), lrb_map / lrb_hessian."""
import numpy as np
import pytest

from oracle import logreg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lr():
    import logreg_b200
    assert logreg_b200.device_count() >= 1
    return logreg_b200


def numpy_hessian(X, beta, pscale):
    pr = 1 / (1 + np.exp(-X.dot(beta)))
    return X.T @ ((pr * (1 - pr))[:, None] * X) + np.diag(1 / pscale ** 2)


def numpy_newton(tgt, X, pscale, init, tol, maxit=500):
    """fit-jax.py:62-79 in NumPy (float64), the oracle of lrb_map.  lpost uses the oracle's
    overflow-free log-likelihood: the device kernel never returns -inf (hard
    part 8), and from a far-off start the halving decisions must be taken on the same numbers."""
    lpost = lambda b: O.stable_ll(X, tgt.y, b) + tgt.lprior(b)
    beta = init.copy()
    its = 0
    for its in range(1, maxit + 1):
        g = tgt.glp(beta)
        step = np.linalg.solve(numpy_hessian(X, beta, pscale), g)
        for _ in range(15):
            if lpost(beta + step) > lpost(beta):
                break
            step = step / 2
        beta = beta + step
        if np.linalg.norm(g) < tol:
            break
    return beta, its


def numpy_mode(tgt, X, pscale, start):
    """The exact mode: full Newton steps in float64 until the step vanishes.  (No halving rule: near
    the mode an improvement of lpost smaller than its own rounding is invisible to `>`, which is why
    the reference stops at ||g|| < 0.01 and why tests at large n must not ask its loop for more.)"""
    beta = start.copy()
    for _ in range(50):
        step = np.linalg.solve(numpy_hessian(X, beta, pscale), tgt.glp(beta))
        beta = beta + step
        if np.linalg.norm(step) < 1e-13 * max(1.0, np.linalg.norm(beta)):
            break
    return beta


@pytest.mark.parametrize("mode,n,p", [("fp64", 200, 8), ("fp32", 50_003, 64), ("fp64", 20_001, 13), ("fp64", 9_001, 200),
                                      ("fp32", 30_000, 100)])
def test_hessian_equals_numpy(lr, pima, mode, n, p):
    if n == 200:
        prob = lr.Problem().bind_data(np.asfortranarray(pima["X"]), pima["y"], pima["pscale"], mode=mode)
        X, ps, b = pima["X"], pima["pscale"], pima["B"][1]
    else:
        prob = lr.Problem()
        bt = prob.gen_synthetic(n, p, mode=mode, seed=4)
        X, _ = prob.copy_rows(0, n)
        ps, b = prob.pscale, bt + 0.05 * np.random.RandomState(2).randn(p)
    H = prob.hessian(b)
    ref = numpy_hessian(X, b, ps)
    np.testing.assert_allclose(H, H.T, rtol=0, atol=0)                  # mirrored blocks
    assert np.max(np.abs(H - ref)) <= 1e-11 * np.max(np.abs(ref))       # float64 products and sums in both modes
    assert np.all(np.linalg.eigvalsh(H) > 0)


def test_pima_map_by_device_newton(lr, pima):
    """From the scripts' own starting point (np.random.seed(41); randn(p)*0.1, fit-jax.py:37-38):
    Appendix B's MAP to 1e-8 in at most 10 Newton iterations."""
    from logreg_b200.workflow import map_estimate
    X = np.asfortranarray(pima["X"])
    prob = lr.Problem().bind_data(X, pima["y"], pima["pscale"])
    tgt = O.Target(X, pima["y"], pima["pscale"])
    np.random.seed(41)
    init = np.random.randn(8) * 0.1
    e0 = prob.info()["eval_launches"]
    beta, info = prob.map(init, tol=1e-9, maxit=50)
    assert info["converged"] == 1 and info["iterations"] <= 10
    np.testing.assert_allclose(beta, pima["map"], rtol=1e-5, atol=1e-6)       # the fixture is a BFGS optimum (gtol 1e-5)
    np.testing.assert_allclose(beta, numpy_mode(tgt, X, pima["pscale"], beta), rtol=1e-8, atol=1e-10)   # Appendix B to 1e-8
    assert info["lpost"] == pytest.approx(-100.44943693563214, rel=1e-11)
    assert np.max(np.abs(tgt.glp(beta))) < 1e-8                               # Newton lands on the exact mode
    assert prob.info()["eval_launches"] - e0 == info["evals"]
    ref, its = numpy_newton(tgt, X, pima["pscale"], init, 1e-9)
    np.testing.assert_allclose(beta, ref, rtol=1e-8, atol=1e-10)
    assert abs(info["iterations"] - its) <= 1
    # the reference's own stopping rule (||g|| < 0.01, fit-jax.py:76) through the workflow helper
    res = map_estimate(prob, init)
    assert res.success and res.method == "newton" and res.nit <= its
    np.testing.assert_allclose(res.x, pima["map"], rtol=1e-4, atol=1e-5)
    # BFGS with the hand gradient (fit-np-ul.py:54) finds the same point
    res_b = map_estimate(prob, np.array([-9.8, 0.1, 0.03, -0.005, 0.0, 0.08, 1.8, 0.04]), method="BFGS")
    np.testing.assert_allclose(res_b.x, beta, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("mode,n,p,tol", [("fp64", 100_003, 64, 1e-9), ("fp32", 200_000, 64, 1e-6), ("fp64", 30_001, 150, 1e-9)])
def test_synthetic_map_equals_exact_mode(lr, mode, n, p, tol):
    """The reference's loop with the reference's stopping rule (||g|| < 0.01) on the device, against
    the same loop in NumPy float64 on the same rows (agreement 1e-9; 1e-6 in FP32 mode, whose
    gradient has a rounding floor of about 1e-8 of its un-cancelled magnitude) and against the
    exact mode (one Newton step past ||g|| < 0.01 is within 1e-7 of it)."""
    prob = lr.Problem()
    bt = prob.gen_synthetic(n, p, mode=mode, seed=6)
    X, y = prob.copy_rows(0, n)
    tgt = O.Target(X, y, prob.pscale)
    init = np.zeros(p)
    beta, info = prob.map(init, tol=0.01, maxit=30)
    ref = numpy_mode(tgt, X, prob.pscale, beta)
    same_path, its = numpy_newton(tgt, X, prob.pscale, init, 0.01)
    assert info["converged"] == 1 and info["iterations"] <= 12 and abs(info["iterations"] - its) <= 1
    assert np.max(np.abs(beta - ref)) <= 1e-7 * max(1.0, np.max(np.abs(ref)))      # set by the stopping rule
    if info["iterations"] == its:                                                  # same path: arithmetic parity
        assert np.max(np.abs(beta - same_path)) <= tol * max(1.0, np.max(np.abs(ref)))
    assert abs(info["lpost"] - tgt.lpost(ref)) <= (1e-10 if mode == "fp64" else 1e-5) * abs(tgt.lpost(ref))
    assert np.max(np.abs(beta - bt)) < 0.2                                    # and it is the right neighbourhood


def test_step_halving_and_errors(lr, pima):
    """A start from which the full Newton step overshoots exercises the halving branch; bad
    arguments are refused."""
    X = np.asfortranarray(pima["X"])
    prob = lr.Problem().bind_data(X, pima["y"], pima["pscale"])
    tgt = O.Target(X, pima["y"], pima["pscale"])
    init = np.array([3.0, 0.5, -0.2, 0.1, 0.1, -0.3, 2.0, 0.2])
    beta, info = prob.map(init, tol=1e-6, maxit=100)
    ref, its = numpy_newton(tgt, X, pima["pscale"], init, 1e-6)
    assert info["converged"] == 1 and abs(info["iterations"] - its) <= 1
    np.testing.assert_allclose(beta, numpy_mode(tgt, X, pima["pscale"], ref), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(beta, pima["map"], rtol=1e-4, atol=1e-5)
    with pytest.raises(lr.LogregB200Error):
        prob.map(init, tol=-1.0)
    with pytest.raises((lr.LogregB200Error, ValueError)):
        lr.Problem().map(np.zeros(8))
