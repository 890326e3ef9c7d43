"""GPU parity of the on-device samplers against the reference chains.

The golden fixtures hold chains produced by the reference's own mcmc/kernels under
np.random.seed(s) together with the (Z, U) stream they consumed; the device
replays that stream (LRB_RNG_REPLAY). BASELINE.json north_star: "the accept/reject
sequence must be identical except where |log alpha - log u| falls inside the
stated tolerance".

Two kinds of comparison:
  * whole replayed chains (RWMH, UL, HMC, scalar-pre MALA): accept sequences must match
    exactly and states to 1e-7;
  * MALA with the reference's Pima tuning (pre=[100,..,25,..], dt=1e-5) is a CHAOTIC map:
    0.5*dt*pre*Hessian has eigenvalues > 2, so a 1e-16 perturbation (e.g. the BLAS
    summation order: the oracle itself diverges from the reference chain when X is merely
    stored row-major instead of column-major) grows ~4x per accepted step and flips a
    decision after a few dozen steps.  There the parity statement is per step: starting
    from EVERY state of the reference trajectory, the device's next state and decision
    equal the reference's (ties |log alpha - log u| < 1e-9 excepted, none occur), plus a
    short whole-chain prefix.
"""
import numpy as np
import pytest

from oracle import logreg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lr():
    import logreg_b200
    assert logreg_b200.device_count() >= 1
    return logreg_b200


def make_kernel(lr, prob, pima, kind):
    pre = pima["pre"]
    if kind == "rwmh":
        return lr.mhKernel(prob.lpost, lr.RandomWalk(0.02 * pima["pre_rw"]))
    if kind == "mala":
        return lr.malaKernel(prob.lpost, prob.glp, dt=1e-5, pre=pre)
    if kind == "mala_scalar":
        return lr.malaKernel(prob.lpost, prob.glp, dt=1e-6)
    if kind == "ul":
        return lr.ulKernel(prob.glp, dt=1e-6, pre=pre)
    if kind == "hmc":
        return lr.hmcKernel(prob.lpost, prob.glp, eps=1e-3, l=50, dmm=1 / pre)
    if kind == "hmc_l7":
        return lr.hmcKernel(prob.lpost, prob.glp, eps=2e-3, l=7, dmm=1 / pre)
    raise AssertionError(kind)


CHAINS = [("rwmh_t1", "rwmh"), ("rwmh_t50", "rwmh"), ("ul_t1", "ul"), ("ul_t40", "ul"),
          ("mala_t1", "mala"), ("mala_t25", "mala"), ("mala_scalar_t1", "mala_scalar"),
          ("hmc_t1", "hmc"), ("hmc_t5", "hmc"), ("hmc_l7_t1", "hmc_l7")]


@pytest.mark.parametrize("tag,kind", CHAINS)
def test_replayed_reference_chains(lr, pima, tag, kind):
    prob = lr.Problem().bind_data(np.asfortranarray(pima["X"]), pima["y"], pima["pscale"], mode="fp64")
    k = make_kernel(lr, prob, pima, kind)
    assert isinstance(k, lr.DeviceKernel)
    seed, thin, iters = (int(v) for v in pima[tag + "_cfg"])
    U = pima[tag + "_U"]
    mat, acc = prob.run(k, pima["chain_init"], thin, iters, Z=pima[tag + "_Z"], U=U if U.size else None)
    ref = pima[tag + "_mat"]
    assert mat.shape == ref.shape
    if kind == "mala":
        # chaotic tuning (see module docstring): whole-chain parity only on a prefix
        m = 25 if thin == 1 else 1
        np.testing.assert_allclose(mat[:m], ref[:m], rtol=1e-6, atol=1e-6)
        return
    if thin == 1 and kind != "ul":
        prev = np.vstack([pima["chain_init"][None, :], ref[:-1]])
        ref_acc = np.any(ref != prev, axis=1)
        prev = np.vstack([pima["chain_init"][None, :], mat[:-1]])
        my_acc = np.any(mat != prev, axis=1)
        np.testing.assert_array_equal(my_acc, ref_acc)
        assert acc == int(ref_acc.sum())
    np.testing.assert_allclose(mat, ref, rtol=1e-7, atol=1e-7)


def test_mala_one_step_ahead_along_reference_trajectory(lr, pima):
    """Per-step parity for the chaotic Pima MALA tuning: from each reference state, with the
    reference's draws, the device takes the reference's decision and lands on its next state."""
    X = np.asfortranarray(pima["X"])
    prob = lr.Problem().bind_data(X, pima["y"], pima["pscale"], mode="fp64")
    tgt = O.Target(X, pima["y"], pima["pscale"])
    k = lr.malaKernel(prob.lpost, prob.glp, dt=1e-5, pre=pima["pre"])
    ref, Z, U = pima["mala_t1_mat"], pima["mala_t1_Z"], pima["mala_t1_U"]
    x_prev, ll_prev = pima["chain_init"], -np.inf
    flips = 0
    for i in range(400):
        mat, acc = prob.run(k, x_prev, 1, 1, Z=Z[i:i + 1], U=U[i:i + 1], init_lpost=ll_prev)
        ref_acc = bool(np.any(ref[i] != x_prev))
        if bool(acc) != ref_acc:
            flips += 1   # would only be legitimate on a numerical tie of the accept test
        else:
            np.testing.assert_allclose(mat[0], ref[i], rtol=1e-9, atol=1e-9)
        x_prev = ref[i]
        ll_prev = tgt.lpost(x_prev)
    assert flips == 0


def test_mcmc_numpy_rng_reproduces_reference(lr, pima):
    """mcmc(..., rng='numpy') consumes the global NumPy RNG in the reference's order:
    seeding it gives the reference's chain (the script-level drop-in check, config 1)."""
    lr.bind_data(np.asfortranarray(pima["X"]), pima["y"], pima["pscale"])
    for tag, kern in (("rwmh_t50", lambda: lr.mhKernel(lr.lpost, lr.RandomWalk(0.02 * pima["pre_rw"]))),
                      ("mala_t25", lambda: lr.malaKernel(lr.lpost, lr.glp, dt=1e-5, pre=pima["pre"])),
                      ("hmc_t5", lambda: lr.hmcKernel(lr.lpost, lr.glp, eps=1e-3, l=50, dmm=1 / pima["pre"])),
                      ("ul_t40", lambda: lr.ulKernel(lr.glp, dt=1e-6, pre=pima["pre"]))):
        seed, thin, iters = (int(v) for v in pima[tag + "_cfg"])
        np.random.seed(seed)
        mat = lr.mcmc(pima["chain_init"], kern(), thin=thin, iters=iters, verb=False, rng="numpy")
        m = 1 if tag.startswith("mala") else iters   # Pima MALA tuning is chaotic: see module docstring
        np.testing.assert_allclose(mat[:m], pima[tag + "_mat"][:m], rtol=1e-6, atol=1e-6)


def test_synthetic_mala_chain_both_modes(lr, synth):
    Xd = synth["X32"].astype(np.float64)
    for mode, tol in (("fp64", 1e-7), ("fp32", 2e-3)):
        prob = lr.Problem().bind_data(synth["X32"], synth["y"], synth["pscale"], mode=mode)
        k = lr.malaKernel(prob.lpost, prob.glp, dt=2e-3)
        mat, acc = prob.run(k, synth["beta_true"], 1, 200, Z=synth["mala_Z"], U=synth["mala_U"])
        ref = synth["mala_mat"]
        if mode == "fp64":
            np.testing.assert_allclose(mat, ref, rtol=tol, atol=tol)
        else:
            # float32 arithmetic may flip a decision that is a near-tie; until the first
            # flip (if any) the chains agree to float32-level error
            same = np.all(np.abs(mat - ref) < tol, axis=1)
            first_bad = len(same) if same.all() else int(np.argmin(same))
            assert first_bad >= 50, first_bad
        assert 0 < acc <= 200


def test_single_step_calls_match_reference_kernels(lr, pima):
    """kernel(x, ll) / kernel(x) one step at a time, RNG drawn from np.random in the
    reference's order."""
    X = np.asfortranarray(pima["X"])
    prob = lr.bind_data(X, pima["y"], pima["pscale"])
    tgt = O.Target(X, pima["y"], pima["pscale"])
    pre = pima["pre"]
    x0 = pima["chain_init"]
    # MALA (threaded)
    np.random.seed(5)
    k_dev = lr.malaKernel(lr.lpost, lr.glp, dt=1e-5, pre=pre)
    x, l = x0, lr.lpost(x0)
    dev = []
    for _ in range(40):
        x, l = k_dev(x, l)
        dev.append(x)
    np.random.seed(5)
    k_ref = O.mala_kernel(tgt.lpost, tgt.glp, 8, dt=1e-5, pre=pre)
    x, l2 = x0, tgt.lpost(x0)
    ref = []
    for _ in range(40):
        x, l2 = k_ref(x, l2)
        ref.append(x)
    np.testing.assert_allclose(np.array(dev), np.array(ref), rtol=1e-8, atol=1e-8)
    assert l == pytest.approx(l2, rel=1e-10)
    # HMC (not threaded)
    np.random.seed(6)
    k_dev = lr.hmcKernel(lr.lpost, lr.glp, eps=1e-3, l=12, dmm=1 / pre)
    np.random.seed(6)
    xd = x0
    for _ in range(10):
        xd = k_dev(xd)
    np.random.seed(6)
    k_ref = O.hmc_kernel(tgt.lpost, tgt.glp, eps=1e-3, l=12, dmm=1 / pre)
    xr = x0
    for _ in range(10):
        xr = k_ref(xr)
    np.testing.assert_allclose(xd, xr, rtol=1e-8, atol=1e-8)


def test_user_python_kernels_use_device_density(lr, pima):
    """A script's own rprop (fit-numpy.py:83-84) still works: mhKernel falls back to the
    reference's host closure around the device lpost."""
    X = np.asfortranarray(pima["X"])
    lr.bind_data(X, pima["y"], pima["pscale"])
    pre = pima["pre_rw"]
    def rprop(beta):
        return beta + 0.02 * pre * np.random.randn(8)
    seed, thin, iters = (int(v) for v in pima["rwmh_t50_cfg"])
    np.random.seed(seed)
    mat = lr.mcmc(pima["chain_init"], lr.mhKernel(lr.lpost, rprop), thin=thin, iters=10, verb=False)
    np.testing.assert_allclose(mat, pima["rwmh_t50_mat"][:10], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("kind", ["rwmh", "ul", "mala", "hmc_l7"])
def test_philox_chain_matches_oracle_with_dumped_draws(lr, pima, kind):
    """Device RNG path: dump the Philox stream, feed it to the oracle's reference kernels."""
    X = np.asfortranarray(pima["X"])
    prob = lr.Problem().bind_data(X, pima["y"], pima["pscale"])
    tgt = O.Target(X, pima["y"], pima["pscale"])
    k = make_kernel(lr, prob, pima, kind)
    thin, iters, seed = 3, 40, 987654321
    mat, acc = prob.run(k, pima["chain_init"], thin, iters, seed=seed)
    Z, U = prob.rng_dump(seed, 0, thin * iters)
    assert abs(Z.mean()) < 0.15 and abs(Z.std() - 1) < 0.1 and 0 < U.min() and U.max() < 1
    rng = O.ReplayRNG(Z, U)
    pre = pima["pre"]
    if kind == "rwmh":
        ref = O.mcmc_threaded(pima["chain_init"], O.mh_kernel(tgt.lpost, O.rw_proposal(0.02 * pima["pre_rw"], rng), rng=rng), thin, iters)
    elif kind == "ul":
        ref = O.mcmc_plain(pima["chain_init"], O.ul_kernel(tgt.glp, 8, dt=1e-6, pre=pre, rng=rng), thin, iters)
    elif kind == "mala":
        ref = O.mcmc_threaded(pima["chain_init"], O.mala_kernel(tgt.lpost, tgt.glp, 8, dt=1e-5, pre=pre, rng=rng), thin, iters)
    else:
        ref = O.mcmc_plain(pima["chain_init"], O.hmc_kernel(tgt.lpost, tgt.glp, eps=2e-3, l=7, dmm=1 / pre, rng=rng), thin, iters)
    np.testing.assert_allclose(mat, ref, rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("kind", ["rwmh", "ul", "mala", "hmc_l7"])
def test_continued_run_equals_one_run(lr, pima, kind):
    prob = lr.Problem().bind_data(pima["X"], pima["y"], pima["pscale"])
    k = make_kernel(lr, prob, pima, kind)
    full, acc_full = prob.run(k, pima["chain_init"], 4, 30, seed=77)
    a, _ = prob.run(k, pima["chain_init"], 4, 12, seed=77)
    b, acc_b = prob.run(k, None, 4, 18, seed=77)
    np.testing.assert_array_equal(np.vstack([a, b]), full)
    assert acc_b == acc_full
    x, lp, t = prob.chain_state()
    np.testing.assert_array_equal(x, full[-1])
    assert t == 120


def test_mcmc_philox_posterior_close_to_reference_sampler(lr, pima):
    """Config 1 acceptance: posterior means / sds on Pima within Monte Carlo error of the
    reference sampler (here: the oracle's HMC, same tuning, independent RNG)."""
    X = np.asfortranarray(pima["X"])
    lr.bind_data(X, pima["y"], pima["pscale"])
    pre = pima["pre"]
    np.random.seed(123)
    mat = lr.mcmc(pima["map"], lr.hmcKernel(lr.lpost, lr.glp, eps=1e-3, l=50, dmm=1 / pre),
                  thin=2, iters=3000, verb=False)
    assert 0.85 < lr.current().last_accept_rate <= 1.0
    tgt = O.Target(X, pima["y"], pima["pscale"])
    np.random.seed(321)
    ref = O.mcmc_plain(pima["map"], O.hmc_kernel(tgt.lpost, tgt.glp, eps=1e-3, l=50, dmm=1 / pre), 2, 1500)
    m1, m2 = mat.mean(0), ref.mean(0)
    s1, s2 = mat.std(0), ref.std(0)
    # effective sample sizes are O(1000): allow 5 standard errors with ESS >= 300
    assert np.all(np.abs(m1 - m2) < 5 * s2 * np.sqrt(1 / 300 + 1 / 300))
    assert np.all(np.abs(s1 / s2 - 1) < 0.35)


def test_run_argument_errors(lr, pima):
    prob = lr.Problem().bind_data(pima["X"], pima["y"], pima["pscale"])
    k = lr.hmcKernel(prob.lpost, prob.glp, eps=1e-3, l=5, dmm=1.0)
    with pytest.raises(lr.LogregB200Error, match="no paused chain"):
        prob.run(k, None, 1, 1)
    k.l = 0
    with pytest.raises(lr.LogregB200Error, match="l >= 1"):
        prob.run(k, pima["chain_init"], 1, 1)
    k2 = lr.ulKernel(prob.glp, dt=-1.0)
    with pytest.raises(lr.LogregB200Error, match="step"):
        prob.run(k2, pima["chain_init"], 1, 1)


def test_device_philox_matches_its_specification(lr, pima):
    """lrb_rng_dump against a plain-Python Philox4x32-10 + Box-Muller (the documented stream:
    counter = (t_lo, t_hi, coordinate, stream), key = seed; 53-bit uniforms)."""
    from tests.test_host_logic import philox_ref
    prob = lr.Problem().bind_data(pima["X"], pima["y"], pima["pscale"])
    seed, t0, cnt = 0x0123456789ABCDEF, (1 << 33) + 5, 6
    Z, U = prob.rng_dump(seed, t0, cnt)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    def u53(hi, lo):
        return (((hi >> 5) << 26 | (lo >> 6)) + 0.5) / 9007199254740992.0
    for i in range(cnt):
        t = t0 + i
        r = philox_ref([t & 0xFFFFFFFF, t >> 32, 0, 1], key)
        assert U[i] == u53(r[0], r[1])
        for j in range(8):
            r = philox_ref([t & 0xFFFFFFFF, t >> 32, j, 0], key)
            z = np.sqrt(-2.0 * np.log(u53(r[0], r[1]))) * np.cos(2 * np.pi * u53(r[2], r[3]))
            assert Z[i, j] == pytest.approx(z, rel=1e-12, abs=1e-14)


@pytest.mark.parametrize("mode", ["fp64", "fp32"])
def test_many_chain_evaluation_matches_single_chain(lr, synth, mode):
    """lrb_eval with C >= 2 goes through the many-chain kernel (X batch reused for 4 chains)."""
    prob = lr.Problem().bind_data(synth["X32"], synth["y"], synth["pscale"], mode=mode)
    rs = np.random.RandomState(8)
    for c in (2, 3, 4, 7, 13):
        B = synth["beta_true"] + 0.2 * rs.randn(c, 32)
        lp, l, g = prob.eval_many(B)
        for i in range(c):
            prob._cache_key = None
            lp1, l1, g1 = prob.eval(B[i])
            assert lp[i] == pytest.approx(lp1, rel=1e-13)
            assert l[i] == pytest.approx(l1, rel=1e-13)
            np.testing.assert_allclose(g[i], g1, rtol=1e-11, atol=1e-9)
    Xd = synth["X32"].astype(np.float64)
    tgt = O.Target(Xd, synth["y"], synth["pscale"])
    lp, l, g = prob.eval_many(synth["B"])
    tol = 1e-10 if mode == "fp64" else 1e-5
    np.testing.assert_allclose(lp, synth["lpost"], rtol=tol)


@pytest.mark.parametrize("kind", ["rwmh", "ul", "mala", "hmc_l7"])
def test_lockstep_chains_equal_independent_runs(lr, pima, kind):
    """C chains advanced in lock-step by the many-chain kernel == C separate single-chain runs
    with the same per-chain Philox keys (fp64: identical decisions, states to 1e-9)."""
    prob = lr.Problem().bind_data(pima["X"], pima["y"], pima["pscale"])
    k = make_kernel(lr, prob, pima, kind)
    rs = np.random.RandomState(4)
    C = 6
    inits = pima["chain_init"] + 0.01 * rs.randn(C, 8) * np.array([1., .02, .005, .005, .005, .01, .3, .01])
    seed = 424242
    mats, accs = prob.run_chains(k, inits, 3, 15, seed=seed)
    assert mats.shape == (C, 15, 8)
    for c in range(C):
        m1, a1 = prob.run(k, inits[c], 3, 15, seed=(seed + c * 0x9E3779B97F4A7C15) % 2 ** 64)
        assert a1 == accs[c]
        np.testing.assert_allclose(mats[c], m1, rtol=1e-9, atol=1e-9)


def test_lockstep_chains_replay_reference(lr, pima):
    """Replay mode with per-chain streams: every chain reproduces the reference HMC chain."""
    prob = lr.Problem().bind_data(np.asfortranarray(pima["X"]), pima["y"], pima["pscale"])
    k = make_kernel(lr, prob, pima, "hmc_l7")
    Z, U, ref = pima["hmc_l7_t1_Z"][:60], pima["hmc_l7_t1_U"][:60], pima["hmc_l7_t1_mat"][:60]
    C = 5
    mats, accs = prob.run_chains(k, np.tile(pima["chain_init"], (C, 1)), 1, 60,
                                 Z=np.tile(Z, (C, 1, 1)), U=np.tile(U, (C, 1)))
    for c in range(C):
        np.testing.assert_allclose(mats[c], ref, rtol=1e-7, atol=1e-7)


# ---------------------------------------------------------------- tensor-core (tcgen05) many-chain path
def _tc_problem(lr, n, p=64, seed=42):
    prob = lr.Problem()
    bt = prob.gen_synthetic(n, p, mode="fp32", seed=seed)
    return prob, bt


@pytest.mark.parametrize("n,C", [(64, 32), (100, 40), (8192, 128), (100_003, 130), (300_000, 512)])
def test_tensor_core_many_chain_eval_against_oracle(lr, n, C):
    """C >= 12 chains in FP32 mode with p = 64 run on the tcgen05 kernel (3xTF32): compare with
    the oracle on the same rows -- FP32-mode tolerance 1e-5 (lpost relative to itself, glp
    relative to the un-cancelled gradient magnitude). Covers ragged n (tail tile, TMA
    zero-fill) and chain counts that are not a multiple of 128."""
    from tests.helpers import ungrad_scale
    prob, bt = _tc_problem(lr, n)
    X, y = prob.copy_rows(0, n)
    tgt = O.Target(X, y, prob.pscale)
    rs = np.random.RandomState(7)
    B = bt + 0.3 * rs.randn(C, 64) / 8
    B[0] = 0.0
    lp, l, g = prob.eval_many(B)
    for c in list(range(0, C, max(1, C // 6))) + [C - 1]:
        with np.errstate(over="ignore"):
            ref_lp, ref_g = O.stable_ll(X, y, B[c]) + tgt.lprior(B[c]), tgt.glp(B[c])
        assert abs(lp[c] - ref_lp) <= 1e-5 * max(1.0, abs(ref_lp)), (c, lp[c], ref_lp)
        assert np.max(np.abs(g[c] - ref_g)) <= 1e-5 * ungrad_scale(X, y, B[c], prob.pscale), c
    if n >= 64:
        assert l[0] == pytest.approx(-n * np.log(2.0), rel=1e-6)


def test_tensor_core_eta_tile(lr):
    """The contraction itself: eta of the first row tile as the tensor cores produce it."""
    import ctypes as C
    from logreg_b200 import _native as N
    prob, bt = _tc_problem(lr, 4096)
    rows = prob._lib.lrb_tc_tile_rows()
    X, y = prob.copy_rows(0, rows)
    B = bt + 0.1 * np.random.RandomState(1).randn(256, 64)
    eta = np.zeros((256, rows), dtype=np.float32)
    prob._ck(prob._lib.lrb_debug_tc_eta(prob._h, N.as_dp(np.ascontiguousarray(B)), 256,
                                        eta.ctypes.data_as(C.POINTER(C.c_float))))
    ref = (X @ B.T).T
    assert np.max(np.abs(eta - ref)) <= 2e-6 * np.max(np.abs(X)) * np.max(np.abs(B)) * 64


def test_tensor_core_lockstep_mala_matches_simt_path(lr):
    """512 MALA chains for 30 steps on the tcgen05 path vs the same chains on the single-chain
    float32 kernel: same Philox keys; float32-level agreement until a near-tie decision."""
    prob, bt = _tc_problem(lr, 200_000)
    sd = 2.2 / np.sqrt(200_000)
    k = lr.malaKernel(prob.lpost, prob.glp, dt=(0.5 * sd) ** 2, pre=1.0)
    C = 512
    inits = bt + 0.5 * sd * np.random.RandomState(3).randn(C, 64)
    mats, accs = prob.run_chains(k, inits, 1, 30, seed=5)
    assert 0.5 < accs.mean() / 30 <= 1.0
    for c in (0, 17, 255, 511):
        m1, a1 = prob.run(k, inits[c], 1, 30, seed=(5 + c * 0x9E3779B97F4A7C15) % 2 ** 64)
        same = np.all(np.abs(mats[c] - m1) < 5e-3 * sd, axis=1)
        first_bad = len(same) if same.all() else int(np.argmin(same))
        assert first_bad >= 10, (c, first_bad)


@pytest.mark.parametrize("kind,thin,iters", [("rwmh", 200, 4000), ("ul", 400, 3000), ("mala", 200, 4000), ("hmc", 10, 3000)])
def test_pima_posterior_within_mc_error_of_reference_samplers(lr, pima, kind, thin, iters):
    """BASELINE.json config 1: posterior means / sds on Pima within Monte Carlo error of the
    REFERENCE samplers. tests/golden/pima_posterior.npz holds mean / sd / ESS of chains produced by
    the reference's own mcmc + kernels (make_posterior.py); the device chain uses the same tuning
    constants (fit-numpy.py:81-86, fit-np-ul.py:88, fit-np-mala.py:99, fit-np-hmc.py:108) with the
    on-device Philox stream. UL is biased by construction -- in the reference too -- so it is
    compared with the reference's UL, not with the exact posterior."""
    import os
    from logreg_b200.workflow import effective_sample_size
    post = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "pima_posterior.npz")))
    prob = lr.Problem().bind_data(np.asfortranarray(pima["X"]), pima["y"], pima["pscale"])
    k = make_kernel(lr, prob, pima, kind)
    mat, acc = prob.run(k, pima["map"], thin, iters, seed=20260101)
    m, s, ess = mat.mean(0), mat.std(0, ddof=1), effective_sample_size(mat)
    rm, rsd, ress = post[kind + "_mean"], post[kind + "_sd"], post[kind + "_ess"]
    se = np.sqrt(rsd ** 2 / np.maximum(ress, 4.0) + s ** 2 / np.maximum(ess, 4.0))
    z = np.abs(m - rm) / se
    assert np.all(z < 4.5), (kind, z)
    # standard deviations: relative MC error of an sd estimate ~ 1/sqrt(2*ESS) per chain
    rel = np.abs(s / rsd - 1.0)
    tol = 4.5 * np.sqrt(0.5 / np.maximum(ress, 4.0) + 0.5 / np.maximum(ess, 4.0)) + 0.05
    assert np.all(rel < tol), (kind, rel, tol)
    if kind != "ul":
        rate = acc / (thin * iters)
        lo, hi = {"rwmh": (0.02, 0.2), "mala": (0.1, 0.5), "hmc": (0.85, 1.0)}[kind]   # 0.05 / 0.26 / 0.96
        assert lo < rate <= hi, (kind, rate)


@pytest.mark.parametrize("kind,epi", [("hmc_l7", 7), ("mala", 1)])
def test_stepwise_calls_reuse_the_cached_state(lr, pima, kind, epi):
    """mcmc / kernel calls that start exactly where the previous call stopped skip the evaluation
    at init (LRB_RUN_REUSE_CACHE): same chain as a cold start from that state, one pass less."""
    X = np.asfortranarray(pima["X"])
    a = lr.Problem().bind_data(X, pima["y"], pima["pscale"])
    ka = make_kernel(lr, a, pima, kind)
    m1, _ = a.run(ka, pima["chain_init"], 1, 5, seed=1)
    e0 = a.info()["eval_launches"]
    m2, acc2 = a.run(ka, m1[-1], 1, 6, seed=2)
    warm = a.info()["eval_launches"] - e0
    b = lr.Problem().bind_data(X, pima["y"], pima["pscale"])
    kb = make_kernel(lr, b, pima, kind)
    e0 = b.info()["eval_launches"]
    m3, acc3 = b.run(kb, m1[-1], 1, 6, seed=2)
    cold = b.info()["eval_launches"] - e0
    np.testing.assert_array_equal(m2, m3)
    assert acc2 == acc3
    assert cold == 6 * epi + 1 and warm == 6 * epi
    # a different starting point must NOT reuse anything
    e0 = a.info()["eval_launches"]
    a.run(ka, m1[-1] + 1e-9, 1, 2, seed=3)
    assert a.info()["eval_launches"] - e0 == 2 * epi + 1
