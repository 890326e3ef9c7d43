"""Development check of the tcgen05 many-chain kernel: eta of tile 0, then full parity + timing."""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
import logreg_b200 as lr
from logreg_b200 import _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 64
Cn = int(sys.argv[3]) if len(sys.argv) > 3 else 128
prob = lr.Problem()
bt = prob.gen_synthetic(n, p, mode="fp32", seed=42)
rs = np.random.RandomState(1)
B = bt + 0.1 * rs.randn(Cn, p)
X, y = prob.copy_rows(0, min(n, 4096))
groups = (Cn + 127) // 128
TR = prob._lib.lrb_tc_tile_rows()
eta = np.zeros((groups * 128, TR), dtype=np.float32)
prob._ck(prob._lib.lrb_debug_tc_eta(prob._h, N.as_dp(np.ascontiguousarray(B)), Cn, eta.ctypes.data_as(C.POINTER(C.c_float))))
ref = (X[:TR] @ B.T).T            # [chain][row]
err = np.abs(eta[:Cn, :min(n, TR)] - ref[:, :min(n, TR)])
print("eta tile0: max abs err", err.max(), "max |eta|", np.abs(ref).max(), "rel", err.max() / np.abs(ref).max(), flush=True)
if err.max() > 1e-3:
    print("eta sample (dev):", eta[0, :6], "\n           (ref):", ref[0, :6])
    print("chain1 dev:", eta[1, :6], " ref:", ref[1, :6])
    # diagnose: is it a transposition / permutation?
    print("corr of dev[0] with ref rows:", [float(np.corrcoef(eta[0], ref[c])[0, 1]) for c in range(3)])
t0 = time.perf_counter(); lp, l, g = prob.eval_many(B); dt = time.perf_counter() - t0
worst_lp = worst_g = 0.0
for i in range(0, Cn, max(1, Cn // 8)):
    prob._cache_key = None
    lp1, l1, g1 = prob.eval(B[i])
    worst_lp = max(worst_lp, abs(lp[i] - lp1) / abs(lp1))
    U = np.abs(g1).max() + np.sqrt(n)
    worst_g = max(worst_g, np.abs(g[i] - g1).max() / U)
    if i == 0:
        print("chain0: lp tc", lp[0], "single", lp1, "| g tc", g[0][:4], "single", g1[:4], flush=True)
print(f"parity vs single-chain fp32 kernel: lpost rel {worst_lp:.3e}, glp rel(to |g|max+sqrt n) {worst_g:.3e}")
for rep in range(3):
    t0 = time.perf_counter(); prob.eval_many(B); dt = time.perf_counter() - t0
print(f"all-chain eval C={Cn} n={n} p={p}: {dt*1e3:.3f} ms host-timed -> {dt/Cn*1e6:.2f} us per chain-eval, "
      f"{4*n*p*Cn/dt/1e12:.2f} TFLOP/s (1-pass flops)")

# device-side timing of the all-chain evaluation + a lock-step MALA run
import torch
def dev_time(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
sd = 2.2 / np.sqrt(n)
k = lr.malaKernel(prob.lpost, prob.glp, dt=(0.5 * sd) ** 2, pre=1.0)
inits = np.tile(bt, (Cn, 1))
prob.run_chains(k, inits, 1, 5, seed=3)
for rep in range(3):
    t0 = time.perf_counter(); mats, acc = prob.run_chains(k, inits, 1, 50, seed=3); dt = time.perf_counter() - t0
    print(f"lock-step MALA C={Cn}: 50 iters in {dt*1e3:.2f} ms -> {dt/50*1e3:.3f} ms per all-chain step, "
          f"{Cn*50/dt:.0f} chain-iters/s, accept {acc.mean()/50:.2f}, {4*n*p*Cn*51/dt/1e12:.1f} TFLOP/s")
