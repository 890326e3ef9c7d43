"""Ad-hoc timing of the fused kernel (development aid, not the bench contract)."""
import sys, time, ctypes as C
import numpy as np
import torch
sys.path.insert(0, ".")
import logreg_b200 as lr
from logreg_b200 import _native as N

def time_evals(prob, reps=20, grad=True):
    p = prob.p
    s = torch.cuda.Stream()
    d_beta = torch.zeros(p, dtype=torch.float64, device="cuda")
    d_out = torch.zeros(p + 3, dtype=torch.float64, device="cuda")
    prob.set_stream(s.cuda_stream)
    with torch.cuda.stream(s):
        for _ in range(3):
            prob._ck(prob._lib.lrb_eval_device(prob._h, d_beta.data_ptr(), d_out.data_ptr(), int(grad)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.synchronize()
        e0.record(s)
        for _ in range(reps):
            prob._ck(prob._lib.lrb_eval_device(prob._h, d_beta.data_ptr(), d_out.data_ptr(), int(grad)))
        e1.record(s)
        s.synchronize()
    prob.set_stream(None)
    return e0.elapsed_time(e1) / reps

if __name__ == "__main__" and "--latency" not in sys.argv:
    torch.cuda.init()
    for mode, n, p in [("fp32", 20_000_000, 64), ("fp32", 100_000_000, 64), ("fp32", 1_000_000, 32),
                       ("fp64", 20_000_000, 64), ("fp64", 10_000_000, 128), ("fp32", 40_000_000, 32), ("fp32", 20_000_000, 128)]:
        prob = lr.Problem()
        t0 = time.time(); prob.gen_synthetic(n, p, mode=mode); tg = time.time() - t0
        inf = prob.info()
        for grad in (True, False):
            ms = time_evals(prob, grad=grad)
            print(f"{mode} n={n} p={p} grad={grad} grid={inf['grid']} gen={tg:.2f}s  {ms:.4f} ms/eval  "
                  f"{inf['bytes_per_eval'] / ms / 1e6:.1f} GB/s", flush=True)
        prob.close()


def latency_probe():
    """Fixed per-launch cost: tiny n, eval-only loop and sampler (graph) loop."""
    import numpy as np
    for mode, n, p in [("fp32", 1000, 64), ("fp32", 100_000, 64), ("fp32", 1_000_000, 32), ("fp32", 1_000_000, 64)]:
        prob = lr.Problem()
        bt = prob.gen_synthetic(n, p, mode=mode)
        ms = time_evals(prob, reps=200)
        for name, kern in (("hmc", lr.hmcKernel(prob.lpost, prob.glp, eps=1e-4, l=20, dmm=1.0)),
                           ("mala", lr.malaKernel(prob.lpost, prob.glp, dt=1e-6, pre=1.0)),
                           ("rwmh", lr.mhKernel(prob.lpost, lr.RandomWalk(1e-4 * np.ones(p))))):
            prob.run(kern, bt, 1, 50, seed=1)
            t0 = time.perf_counter(); prob.run(kern, None, 1, 2000 if name != "hmc" else 200, seed=1); dt = time.perf_counter() - t0
            ev = 2000 if name != "hmc" else 200 * 20
            print(f"latency {mode} n={n} p={p}: eval-only {ms*1e3:.1f} us | {name} graph loop {dt/ev*1e6:.1f} us/eval", flush=True)
        # many-chain SIMT kernel
        for C in (4, 16, 64):
            B = np.tile(bt, (C, 1))
            prob.eval_many(B)
            t0 = time.perf_counter(); prob.eval_many(B); dt = time.perf_counter() - t0
            print(f"   many-chain C={C}: {dt*1e3:.3f} ms per all-chain eval = {dt/C*1e6:.1f} us/chain-eval (host-timed)", flush=True)
        prob.close()

if __name__ == "__main__" and "--latency" in sys.argv:
    latency_probe()
