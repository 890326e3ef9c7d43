"""Where does the fixed per-evaluation cost go?  Per-CTA %globaltimer timeline of the fused
kernel inside the sampler loop (development aid; uses lrb_debug_timeline)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import logreg_b200 as lr
from logreg_b200 import _native as N


def run_loop(prob, kern, bt, iters):
    prob.run(kern, bt, 1, 3, seed=1)
    s = torch.cuda.Stream(); prob.set_stream(s.cuda_stream)
    sp = prob._params(kern, seed=1, rng=N.RNG_PHILOX, init_lpost=-np.inf)
    N.check(prob._lib.lrb_run_begin(prob._h, C.byref(sp), None, 1, iters, None, None), prob._h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N.check(prob._lib.lrb_run_launch(prob._h), prob._h); s.synchronize()
    e0.record(s); N.check(prob._lib.lrb_run_launch(prob._h), prob._h); e1.record(s); s.synchronize()
    ev = C.c_int64(); prob._lib.lrb_run_evals_per_launch(prob._h, C.byref(ev))
    prob.set_stream(None)
    return e0.elapsed_time(e1) * 1e3 / ev.value


def timeline(prob):
    buf = np.zeros(prob.info()["grid"] * 4 + 8, dtype=np.int64)
    g = C.c_int64()
    N.check(prob._lib.lrb_debug_timeline_read(prob._h, buf.ctypes.data_as(C.POINTER(C.c_int64)), buf.size, C.byref(g)), prob._h)
    G = g.value
    t = buf[:G * 4].reshape(G, 4).astype(np.float64)
    last = buf[G * 4:].astype(np.float64)
    t0 = t[:, 0].min()
    q = lambda a: "min %.1f p10 %.1f med %.1f p90 %.1f max %.1f" % tuple(np.percentile((a - t0) / 1e3, [0, 10, 50, 90, 100]))
    print("   grid", G, "| us since first CTA entry:")
    print("   entry      ", q(t[:, 0])); print("   after wait ", q(t[:, 1])); print("   stream end ", q(t[:, 2])); print("   ticket     ", q(t[:, 3]))
    se = np.sort(t[:, 2] - t0) / 1e3
    print("   CTAs still streaming at (max - x us): " + ", ".join(f"{x}us:{int((se > se[-1] - x).sum())}" for x in (1, 2, 5, 10, 20, 40, 80)))
    print("   last CTA: prev finish %.1f | ticket seen %.1f | reduced %.1f | finished %.1f" %
          tuple((last[i] - t0) / 1e3 for i in (4, 0, 1, 2)))


if __name__ == "__main__":
    cfgs = [(12_500_000, 64, 20, "hmc"), (1_000_000, 32, 1, "mala"), (50_000_000, 64, 20, "hmc")]
    for pdl in ("drive", "1", "0"):
        os.environ["LRB_DRIVE"] = "1" if pdl == "drive" else "0"
        os.environ["LRB_PDL"] = "1" if pdl == "drive" else pdl
        for n, p, L, samp in cfgs:
            prob = lr.Problem(); bt = prob.gen_synthetic(n, p, mode="fp32")
            sd = 2.2 / np.sqrt(n)
            kern = (lr.hmcKernel(prob.lpost, prob.glp, eps=5 * sd / L, l=L, dmm=1.0) if samp == "hmc"
                    else lr.malaKernel(prob.lpost, prob.glp, dt=(0.6 * sd) ** 2, pre=1.0))
            iters = 20 if samp == "hmc" else 2000
            us = run_loop(prob, kern, bt, iters)
            byt = prob.info()["bytes_per_eval"]
            print(f"PDL={pdl} n={n} p={p} {samp}: {us:.1f} us/eval  ({byt / us / 1e3:.0f} GB/s; streaming at 7.3 TB/s would be {byt / 7.3e6:.1f} us)", flush=True)
            N.check(prob._lib.lrb_debug_timeline(prob._h, 1), prob._h)
            us2 = run_loop(prob, kern, bt, iters)
            print(f"   with stamps: {us2:.1f} us/eval")
            timeline(prob)
            prob.close()
    # T(n) sweep
    for pdl in ("drive", "1"):
        os.environ["LRB_DRIVE"] = "1" if pdl == "drive" else "0"
        os.environ["LRB_PDL"] = "1"
        pts = []
        for n in (1_562_500, 3_125_000, 6_250_000, 12_500_000, 25_000_000, 50_000_000):
            prob = lr.Problem(); bt = prob.gen_synthetic(n, 64, mode="fp32")
            kern = lr.hmcKernel(prob.lpost, prob.glp, eps=5 * 2.2 / np.sqrt(n) / 20, l=20, dmm=1.0)
            us = min(run_loop(prob, kern, bt, 20) for _ in range(3))
            pts.append((prob.info()["bytes_per_eval"], us)); prob.close()
        b = np.array(pts)
        A = np.vstack([np.ones(len(b)), b[:, 0]]).T
        coef = np.linalg.lstsq(A, b[:, 1], rcond=None)[0]
        print(f"T(n) PDL={pdl}: " + " ".join(f"{x / 1e9:.2f}GB:{t:.1f}us" for x, t in pts) + f" | fit a={coef[0]:.1f} us, B={1 / coef[1] / 1e6:.2f} TB/s", flush=True)
