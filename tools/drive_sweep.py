"""Drive mode: sweep the static share of the batch schedule (development aid)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from tail_diag import run_loop, timeline
import logreg_b200 as lr
from logreg_b200 import _native as N

cfgs = [(12_500_000, 64, 20, "hmc"), (50_000_000, 64, 20, "hmc"), (1_000_000, 32, 1, "mala")]
modes = [("static-kernel", "0", "4")] + [(f"drive static={e}/8", "1", str(e)) for e in (8, 6, 4, 2, 0)]
for n, p, L, samp in cfgs:
    for name, drive, eighths in modes:
        os.environ["LRB_DRIVE"] = drive
        os.environ["LRB_DRIVE_STATIC"] = eighths
        prob = lr.Problem(); bt = prob.gen_synthetic(n, p, mode="fp32")
        sd = 2.2 / np.sqrt(n)
        kern = (lr.hmcKernel(prob.lpost, prob.glp, eps=5 * sd / L, l=L, dmm=1.0) if samp == "hmc"
                else lr.malaKernel(prob.lpost, prob.glp, dt=(0.6 * sd) ** 2, pre=1.0))
        iters = 20 if samp == "hmc" else 2000
        us = min(run_loop(prob, kern, bt, iters) for _ in range(3))
        byt = prob.info()["bytes_per_eval"]
        print(f"n={n} p={p} {samp} {name}: {us:.1f} us/eval ({byt / us / 1e3:.0f} GB/s)", flush=True)
        if "--timeline" in sys.argv and eighths in ("4", "8"):
            N.check(prob._lib.lrb_debug_timeline(prob._h, 1), prob._h)
            run_loop(prob, kern, bt, iters)
            timeline(prob)
        prob.close()
