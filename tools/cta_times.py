import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["LRB_CTA_TIMES"] = "1"
import logreg_b200 as lr
for n in (100_000_000, 12_500_000, 1_000_000):
    prob = lr.Problem(); bt = prob.gen_synthetic(n, 64, mode="fp32")
    print("n =", n, flush=True)
    prob.eval(bt)
    prob.close()
