import sys, numpy as np
sys.path.insert(0, ".")
import logreg_b200 as lr
prob = lr.Problem(); bt = prob.gen_synthetic(1000, 64, mode="fp32")
for name, kern in (("hmc", lr.hmcKernel(prob.lpost, prob.glp, eps=1e-4, l=20, dmm=1.0)),
                   ("mala", lr.malaKernel(prob.lpost, prob.glp, dt=1e-6, pre=1.0))):
    prob.run(kern, bt, 1, 3, seed=1)
