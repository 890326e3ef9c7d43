#!/bin/bash
# compute-sanitizer pass over small instances of every kernel (memcheck; racecheck on the SIMT kernels)
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import logreg_b200 as lr
g = dict(np.load("tests/golden/pima.npz"))
prob = lr.Problem().bind_data(np.asfortranarray(g["X"]), g["y"], g["pscale"])
print("pima lpost", prob.lpost(g["B"][1]))
for kern in (lr.mhKernel(prob.lpost, lr.RandomWalk(0.02 * g["pre_rw"])), lr.ulKernel(prob.glp, dt=1e-6, pre=g["pre"]),
             lr.malaKernel(prob.lpost, prob.glp, dt=1e-5, pre=g["pre"]), lr.hmcKernel(prob.lpost, prob.glp, eps=1e-3, l=5, dmm=1 / g["pre"])):
    m, a = prob.run(kern, g["chain_init"], 2, 5, seed=1)
m, a = prob.run_chains(kern, np.tile(g["chain_init"], (5, 1)), 1, 3, seed=2)
for mode in ("fp32", "fp64"):
    for p in (13, 64, 200):
        q = lr.Problem(); bt = q.gen_synthetic(3001, p, mode=mode)
        q.eval(bt); q.eval_many(np.tile(bt, (3, 1))); q.copy_rows(5, 17); q.close()
q = lr.Problem(); bt = q.gen_synthetic(1000, 64, mode="fp32")
lp, l, gg = q.eval_many(np.tile(bt, (140, 1)))          # tensor-core path, 2 chain groups, ragged tail tile
k = lr.malaKernel(q.lpost, q.glp, dt=1e-4, pre=1.0)
q.run_chains(k, np.tile(bt, (20, 1)), 1, 3, seed=3)
print("ok", lp[0])
PY
compute-sanitizer --tool memcheck --error-exitcode 1 python /tmp/san_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 1 python /tmp/san_small.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/sanitize_racecheck.log
