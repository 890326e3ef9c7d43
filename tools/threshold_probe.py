"""Chain-count threshold between the SIMT many-chain kernel and the tcgen05 kernel
(n=1e6, p=64, fp32): time per all-chain evaluation on each path."""
import os, sys, time, subprocess, json
import numpy as np
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import logreg_b200 as lr, torch
    prob = lr.Problem(); bt = prob.gen_synthetic(1_000_000, 64, mode="fp32")
    out = {}
    for C in (2, 4, 8, 16, 32, 64, 128, 256):
        B = np.tile(bt, (C, 1)) + 0.01 * np.random.RandomState(C).randn(C, 64)
        prob.eval_many(B); prob.eval_many(B)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); prob.eval_many(B); ts.append(time.perf_counter() - t0)
        out[C] = min(ts) * 1e3
    print(json.dumps(out))
else:
    res = {}
    for name, thr in (("simt", "1000000"), ("tc", "2")):
        env = dict(os.environ, LRB_TC_MIN_CHAINS=thr)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        res[name] = json.loads(r.stdout.strip().splitlines()[-1])
    print("C    simt_ms   tc_ms   (host-timed all-chain evaluation, n=1e6 p=64 fp32)")
    for C in res["simt"]:
        print(f"{C:>4s} {res['simt'][C]:8.3f} {res['tc'][C]:8.3f}")
