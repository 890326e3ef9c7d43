#!/usr/bin/env python3
"""bench.py -- the driver's measurement contract for logreg_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one `metric` is quoted on; 25.6 GB fits one
B200): HMC with a diagonal mass matrix, L=20 leap-frog steps, synthetic n=1e8, p=64,
float32 X, row-sharded over the N GPUs (strong scaling: n is the TOTAL row count).
A "step" is one HMC iteration = L fused lpost+glp evaluations over all n rows.

metric  "lpost+grad evals/s" = L * (HMC iterations / s): leap-frog gradient
        evaluations completed per second.  The same definition is used for the
        reference arm (which spends 2L+4 passes over X per iteration to deliver the
        same L leap-frog evaluations, fit-np-hmc.py:56-87), so the ratio of the two
        arms is the ratio of HMC iterations/s.
value   device-resident: data and chain state in HBM, CUDA-event timed, max over ranks.
e2e     the same K iterations through the public Python API, one
        mcmc(x, hmcKernel(lpost, glp, ...), thin=1, iters=1) call per step with HOST
        state in and HOST samples out (pinned staging inside the library).  X itself is
        bound once (bind_data / gen_synthetic), exactly as the reference keeps X in a
        module global; its one-off H2D cost is not part of a sampler step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: sampler, n, p, mode, L, step(n) -> tuned so that acceptance is healthy
    "c3": dict(sampler="hmc", n=100_000_000, p=64, mode="fp32", L=20, desc="HMC L=20 diag mass, n=1e8 p=64 fp32 X"),
    "c2": dict(sampler="mala", n=1_000_000, p=32, mode="fp32", L=1, desc="MALA diag precond, n=1e6 p=32 fp32 X"),
    "c5": dict(sampler="ul", n=400_000_000, p=128, mode="fp64", L=1, desc="UL, n=4e8 p=128 fp64 X (needs >= 3 GPUs)"),
    # config 4: many chains, X replicated, chains sharded over the GPUs, no collective
    "c4": dict(sampler="mala", n=1_000_000, p=64, mode="fp32", L=1, chains=4096,
               desc="4096 MALA chains, n=1e6 p=64 fp32 X (tcgen05 3xTF32 many-chain kernel)"),
}
SAMPLE_ROWS = 1_000_000     # rows of the same workload the CPU arm is timed on


def step_size(w, n):
    # posterior sd ~ 2.2/sqrt(n) per coefficient (unit-variance covariates): scale steps with it
    sd = 2.2 / np.sqrt(n)
    if w["sampler"] == "hmc":
        return 5.0 * sd / w["L"]          # 0.25 sd per leap-frog step, trajectory ~ 5 sd
    if w["sampler"] == "mala":
        return (0.6 * sd) ** 2            # dt = sd_prop^2
    return (0.3 * sd) ** 2


# ---------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------- CPU arm (oracle port of the reference)
def synth_host(n, p, seed=42):
    """Same distribution as the device generator: X[:,0]=1, X[:,1:]~N(0,1), y~Bernoulli(expit(X bt))."""
    rs = np.random.RandomState(seed)
    bt = np.random.RandomState(41).randn(p) / np.sqrt(p)
    X = np.empty((n, p), order="F")        # the reference's X is column-major float64
    X[:, 0] = 1.0
    for j in range(1, p):
        X[:, j] = rs.randn(n)
    y = (rs.rand(n) < 1 / (1 + np.exp(-X.dot(bt)))).astype(np.float32)
    ps = np.ones(p); ps[0] = 10.0
    return X, y, ps, bt


def cpu_kernel(w, n_full, ns):
    from oracle import logreg_oracle as O
    X, y, ps, bt = synth_host(ns, w["p"])
    tgt = O.Target(X, y, ps)
    h = step_size(w, ns)      # tuned for the sample it runs on (same acceptance regime)
    if w["sampler"] == "hmc":
        k = O.hmc_kernel(tgt.lpost, tgt.glp, eps=h, l=w["L"], dmm=1.0)
        step = lambda st: (k(st[0]), None)
    elif w["sampler"] == "mala":
        k = O.mala_kernel(tgt.lpost, tgt.glp, w["p"], dt=h, pre=1.0)
        step = lambda st: k(st[0], st[1])
    else:
        k = O.ul_kernel(tgt.glp, w["p"], dt=h, pre=1.0)
        step = lambda st: (k(st[0]), None)
    return step, (bt.copy(), -np.inf)


def time_cpu(w, n_full, steps, warmup, budget_s=None):
    # use every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently make the reference arm single-threaded
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    ns = min(SAMPLE_ROWS, n_full)
    step, st = cpu_kernel(w, n_full, ns)
    np.random.seed(7)
    for _ in range(warmup):
        st = step(st)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        st = step(st)
        done += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    scale = n_full / ns                     # the path is a stream over rows: time is linear in n
    ms_step_full = dt * 1e3 * scale
    try:
        from threadpoolctl import threadpool_info
        blas = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        blas = os.cpu_count()
    return {"ms_per_step": ms_step_full, "iters_per_s": 1e3 / ms_step_full, "steps": done,
            "cores": int(blas), "host_cpus": os.cpu_count(),
            "sample": f"{w['desc'].split(',')[0]} reference kernel (oracle port, NumPy/OpenBLAS, X float64 "
                      f"column-major as the reference builds it) on {ns} rows x {done} iterations, "
                      f"time scaled x{scale:g} to n={n_full} (linear in n)"}


def bench_chains(args, w, prob, kern, bt, rank, world, local, barrier, config, metric):
    """Config 4: C chains in lock-step per GPU (tensor-core many-chain kernel), chains sharded
    over the ranks with no collective. A step = one MALA iteration of every chain."""
    import torch
    import torch.distributed as dist
    import logreg_b200 as lr
    K, W = max(args.steps, 40), max(args.warmup, 10)   # ~7 ms per step at 4096 chains: keep the timed region >= 0.25 s
    n, p, C = w["n"], w["p"], w["chains"]
    c_lo, c_hi = (rank * C) // world, ((rank + 1) * C) // world
    Cl = c_hi - c_lo
    sd = 2.2 / np.sqrt(n)
    inits = bt + 0.5 * sd * np.random.RandomState(100 + rank).randn(Cl, p)
    prob.run_chains(kern, inits, 1, W, seed=7 + rank)
    inf0 = prob.info()
    barrier()
    t0 = time.perf_counter()
    mats, acc = prob.run_chains(kern, inits, 1, K, seed=7 + rank)     # host inits in, host samples out
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    inf1 = prob.info()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    value = C * K / dt
    flops = 4.0 * n * p * Cl * (K + 1)             # algorithmic (1-pass) flops on this rank, incl. the init evaluation
    peak_tf32 = float(peaks.get("bf16_tflops", 1590.0)) / 2
    line = {"metric": metric, "value": value, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak" if world > 1 else "strong",
            "vs_baseline": None, "dtype": "tf32x3", "data": "synthetic",
            "config": dict(config, chains=C, chains_per_gpu=Cl, parallelism=f"chain-sharded x{world}", l2="X (256 MB) exceeds L2 (126 MB)"),
            "chain_iters_per_s": value, "accept_rate": float(acc.mean() / K),
            "roofline": {"bound": "tensor", "achieved": flops / dt / 1e12, "peak": peak_tf32, "unit": "TFLOP/s",
                         "frac": flops / dt / 1e12 / peak_tf32, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 dense runs at half the bf16 rate)",
                         "note": "algorithmic 4*n*p*C flops per all-chain evaluation; the 3xTF32 kernel executes 3x that on the "
                                 "tensor pipe and is co-limited by 3 MUFU ops per (row, chain) element"},
            "cpu_baseline": None,
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": Cl * p * 8 / K, "d2h_bytes_per_step": Cl * p * 8,
                    "note": "timed through Problem.run_chains (host inits in, host samples out)"},
            "gpu_launches": int(inf1["kernel_launches"] - inf0["kernel_launches"]), "clocks": None}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--comm", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--n", type=int, default=0, help="override total rows (development only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.n:
        w["n"] = args.n
    K, W, L = args.steps, max(args.warmup, 0), w["L"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "lpost+grad evals/s"
    config = {"workload": f"{w['desc']}, {w['sampler'].upper()} on-device loop", "n": w["n"], "p": w["p"],
              "L": L, "x_dtype": w["mode"], "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
              "l2": "inputs exceed L2 (X is %.1f GB per GPU vs 126 MB L2)" % (w["n"] / world * w["p"] * (4 if w["mode"] == "fp32" else 8) / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = time_cpu(w, w["n"], K, W)
        val = L * r["iters_per_s"]
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "evals/s", "n_gpus": args.gpus,
                "steps": r["steps"], "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "iters_per_s": r["iters_per_s"],
                "cpu_baseline": {"value": val, "unit": "evals/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import logreg_b200 as lr
    from logreg_b200 import dist as lrd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; logreg_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, p = w["n"], w["p"]
    lo, hi = (0, n) if "chains" in w else lrd.shard_rows(n, rank, world)
    prob = lr.Problem(local)
    prob.n_global = n
    bt = prob.gen_synthetic(hi - lo, p, mode=w["mode"], seed=42, row_offset=lo)
    if world > 1 and "chains" not in w:
        lrd.init_comm(prob, args.comm)
    h = step_size(w, n)
    if w["sampler"] == "hmc":
        kern = lr.hmcKernel(prob.lpost, prob.glp, eps=h, l=L, dmm=1.0)
    elif w["sampler"] == "mala":
        kern = lr.malaKernel(prob.lpost, prob.glp, dt=h, pre=1.0)
    else:
        kern = lr.ulKernel(prob.glp, dt=h, pre=1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if "chains" in w:
        return bench_chains(args, w, prob, kern, bt, rank, world, local, barrier, config, metric)

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib, hd = prob._lib, prob._h
    import ctypes as C
    from logreg_b200 import _native as N
    stream = torch.cuda.Stream()
    prob.set_stream(stream.cuda_stream)
    sp = prob._params(kern, seed=2026, rng=N.RNG_PHILOX, init_lpost=-np.inf)

    # ---- device-resident timing: warm-up run (W iterations), then K iterations per launch
    def begin(iters):
        N.check(lib.lrb_run_begin(hd, C.byref(sp), N.as_dp(np.ascontiguousarray(bt)), 1, iters, None, None), hd)
    with torch.cuda.stream(stream):
        begin(max(W, 1))
        N.check(lib.lrb_run_launch(hd), hd)       # includes the one-off evaluation at init
        acc = C.c_int64()
        N.check(lib.lrb_run_finish(hd, None, C.byref(acc)), hd)
        # continue the same chain for K timed iterations
        sp2 = prob._params(kern, seed=2026, rng=N.RNG_PHILOX, init_lpost=-np.inf)
        N.check(lib.lrb_run_begin(hd, C.byref(sp2), None, 1, K, None, None), hd)
        inf0 = prob.info()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
            time.sleep(0.3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.time()
        e0.record(stream)
        N.check(lib.lrb_run_launch(hd), hd)
        e1.record(stream)
        stream.synchronize()
        barrier()
        tw1 = time.time()
        ms_total = maxr(e0.elapsed_time(e1))
        clk = clocks.stop(tw0, tw1) if rank == 0 else None
        inf1 = prob.info()
        out = np.empty((K, p))
        acc1 = C.c_int64()
        N.check(lib.lrb_run_finish(hd, N.as_dp(out), C.byref(acc1)), hd)
    accept_rate = (acc1.value - acc.value) / K
    launches = inf1["kernel_launches"] - inf0["kernel_launches"]
    evals = inf1["eval_launches"] - inf0["eval_launches"]
    assert evals == K * L, (evals, K, L)
    ms_step = ms_total / K
    value = K * L / (ms_total / 1e3)
    prob.set_stream(None)

    # ---- e2e through the public API: host state in, host sample out, every step
    x = out[-1].copy()
    for _ in range(min(W, 2)):
        x = lr.mcmc(x, kern, thin=1, iters=1, verb=False, seed=11)[0]
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        x = lr.mcmc(x, kern, thin=1, iters=1, verb=False, seed=100 + i)[0]
    torch.cuda.synchronize()
    e2e_s = maxr(time.perf_counter() - t0)
    e2e_val = K * L / e2e_s
    # one call for all K iterations (how a script calls it)
    barrier()
    t0 = time.perf_counter()
    lr.mcmc(x, kern, thin=1, iters=K, verb=False, seed=5)
    e2e_single = K * L / maxr(time.perf_counter() - t0)

    # ---- roofline of the fused kernel (it is the only kernel in the timed region)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_eval_rank = inf1["bytes_per_eval"]            # this rank's shard
    kern_ms = ms_total / evals                           # includes the in-kernel allreduce + sampler update
    achieved = bytes_eval_rank / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        if world == 1 and not args.n:   # the committed ncu capture is of the single-GPU launch
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
    except Exception:
        pass
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
            "frac_of_nominal_8TBs": achieved / 8000.0, "kernel": "lrb::eval_kernel<%s,%d,grad>" % ("float" if w["mode"] == "fp32" else "double", inf1["p_pad"]),
            "algorithmic_bytes_per_launch": bytes_eval_rank, "avg_launch_ms": kern_ms}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    # one-off ingest (bind_data) of a host-resident X in the reference's own layout (float64,
    # column-major), on a 2e6-row sample: reported beside e2e, not part of a sampler step
    ingest = None
    if world == 1:
        try:
            ns_i = 2_000_000
            Xh = np.asfortranarray(np.random.RandomState(1).randn(ns_i, p))
            yh = (np.random.RandomState(2).rand(ns_i) < 0.5).astype(np.float32)
            pi = lr.Problem(local)
            t0 = time.perf_counter()
            pi.bind_data(Xh, yh, np.ones(p), mode=w["mode"])
            dt_i = time.perf_counter() - t0
            pi.close()
            ingest = {"rows": ns_i, "host_layout": "float64 column-major (reference)", "seconds": dt_i,
                      "host_GB_per_s": Xh.nbytes / dt_i / 1e9,
                      "extrapolated_seconds_full_n": dt_i * n / ns_i}
            del Xh, yh
        except Exception as e:  # never let the optional figure break the contract line
            ingest = {"error": str(e)}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = time_cpu(w, n, steps=12, warmup=1, budget_s=25.0)
        cpu = {"value": L * r["iters_per_s"], "unit": "evals/s", "cores": r["cores"], "kind": "port",
               "sample": r["sample"], "iters_per_s": r["iters_per_s"], "host_cpus": r["host_cpus"]}
    line = {"metric": metric, "value": value, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if w["mode"] == "fp32" else "f64", "data": "synthetic", "config": config,
            "iters_per_s": 1e3 / ms_step, "accept_rate": accept_rate, "step_size": h,
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": 2 * p * 8, "d2h_bytes_per_step": p * 8 + 8 + 4,
                    "single_call_value": e2e_single, "one_off_ingest": ingest,
                    "note": "K mcmc(x, kernel, thin=1, iters=1) calls, host state in/out each step (a call that starts "
                            "where the previous one stopped reuses the cached gradient: L passes per step); "
                            "single_call_value = one mcmc(iters=K) call"},
            "gpu_launches": int(launches), "comm": (getattr(prob, "comm_kind", args.comm) if world > 1 else None), "clocks": clk}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
