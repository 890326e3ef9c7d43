"""Worker for the multi-GPU parity test (launched by torchrun, one rank per GPU):
row-sharded evaluation and sampler runs must equal the single-GPU result, on every rank,
bit-identically across ranks, for both communicators."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import logreg_b200 as lr  # noqa: E402
from logreg_b200 import dist as lrd  # noqa: E402


def main():
    kind = sys.argv[1]
    rank, world, local = lrd.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, p = 1_000_003, 64
    for mode, tol in (("fp32", 1e-6), ("fp64", 1e-11)):
        lo, hi = lrd.shard_rows(n, rank, world)
        prob = lr.Problem(local)
        bt = prob.gen_synthetic(hi - lo, p, mode=mode, seed=42, row_offset=lo)
        lrd.init_comm(prob, kind)
        full = lr.Problem(local)
        full.gen_synthetic(n, p, mode=mode, seed=42, beta_true=bt)
        rs = np.random.RandomState(3)
        for b in (bt, np.zeros(p), bt + 0.01 * rs.randn(p)):
            lp, l, g = prob.eval(b)
            lp1, l1, g1 = full.eval(b)
            assert abs(lp - lp1) <= tol * abs(lp1), (kind, mode, lp, lp1)
            assert np.max(np.abs(g - g1)) <= tol * max(1.0, np.max(np.abs(g1))) * 1e3, (kind, mode)
            # identical bits on all ranks
            t = torch.tensor(np.concatenate(([lp, l], g)), device="cuda")
            ts = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(ts, t)
            for o in ts:
                assert torch.equal(o, ts[0]), (kind, mode, "ranks disagree")
        # sampler loops: sharded chain == single-GPU chain (same Philox stream)
        sd = 2.2 / np.sqrt(n)
        for kern_s, kern_f in (
            (lr.hmcKernel(prob.lpost, prob.glp, eps=1.5 * sd / 8, l=8, dmm=1.0),
             lr.hmcKernel(full.lpost, full.glp, eps=1.5 * sd / 8, l=8, dmm=1.0)),
            (lr.malaKernel(prob.lpost, prob.glp, dt=(0.5 * sd) ** 2, pre=1.0),
             lr.malaKernel(full.lpost, full.glp, dt=(0.5 * sd) ** 2, pre=1.0)),
            (lr.mhKernel(prob.lpost, lr.RandomWalk(0.2 * sd * np.ones(p))),
             lr.mhKernel(full.lpost, lr.RandomWalk(0.2 * sd * np.ones(p)))),
        ):
            ms, accs = prob.run(kern_s, bt, 2, 12, seed=99)
            mf, accf = full.run(kern_f, bt, 2, 12, seed=99)
            assert accs == accf, (kind, mode, accs, accf)
            assert np.max(np.abs(ms - mf)) <= (1e-9 if mode == "fp64" else 1e-6), (kind, mode, np.max(np.abs(ms - mf)))
            t = torch.tensor(ms, device="cuda")
            ts = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(ts, t)
            for o in ts:
                assert torch.equal(o, ts[0]), (kind, mode, "chains disagree across ranks")
        # several chains on a row-sharded handle (NCCL: lock-step many-chain kernel; fused P2P:
        # chain by chain) == the same chains on one GPU
        k_s = lr.malaKernel(prob.lpost, prob.glp, dt=(0.5 * sd) ** 2, pre=1.0)
        k_f = lr.malaKernel(full.lpost, full.glp, dt=(0.5 * sd) ** 2, pre=1.0)
        inits = bt + 0.3 * sd * np.random.RandomState(9).randn(3, p)
        ms, accs = prob.run_chains(k_s, inits, 1, 8, seed=4)
        mf, accf = full.run_chains(k_f, inits, 1, 8, seed=4)
        assert np.array_equal(accs, accf), (kind, mode, accs, accf)
        assert np.max(np.abs(ms - mf)) <= (1e-9 if mode == "fp64" else 1e-6), (kind, mode)
        prob.close()
        full.close()
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK kind={kind} world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
