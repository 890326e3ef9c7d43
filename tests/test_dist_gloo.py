"""World-size-2 test of the row-sharded decomposition on CPU (gloo): each rank
evaluates the partial sums [ll, X'(y-p)] of its row block (with the oracle as the
per-shard arithmetic), the (p+1)-vector is all-reduced, the prior is added once --
the exchange step the GPU path performs with NCCL / NVLink peer memory
(precedent: Scala/spark/src/main/scala/fit-spark.scala:54-58)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        from logreg_b200 import dist as lrd
        from oracle import logreg_oracle as O
        g = dict(np.load(os.path.join(root, "tests", "golden", "synth2000x32.npz")))
        X = g["X32"].astype(np.float64)
        t = O.Target(X, g["y"], g["pscale"])
        lo, hi = lrd.shard_rows(t.n, rank, world)
        out = []
        for b in g["B"]:
            ll_loc, gll_loc = t.ll_gll_rows(b, lo, hi)
            ll, gll = lrd.allreduce_partials_host(ll_loc, gll_loc)
            lpost = ll + t.lprior(b)                       # prior added once, after the exchange
            glp = gll - b / (t.pscale * t.pscale)
            out.append((lpost, glp))
        # bootstrap byte plumbing used for the NCCL id / IPC handles
        raw = lrd._bcast_bytes(bytes(range(128)) if rank == 0 else b"", 128)
        allh = lrd._allgather_bytes(bytes([rank]) * 64)
        q.put((rank, out, raw == bytes(range(128)), allh == b"".join(bytes([r]) * 64 for r in range(world))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_row_sharded_sum_equals_full_evaluation():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "synth2000x32.npz")))
    results.sort(key=lambda r: r[0])
    for rank, out, ok_b, ok_g in results:
        assert ok_b and ok_g
        for i, (lpost, glp) in enumerate(out):
            assert lpost == pytest.approx(g["lpost"][i], rel=1e-12)
            np.testing.assert_allclose(glp, g["glp"][i], rtol=1e-9, atol=1e-9)
    # every rank holds bit-identical results (what keeps replicated sampler state in lock-step)
    for (l0, g0), (l1, g1) in zip(results[0][1], results[1][1]):
        assert l0 == l1
        np.testing.assert_array_equal(g0, g1)


class _FakeLib:
    """Stands in for liblogreg_b200 in the communicator bootstrap: records the calls and fails
    where told, so the collective sequence of dist.init_comm can be checked on CPU."""

    def __init__(self, fail_export=False, fail_connect=False):
        self.fail_export, self.fail_connect, self.calls = fail_export, fail_connect, []

    def lrb_comm_p2p_export(self, h, buf):
        self.calls.append("export")
        return 3 if self.fail_export else 0

    def lrb_comm_p2p_connect(self, h, rank, world, handles):
        self.calls.append("connect")
        return 3 if self.fail_connect else 0

    def lrb_nccl_unique_id(self, uid, path):
        self.calls.append("uid")
        return 0

    def lrb_comm_init_nccl(self, h, rank, world, uid, path):
        self.calls.append("nccl")
        return 0


class _FakeProblem:
    def __init__(self, lib):
        self._lib, self._h = lib, None


def _comm_worker(rank, world, port, q, scenario):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        from logreg_b200 import dist as lrd
        lib = _FakeLib(fail_export=(scenario == "export" and rank == 1), fail_connect=(scenario == "connect" and rank == 0))
        prob = _FakeProblem(lib)
        lrd.init_comm(prob, "auto")
        q.put((rank, prob.comm_kind, lib.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("scenario,kind", [("ok", "p2p"), ("export", "nccl"), ("connect", "nccl")])
def test_comm_auto_falls_back_jointly_without_a_collective_mismatch(scenario, kind):
    """if ONE rank cannot export / map peer memory, every rank must still run the same
    collectives and all fall back to NCCL together (no rank left waiting in an all_gather)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_comm_worker, args=(r, world, port, q, scenario)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, got, calls in results:
        assert got == kind, (scenario, rank, got, calls)
        assert calls[0] == "export"
        if kind == "p2p":
            assert calls == ["export", "connect"]
        else:
            assert calls[-1] == "nccl"
            if scenario == "export":
                assert "connect" not in calls          # nobody maps peers when an export failed
