"""Row-sharded multi-GPU plumbing (one process per GPU; torch.distributed only
carries the bootstrap bytes -- the data path is NCCL or NVLink peer memory
inside the library).

Precedent in the reference: the Spark map/reduce of the log-likelihood over row
partitions, Scala/spark/src/main/scala/fit-spark.scala:54-58 (ll only; here the
gradient is reduced too).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N
from . import build as _build


def shard_rows(n: int, rank: int, world: int):
    """Contiguous row block [lo, hi) of rank `rank`: the first n % world ranks get one extra row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _bcast_bytes(buf: bytes, n: int, src: int = 0) -> bytes:
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(n, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(buf), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def _allgather_bytes(buf: bytes) -> bytes:
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(buf), dtype=torch.uint8).to(dev)
    outs = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, mine)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in outs)


def init_comm(problem, kind: str = "p2p"):
    """Join `problem` (already created on this rank's GPU) to the row-sharded group
    of the initialised torch.distributed process group. kind: "nccl", "p2p", or "auto"
    (the fused peer-memory allreduce if every rank can map its peers, else NCCL)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = problem._lib
    if world == 1:
        return problem
    if kind == "auto":
        # Every collective stays OUTSIDE the try blocks: a rank whose export or connect fails must
        # still take part in the all_gather / all_reduce, or the others would wait for it forever.
        ok = 1
        mine = C.create_string_buffer(64)
        try:
            N.check(lib.lrb_comm_p2p_export(problem._h, mine), problem._h)
        except Exception:
            ok = 0                                   # gathers a dummy (zero) handle
        allh = _allgather_bytes(mine.raw)
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)     # can everybody export?
        if int(t.item()) == 1:
            try:
                N.check(lib.lrb_comm_p2p_connect(problem._h, rank, world, C.create_string_buffer(allh, 64 * world)), problem._h)
            except Exception:
                ok = 0
            t = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)  # could everybody map its peers?
        if int(t.item()) == 1:
            dist.barrier()                           # nobody writes a mailbox before every rank has cleared its own
            problem.world, problem.rank, problem.comm_kind = world, rank, "p2p"
            return problem
        kind = "nccl"     # some rank could not export / map peer memory: everyone falls back together
    if kind == "nccl":
        libnccl = _build.nccl_library().encode()
        uid = C.create_string_buffer(128)
        if rank == 0:
            N.check(lib.lrb_nccl_unique_id(uid, libnccl))
        raw = _bcast_bytes(uid.raw, 128)
        N.check(lib.lrb_comm_init_nccl(problem._h, rank, world, C.create_string_buffer(raw, 128), libnccl), problem._h)
    elif kind == "p2p":
        mine = C.create_string_buffer(64)
        N.check(lib.lrb_comm_p2p_export(problem._h, mine), problem._h)
        allh = _allgather_bytes(mine.raw)
        N.check(lib.lrb_comm_p2p_connect(problem._h, rank, world, C.create_string_buffer(allh, 64 * world)), problem._h)
        dist.barrier()
    else:
        raise ValueError("kind must be 'nccl' or 'p2p'")
    problem.world, problem.rank, problem.comm_kind = world, rank, kind
    return problem


def allreduce_partials_host(ll_local: float, gll_local: np.ndarray):
    """Host-side statement of the exchange step (used by the gloo tests of the
    sharding logic): sum [ll, gll] over ranks; the prior is added once afterwards."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.concatenate(([ll_local], np.asarray(gll_local, dtype=np.float64))))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    a = t.numpy()
    return float(a[0]), a[1:].copy()
