"""One-process-per-GPU launch plumbing (torchrun / torch.distributed): which rows or individuals a rank owns, how the
NCCL unique id of the pnol communicator reaches every rank, and max-over-ranks timing. Pure host logic -- the collectives
of the data path itself (packed J^T J / J^T r all-reduce, fitness all-gather) are issued by libpnol_b200.so on the
context's stream (csrc/comm.cu); torch.distributed only carries the rendezvous, barriers and the timing reduction.

Replaces what MPI_Init / MPI_Comm_rank / MPI_Comm_size give the reference (e.g. Source/LevenbergMarquardtMPI.cpp:15-17)."""
import os

import numpy as np


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def row_shard(m, world, rank):
    """Contiguous block [lo, hi) of m rows owned by `rank`: the first m % world ranks get one extra row, so any m works
    (the reference splits round-robin by COLUMN, Source/PNOL_Objective.cpp:235-246; rows are what shards naturally)."""
    m, world, rank = int(m), int(world), int(rank)
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank %d of %d" % (rank, world))
    base, extra = divmod(m, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def column_shard(n, world, rank):
    """FD-gradient coordinates [lo, hi) of `rank`, as pnol_fd_gradient deals them out (csrc/capi.cu, fd_gradient_common): equal
    blocks of ceil(n / world) coordinates, the last ones cut at n (possibly empty), so that one all-gather of fixed-size blocks in
    rank order rebuilds the gradient. The reference deals the coordinates round-robin and sums (Source/PNOL_Objective.cpp:110-148)."""
    n, world, rank = int(n), int(world), int(rank)
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank %d of %d" % (rank, world))
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def init_process_group(backend=None):
    """torch.distributed rendezvous on 127.0.0.1 (env:// as torchrun sets it). Returns (rank, local_rank, world)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def broadcast_bytes(payload, nbytes, root=0):
    """Every rank returns root's `payload` (bytes of length nbytes); other ranks may pass None."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(payload)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    if dist.get_rank() == root:
        buf = torch.tensor(list(bytes(payload)), dtype=torch.uint8, device=dev)
        assert buf.numel() == nbytes
    else:
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src=root)
    return bytes(buf.cpu().numpy().tobytes())


def attach_communicator(ctx):
    """Give a capi.Context its NCCL communicator: rank 0 draws the unique id, everyone joins (pnol_comm_init)."""
    import torch.distributed as dist
    from . import capi
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 1
    uid = ctx.comm_unique_id() if dist.get_rank() == 0 else None
    uid = broadcast_bytes(uid, capi.COMM_ID_BYTES, root=0)
    ctx.comm_init(uid, dist.get_world_size(), dist.get_rank())
    return dist.get_world_size()


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value):
    """max of a python float over all ranks (timing: a multi-GPU step takes as long as its slowest rank)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_over_ranks(arr):
    """list of every rank's copy of a float64 array (same shape on all ranks); [arr] without a process group"""
    import torch
    import torch.distributed as dist
    a = np.ascontiguousarray(arr, dtype=np.float64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [a.copy()]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(a.copy()).to(dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().numpy() for o in out]


def sum_over_ranks(arr):
    """element-wise float64 sum over ranks in torch.distributed (CPU tests of the sharded reductions)."""
    import torch
    import torch.distributed as dist
    a = np.ascontiguousarray(arr, dtype=np.float64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return a.copy()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(a.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
