"""GPU, 2 ranks (skipped on a one-GPU box): the NCCL paths of the C-ABI -- row-sharded LM (packed J^T J | J^T r all-reduce, chi^2
all-reduce), column-split FD gradient (all-gather) and the raw collectives -- against the single-rank results / the golden
outputs of the verbatim reference. One process per GPU, rendezvous on 127.0.0.1."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from parallelnonlinearoptimizationlibrary_b200 import capi, hostapi, launch
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
    launch.init_process_group("nccl")
    ctx = capi.Context(rank)
    assert launch.attach_communicator(ctx) == world and ctx.comm_size() == world and ctx.comm_rank() == rank
    # raw all-reduce
    buf = np.arange(5, dtype=np.float64) + rank
    ctx.allreduce_sum(buf, 5)
    assert np.array_equal(buf, 2 * np.arange(5) + 1.0)
    # column-split FD gradient == the reference's gradient, bit for bit (each coordinate is one independent evaluation)
    f = ctx.functor(capi.F_ROSENBROCK)
    g, f0 = ctx.fd_gradient(f, G["fdgrad_rosenbrock_40/x"], G["fdgrad_rosenbrock_40/dx"])
    assert np.array_equal(g, G["fdgrad_rosenbrock_40/g"]) and f0 == G["fdgrad_rosenbrock_40/f"][0]
    # row-sharded LM through the plugin class: every rank holds its block of rows
    hostapi.attach(ctx)
    c = "lm_lorentz_K8"
    t, y = G[c + "/t"], G[c + "/y"]
    lo, hi = launch.row_shard(t.size, world, rank)
    r = hostapi.lm_lorentz(t[lo:hi], y[lo:hi], float(G[c + "/w"]), G[c + "/x0"], 0.001, 10.0, 1e-7, int(G[c + "/iters"]), 0.0)
    np.save(os.path.join(out_dir, "X%d.npy" % rank), r["X"])
    np.save(os.path.join(out_dir, "F0_%d.npy" % rank), r["F0"])
    # GA with the fitness sweep sharded over the ranks (all-gather of F): bit-exact like the single-GPU run
    c = "ga_rastrigin"
    hostapi.set_stream(seed=int(G[c + "/seed"]), scale=float(G[c + "/scale"]))
    r = hostapi.ga("rastrigin", G[c + "/x0"], G[c + "/lb"], G[c + "/ub"], int(G[c + "/npop"]), int(G[c + "/gens"]))
    assert np.array_equal(r["X"], G[c + "/X"]) and r["fOpt"] == G[c + "/fOpt"][0] and r["stream_pos"] == int(G[c + "/stream_pos"][0])
    hostapi.detach()
    launch.barrier()
    ctx.close()
    import torch.distributed as dist
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_rank_lm_and_gradient(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
    X0, X1 = np.load(tmp_path / "X0.npy"), np.load(tmp_path / "X1.npy")
    assert np.array_equal(X0, X1)
    want = G["lm_lorentz_K8/X"]
    assert np.linalg.norm(X0 - want) <= 1e-9 * np.linalg.norm(want)
    F0 = np.concatenate([np.load(tmp_path / "F0_0.npy"), np.load(tmp_path / "F0_1.npy")])
    assert np.array_equal(F0, G["lm_lorentz_K8/F0"])
