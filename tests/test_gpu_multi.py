"""GPU, 2 / 4 / 8 ranks (each case skipped when the box has fewer GPUs): the NCCL paths of the C-ABI against the single-rank
results, the oracle and the golden outputs of the verbatim reference. One process per GPU, rendezvous on 127.0.0.1.

  * raw collectives;
  * column-split FD gradient (pnol_fd_gradient, all-gather; Source/PNOL_Objective.cpp:88-159) at n = 40 (golden) and n = 4096
    (cfg3's size, oracle): bit-exact on every rank;
  * row-sharded Levenberg-Marquardt through LevMarqMPI::findMin (packed J^T J | J^T r all-reduce, chi^2 all-reduce;
    Source/LevenbergMarquardtMPI.cpp:55-141) at n = 16 and at n = 256 (K = 128): X within 1e-9 of the verbatim reference, X
    bit-identical on all ranks, F0 shards bit-exact;
  * GeneticAlgorithmMPI::findMinBnd with the generation sharded over the ranks: bit-exact like the single-GPU run, with an
    explicit stream AND with the default (clock-seeded, rank-0-broadcast) stream, where all ranks must still agree."""
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


def _worker(rank, world, port, out_dir, shard):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import oracle_lib as O
    from parallelnonlinearoptimizationlibrary_b200 import capi, hostapi, launch, problems
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
    GB = np.load(os.path.join(ROOT, "tests", "golden", "baseline_lm_golden.npz"))
    launch.init_process_group("nccl")
    ctx = capi.Context(rank)
    assert launch.attach_communicator(ctx) == world and ctx.comm_size() == world and ctx.comm_rank() == rank
    ctx.ga_set_sharding(shard)                     # 1: GA rows sharded over the ranks, 2: rows replicated + sweep sharded, 3: replicas
    # raw all-reduce
    buf = np.arange(5, dtype=np.float64) + rank
    ctx.allreduce_sum(buf, 5)
    assert np.array_equal(buf, world * np.arange(5) + world * (world - 1) / 2.0)
    # column-split FD gradient == the reference's gradient, bit for bit (each coordinate is one independent evaluation)
    f = ctx.functor(capi.F_ROSENBROCK)
    g, f0 = ctx.fd_gradient(f, G["fdgrad_rosenbrock_40/x"], G["fdgrad_rosenbrock_40/dx"])
    assert np.array_equal(g, G["fdgrad_rosenbrock_40/g"]) and f0 == G["fdgrad_rosenbrock_40/f"][0]
    # ... and at cfg3's size, n = 4096 (coordinates per rank: 2048 / 1024 / 512), against the oracle
    rng = np.random.default_rng(4096)
    x, dx = rng.uniform(-2, 2, 4096), np.full(4096, 1e-6)
    g, f0 = ctx.fd_gradient(f, x, dx)
    gw, f0w = O.fd_gradient(O.OFunctor(capi.F_ROSENBROCK), x, dx)
    assert np.array_equal(g, gw) and f0 == f0w
    # row-sharded LM through the plugin class: every rank holds its block of rows
    hostapi.attach(ctx)
    hostapi.set_jacobian_cache(False)
    # the SERIAL members / classes never touch the communicator (local mode): rank 0 may call them alone, as with the reference
    if rank == 0:
        g = hostapi.gradient("rosenbrock", G["fdgrad_rosenbrock_40/x"], G["fdgrad_rosenbrock_40/dx"], mpi=False)
        assert np.array_equal(g, G["fdgrad_rosenbrock_40/g"])
        rb = hostapi.bfgs("bfgs", "rosenbrock", np.full(10, 3.0), [1e-4, 0.9, 1e-6, 1.0, 1000, 1e-7, 1e-3, 5, 1e-5, 1e-5, 0])
        assert rb["f0"] == G["bfgs_cfg1_it5/f0"][0]
        assert ctx.comm_size() == world            # and the communicator is back afterwards
    c = "lm_lorentz_K8"
    t, y = G[c + "/t"], G[c + "/y"]
    lo, hi = launch.row_shard(t.size, world, rank)
    r = hostapi.lm_lorentz(t[lo:hi], y[lo:hi], float(G[c + "/w"]), G[c + "/x0"], 0.001, 10.0, 1e-7, int(G[c + "/iters"]), 0.0)
    np.save(os.path.join(out_dir, "X%d.npy" % rank), r["X"])
    np.save(os.path.join(out_dir, "F0_%d.npy" % rank), r["F0"])
    # n = 256 (K = 128, the headline parameter count), m = 4096 rows over the ranks
    pr = problems.lorentz_problem(int(GB["lm_K128/m"]), int(GB["lm_K128/K"]))
    lo, hi = launch.row_shard(pr["m"], world, rank)
    r = hostapi.lm_lorentz(pr["t"][lo:hi], pr["y"][lo:hi], pr["w"], pr["x0"], 0.001, 10.0, 1e-7, int(GB["lm_K128/maxiter"]), 0.0)
    np.save(os.path.join(out_dir, "XK128_%d.npy" % rank), r["X"])
    np.save(os.path.join(out_dir, "F0K128_%d.npy" % rank), r["F0"])
    np.save(os.path.join(out_dir, "repK128_%d.npy" % rank), np.array([r["iterations"], r["lam"], ctx.lm_exchange_mode()]))
    # GA with the generation sharded over the ranks: bit-exact like the single-GPU run
    c = "ga_rastrigin"
    hostapi.set_stream(seed=int(G[c + "/seed"]), scale=float(G[c + "/scale"]))
    r = hostapi.ga("rastrigin", G[c + "/x0"], G[c + "/lb"], G[c + "/ub"], int(G[c + "/npop"]), int(G[c + "/gens"]))
    assert np.array_equal(r["X"], G[c + "/X"]) and r["fOpt"] == G[c + "/fOpt"][0] and r["stream_pos"] == int(G[c + "/stream_pos"][0])
    # a larger population (several row blocks per rank) against the oracle, generation by generation
    n, npop, gens = 16, 6000, 3
    lb, ub = np.full(n, -5.12), np.full(n, 5.12)
    x0 = np.full(n, 1.5)
    stream = dict(seed=4242, scale=1.0 - 1.0 / npop)
    fr = ctx.functor(capi.F_RASTRIGIN)
    ga = ctx.ga_create(fr, n, lb, ub, npop, gens, stream)
    assert ga.peer_mode() == (3 if shard == 2 else 4 if shard == 3 else (2 if os.environ.get("PNOL_GA_NO_IPC") == "1" else ga.peer_mode()))
    assert ga.peer_mode() in (1, 2, 3, 4)
    np.save(os.path.join(out_dir, "peer_mode_%d.npy" % rank), np.array([ga.peer_mode()]))
    ga.init(x0)
    for gen in range(1, gens + 1):
        ga.generation()
        want = O.ga(O.OFunctor(capi.F_RASTRIGIN), x0, lb, ub, npop, gens, stream, stop_after=gen)
        X, F = ga.population()
        cross, mut, elite = ga.indices()
        assert ga.status().stream_pos == want["stream_pos"], "stream position, generation %d" % gen
        assert np.array_equal(F, want["F"]) and np.array_equal(X, want["xpop"]), "population, generation %d" % gen
        assert np.array_equal(cross, want["cross_idx"]) and np.array_equal(mut, want["mut_idx"]) and np.array_equal(elite, want["elite_idx"])
    ga.close()
    # no stream set: the default stream is seeded on rank 0 and broadcast, so the ranks must still agree bit for bit
    hostapi.clear_stream()
    r = hostapi.ga("rastrigin", np.full(6, 1.0), np.full(6, -5.12), np.full(6, 5.12), 500, 4)
    np.save(os.path.join(out_dir, "gaX_%d.npy" % rank), np.concatenate([r["X"], [r["fOpt"], float(r["stream_pos"])]]))
    hostapi.detach()
    launch.barrier()
    ctx.close()
    import torch.distributed as dist
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shard,no_ipc,lm_peer", [(2, 1, 0, 1), (2, 1, 1, 1), (2, 2, 0, 0), (2, 3, 0, 1), (4, 1, 0, 1), (4, 2, 0, 0), (8, 1, 0, 1),
                                                        (8, 1, 1, 0), (8, 2, 0, 1), (8, 3, 0, 1)])
def test_sharded_paths_match_the_reference(tmp_path, world, shard, no_ipc, lm_peer, monkeypatch):
    """shard = 1: GA rows sharded over the ranks (no_ipc = 1: through replicas + all-gather, PNOL_GA_NO_IPC=1, instead of CUDA IPC
    peer mappings); shard = 2: GA rows replicated, fitness sweep sharded; shard = 3: replicas (no collective in a generation); lm_peer = 0: the LM step's sums through NCCL
    (PNOL_LM_PEER=0) instead of the fused peer-memory kernels"""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    monkeypatch.setenv("PNOL_GA_NO_IPC", str(no_ipc))
    monkeypatch.setenv("PNOL_LM_PEER", str(lm_peer))
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), shard), nprocs=world, join=True)
    modes = {int(np.load(tmp_path / ("peer_mode_%d.npy" % r))[0]) for r in range(world)}
    assert len(modes) == 1, "every rank takes the same path"
    assert modes == ({3} if shard == 2 else {4} if shard == 3 else ({2} if no_ipc else modes)) and modes <= {1, 2, 3, 4}
    print("GA peer mode at world %d: %s" % (world, modes))
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
    GB = np.load(os.path.join(ROOT, "tests", "golden", "baseline_lm_golden.npz"))
    for tag, want, F0w in (("", G["lm_lorentz_K8/X"], G["lm_lorentz_K8/F0"]), ("K128_", GB["lm_K128/X"], GB["lm_K128/F0"])):
        Xs = [np.load(tmp_path / ("X%s%d.npy" % (tag, r))) for r in range(world)]
        for r in range(1, world):
            assert np.array_equal(Xs[0], Xs[r]), "ranks disagree on X"
        assert np.linalg.norm(Xs[0] - want) <= 1e-9 * np.linalg.norm(want)
        F0 = np.concatenate([np.load(tmp_path / ("F0%s%d.npy" % (tag or "_", r))) for r in range(world)])
        assert np.array_equal(F0, F0w)
    rep = np.load(tmp_path / "repK128_0.npy")
    assert int(rep[0]) == int(GB["lm_K128/iters"]) and rep[1] == float(GB["lm_K128/lam"])
    modes_lm = {int(np.load(tmp_path / ("repK128_%d.npy" % r))[2]) for r in range(world)}
    assert modes_lm == ({2} if lm_peer == 0 else modes_lm) and len(modes_lm) == 1 and modes_lm <= {1, 2}, "LM exchange path: %s" % modes_lm
    print("LM exchange mode at world %d: %s" % (world, modes_lm))
    gas = [np.load(tmp_path / ("gaX_%d.npy" % r)) for r in range(world)]
    for r in range(1, world):
        assert np.array_equal(gas[0], gas[r]), "default-stream GA: ranks disagree"
