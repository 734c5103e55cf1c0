"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU layout (parallelnonlinearoptimizationlibrary_b200/launch.py)
-- row / individual shards, the rendezvous that distributes the NCCL unique id, max-over-ranks timing -- and the algebra
the sharding relies on: the sum over ranks of the per-shard J^T J | J^T r | chi^2 equals the unsharded normal equations, so a
row-sharded LM run reproduces the single-rank iterates (checked with the CPU oracle standing in for the kernels); the coordinate-block split of the FD gradient and the individual-sharded GA
fitness sweep, each with its all-gather, are bit-identical to the unsharded results."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import oracle_lib as O
    from parallelnonlinearoptimizationlibrary_b200 import launch, problems
    r, lr, w = launch.init_process_group("gloo")
    assert (r, w) == (rank, world)
    # 1. the unique id reaches every rank unchanged
    uid = bytes(range(128)) if rank == 0 else None
    got = launch.broadcast_bytes(uid, 128, root=0)
    assert got == bytes(range(128))
    # 2. timing reduction
    assert launch.max_over_ranks(1.0 + rank) == float(world)
    # 3. row-sharded LM == unsharded LM (oracle arithmetic per shard, gloo all-reduce of the packed buffer)
    m, K = 1001, 4                      # m not divisible by the world size
    pr = problems.lorentz_problem(m, K)
    n = pr["n"]
    lo, hi = launch.row_shard(m, world, rank)
    f = O.OFunctor(103, (pr["w"],), (), (pr["t"][lo:hi], pr["y"][lo:hi]), hi - lo)
    X, lam = pr["x0"].copy(), 1e-3
    dx = np.full(n, 1e-7)
    F = O.residual(f, X)
    chi = float(launch.sum_over_ranks(np.array([np.sum(F * F)]))[0])
    for _ in range(6):
        J, _ = O.fd_jacobian(f, X, dx)
        JTJ, _, rhs = O.lm_normal_eq(J, F, 0.0)
        packed = launch.sum_over_ranks(np.concatenate([JTJ.ravel(), rhs]))
        JTJ, rhs = packed[:n * n].reshape(n, n), packed[n * n:]
        A = JTJ.copy()
        A[np.diag_indices(n)] = (1 + lam) * np.diag(JTJ)
        sigma = O.lu_solve(A, rhs)
        Xn = X + sigma
        Fn = O.residual(f, Xn)
        chin = float(launch.sum_over_ranks(np.array([np.sum(Fn * Fn)]))[0])
        if chin >= chi or chin != chin:
            lam *= 10
        else:
            lam /= 10
            X, F, chi = Xn, Fn, chin
    np.save(os.path.join(out_dir, "X%d.npy" % rank), X)
    # 4. FD gradient split by coordinate blocks: every rank evaluates its block of stencil points, one all-gather of
    #    fixed-size blocks in rank order rebuilds the gradient -- bit-identical to the unsharded stencil
    import torch
    fr = O.OFunctor(1)                  # Rosenbrock
    ng = 11                             # not divisible by the world size
    xg, dxg = np.linspace(-1.5, 2.0, ng), np.full(ng, 1e-6)
    g_all, f0 = O.fd_gradient(fr, xg, dxg)
    clo, chi_ = launch.column_shard(ng, world, rank)
    per = (ng + world - 1) // world
    mine = torch.zeros(per, dtype=torch.float64)
    for i in range(clo, chi_):          # this rank's stencil points only
        xp = xg.copy()
        xp[i] = xp[i] + dxg[i]
        mine[i - clo] = (O.obj_eval(fr, xp) - f0) / dxg[i]
    blocks = [torch.zeros(per, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(blocks, mine)
    g = torch.cat(blocks).numpy()[:ng]
    assert np.array_equal(g, g_all)
    # 5. GA fitness sweep sharded by individuals, all-gather of F (every other GA stage is replicated): bit-identical F on every rank
    fa = O.OFunctor(5)                  # Rastrigin
    npop, ngene = 101, 8
    pop = np.random.default_rng(11).uniform(-5.12, 5.12, size=(npop, ngene))
    plo, phi = launch.row_shard(npop, world, rank)
    perp = (npop + world - 1) // world
    minef = torch.zeros(perp, dtype=torch.float64)
    minef[:phi - plo] = torch.from_numpy(O.eval_batch(fa, pop[plo:phi]))
    fblocks = [torch.zeros(perp, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(fblocks, minef)
    F_all = np.concatenate([fblocks[r].numpy()[:launch.row_shard(npop, world, r)[1] - launch.row_shard(npop, world, r)[0]] for r in range(world)])
    assert np.array_equal(F_all, O.eval_batch(fa, pop))
    dist.barrier()
    dist.destroy_process_group()


def test_column_shards_are_equal_blocks_in_rank_order():
    from parallelnonlinearoptimizationlibrary_b200 import launch
    for n in (1, 2, 10, 11, 256, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            per = (n + world - 1) // world
            edges = [launch.column_shard(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert all(b - a == per for a, b in edges if b < n)          # full blocks everywhere before the cut
    with pytest.raises(ValueError):
        launch.column_shard(10, 2, 2)


def test_row_shards_cover_everything():
    from parallelnonlinearoptimizationlibrary_b200 import launch
    for m in (0, 1, 7, 8, 4_000_000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            edges = [launch.row_shard(m, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == m
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        launch.row_shard(10, 2, 2)


def test_world_size_2_gloo(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from parallelnonlinearoptimizationlibrary_b200 import problems
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    X0, X1 = np.load(tmp_path / "X0.npy"), np.load(tmp_path / "X1.npy")
    assert np.array_equal(X0, X1)                       # every rank holds the same iterate
    pr = problems.lorentz_problem(1001, 4)
    f = O.OFunctor(103, (pr["w"],), (), (pr["t"], pr["y"]), 1001)
    w = O.lm(f, pr["x0"], 0.001, 10.0, 1e-7, 6, 0.0)
    assert np.linalg.norm(X0 - w["X"]) <= 1e-9 * np.linalg.norm(w["X"])
