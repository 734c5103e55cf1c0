"""GPU: the BASELINE.json shapes at FULL size, checked through size-independent properties and through the oracle on random
samples of rows / individuals (the oracle cannot run the whole problem in seconds -- SURVEY.md Appendix C):
  cfg5  m = 4M x n = 256: sampled Jacobian rows and residuals bit-exact; J^T J symmetric, additive over row blocks (the property
        the multi-GPU all-reduce relies on) and equal to the oracle on a 3000-row block; LM steps decrease chi^2.
  cfg4  Npop = 1M x 32: F sorted after every generation, F[i] == f(Xpop[i]) on samples, elites survive, the random stream advances
        by at least the draws the fixed-cost stages need.
  cfg3  n = 4096: rank-2 and literal (two DMMA GEMMs) inverse-Hessian updates agree; p = -D g against numpy."""
import numpy as np
import pytest

import oracle_lib as O
from parallelnonlinearoptimizationlibrary_b200 import capi, problems

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def test_cfg5_jacobian_and_normal_equations(ctx):
    m, K = 4_000_000, 128
    pr = problems.lorentz_problem(m, K)
    n = pr["n"]
    f = ctx.functor(capi.F_LORENTZ_SUM, (pr["w"],), (), (pr["t"], pr["y"]), m)
    Jd, Fd = ctx.malloc(m * n * 8), ctx.malloc(m * 8)
    dx = np.full(n, 1e-7)
    ctx.fd_jacobian(f, pr["x0"], dx, J=Jd, F=Fd, n=n)
    # sampled rows against the oracle (a functor over just those rows evaluates the same arithmetic)
    rng = np.random.default_rng(5)
    rows = np.sort(np.concatenate([[0, 1, 31, 32, m - 33, m - 1], rng.integers(0, m, 250)]))
    of = O.OFunctor(capi.F_LORENTZ_SUM, (pr["w"],), (), (pr["t"][rows], pr["y"][rows]), rows.size)
    Jw, Fw = O.fd_jacobian(of, pr["x0"], dx)
    Jrow = np.empty(n)
    Fi = np.empty(1)
    for k, i in enumerate(rows):
        ctx.memcpy(Jrow, Jd + int(i) * n * 8, n * 8)
        ctx.memcpy(Fi, Fd + int(i) * 8, 8)
        assert np.array_equal(Jrow, Jw[k]) and Fi[0] == Fw[k], "row %d differs from the oracle" % i
    # normal equations: full, and as the sum over two row blocks (what two ranks would all-reduce)
    JTJ, A, rhs = ctx.lm_normal_eq(Jd, Fd, m, n, 1e-3)
    assert np.array_equal(JTJ, JTJ.T)
    assert np.array_equal(np.diag(A), (1 + 1e-3) * np.diag(JTJ)) and np.array_equal(A - np.diag(np.diag(A)), JTJ - np.diag(np.diag(JTJ)))
    h = 1_999_968
    J1, A1, r1 = ctx.lm_normal_eq(Jd, Fd, h, n, 0.0)
    J2, A2, r2 = ctx.lm_normal_eq(Jd + h * n * 8, Fd + h * 8, m - h, n, 0.0)
    assert rel(J1 + J2, JTJ) < 1e-12 and rel(r1 + r2, rhs) < 1e-12
    # a 3000-row block against the sequential oracle
    blk = 3000
    Jb = np.empty((blk, n))
    Fb = np.empty(blk)
    ctx.memcpy(Jb, Jd + 1_000_000 * n * 8, blk * n * 8)
    ctx.memcpy(Fb, Fd + 1_000_000 * 8, blk * 8)
    Jg, _, rg = ctx.lm_normal_eq(Jd + 1_000_000 * n * 8, Fd + 1_000_000 * 8, blk, n, 0.0)
    Jo, _, ro = O.lm_normal_eq(Jb, Fb, 0.0)
    assert rel(Jg, Jo) < 1e-12 and rel(rg, ro) < 1e-12
    # three LM steps from x0 decrease chi^2 (zero-noise data, start 10 % off)
    X, lam = pr["x0"].copy(), 1e-3
    _, chi = ctx.residual_eval(f, X, F=Fd, n=n)
    for _ in range(3):
        ctx.fd_jacobian(f, X, dx, J=Jd, F=None, n=n)
        _, A, rhs = ctx.lm_normal_eq(Jd, Fd, m, n, lam)
        X = X + ctx.spd_solve(A, rhs, n)
        _, chin = ctx.residual_eval(f, X, F=Fd, n=n)
        assert chin < chi
        chi, lam = chin, lam / 10
    assert rel(X, pr["x_true"]) < 1e-3
    ctx.free(Jd)
    ctx.free(Fd)


def test_cfg4_ga_generation_invariants(ctx):
    npop, n = 1_000_000, 32
    f = ctx.functor(capi.F_RASTRIGIN)
    of = O.OFunctor(capi.F_RASTRIGIN)
    lb, ub = np.full(n, -5.12), np.full(n, 5.12)
    ga = ctx.ga_create(f, n, lb, ub, npop, 10, dict(seed=12345, scale=1.0 - 2.0 ** -20), nstatic=1e9)
    x0 = np.full(n, 2.5)
    f0 = ga.init(x0)
    assert f0 == O.obj_eval(of, x0)
    rng = np.random.default_rng(1)
    prev_best, prev_pos = None, ga.status().stream_pos
    for gen in range(3):
        if gen:
            ga.generation()
        X, F = ga.population()
        st = ga.status()
        assert np.all(F[1:] >= F[:-1]), "population not sorted by objective"
        assert np.all(X >= -5.12) and np.all(X <= 5.12)
        idx = np.concatenate([[0, 1, npop - 1], rng.integers(0, npop, 500)])
        assert np.array_equal(F[idx], O.eval_batch(of, X[idx])), "F[i] != f(Xpop[i])"
        if prev_best is not None:
            assert F[0] <= prev_best                         # elites are copied unchanged, so the best never gets worse
            # fixed-cost draws of a generation: mutation genes of the Nrand children + 2 draws per elite-mutation gene
            assert st.stream_pos - prev_pos >= (st.n_rand + 2 * st.n_elite_mut) * n + 2 * st.n_cross * n
        prev_best, prev_pos = F[0], st.stream_pos
    assert (st.n_elite, st.n_elite_mut, st.n_cross, st.n_rand) == (100000, 200000, 300000, 400000)
    ga.close()


def test_cfg3_dense_updates_at_n4096(ctx):
    n = 4096
    rng = np.random.default_rng(9)
    d = rng.uniform(0.5, 2.0, n)
    u = rng.normal(size=(n, 3)) / np.sqrt(n)
    D = np.diag(d) + u @ u.T
    g = rng.normal(size=n)
    s = 0.1 * g + 0.05 * rng.normal(size=n)
    Dd = ctx.to_device(D)
    p = ctx.matvec_neg(Dd, g, n)
    assert rel(p, -(D @ g)) < 1e-13
    D1, D2 = ctx.to_device(D), ctx.to_device(D)
    ctx.bfgs_update_hinv(D1, g, s, n, mode=capi.HINV_RANK2)
    ctx.bfgs_update_hinv(D2, g, s, n, mode=capi.HINV_LITERAL)
    A, B = ctx.to_host(D1, (n, n)), ctx.to_host(D2, (n, n))
    assert rel(A, B) < 1e-12
    rho = 1.0 / (g @ s)
    want = (np.eye(n) - rho * np.outer(s, g)) @ D @ (np.eye(n) - rho * np.outer(g, s)) + rho * np.outer(s, s)
    assert rel(B, want) < 1e-12
    # the secant condition of the BFGS inverse update: D_new g = s
    assert rel(B @ g, s) < 1e-9
    for q in (Dd, D1, D2):
        ctx.free(q)
