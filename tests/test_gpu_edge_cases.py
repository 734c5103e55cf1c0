"""GPU: edge cases of the C-ABI -- empty and one-element inputs, ragged sizes around the warp / tile boundaries, invalid
arguments (status codes instead of the reference's exit()), non-finite objectives, a singular damped matrix in LM."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from parallelnonlinearoptimizationlibrary_b200 import capi, problems

pytestmark = pytest.mark.gpu


def test_empty_and_single_row_batches(ctx):
    f = ctx.functor(capi.F_ROSENBROCK)
    of = O.OFunctor(capi.F_ROSENBROCK)
    assert ctx.eval_batch(f, np.zeros((0, 4)), 0, 4).size == 0                   # B = 0 is a no-op, not an error
    for B in (1, 2, 31, 32, 33, 127, 128, 129):
        pts = np.random.default_rng(B).uniform(-2, 2, size=(B, 7))
        assert np.array_equal(ctx.eval_batch(f, pts, B, 7), O.eval_batch(of, pts))


@pytest.mark.parametrize("B,n", [(1, 32), (63, 32), (64, 32), (65, 32), (1000, 4), (1000, 36), (257, 100), (200000, 32), (100001, 64), (3000, 96), (5000, 48)])
def test_separable_sweep_sizes(ctx, B, n):
    # Rastrigin takes the row-wise / warp-tile kernels for aligned n and the generic tile kernel otherwise: same bits everywhere
    f = ctx.functor(capi.F_RASTRIGIN)
    of = O.OFunctor(capi.F_RASTRIGIN)
    pts = np.random.default_rng(B + n).uniform(-5.12, 5.12, size=(B, n))
    assert np.array_equal(ctx.eval_batch(f, pts, B, n), O.eval_batch(of, pts))
    ind = (np.arange(B) % 3 != 0).astype(np.uint8)
    out = np.full(B, -1.0)
    got = ctx.eval_batch(f, pts, B, n, indicator=ind, f_out=out.copy())
    want = O.eval_batch(of, pts, indicator=ind, f_out=out.copy())
    assert np.array_equal(got, want) and np.all(got[ind == 0] == -1.0)


def test_strided_population(ctx):
    # ld > n: rows of a wider matrix (the GA never needs it, bindings may)
    f = ctx.functor(capi.F_RASTRIGIN)
    of = O.OFunctor(capi.F_RASTRIGIN)
    wide = np.random.default_rng(3).uniform(-5, 5, size=(300, 40))
    got = ctx.eval_batch(f, wide, 300, 32, ld=40)
    assert np.array_equal(got, O.eval_batch(of, np.ascontiguousarray(wide[:, :32])))


@pytest.mark.parametrize("m", [1, 2, 31, 33, 95])
def test_tiny_residual_problems(ctx, m):
    pr = problems.lorentz_problem(max(m, 2), 4)
    t, y = pr["t"][:m], pr["y"][:m]
    f = ctx.functor(capi.F_LORENTZ_SUM, (pr["w"],), (), (t, y), m)
    of = O.OFunctor(capi.F_LORENTZ_SUM, (pr["w"],), (), (t, y), m)
    dx = np.full(pr["n"], 1e-7)
    J, F = ctx.fd_jacobian(f, pr["x0"], dx)
    Jw, Fw = O.fd_jacobian(of, pr["x0"], dx)
    assert np.array_equal(J, Jw) and np.array_equal(F, Fw)
    Jb, Fb = ctx.fd_jacobian(f, pr["x0"], dx, mode=capi.JAC_BLACKBOX)
    assert np.array_equal(Jb, Jw) and np.array_equal(Fb, Fw)
    JTJ, A, rhs = ctx.lm_normal_eq(J, F, m, pr["n"], 0.5)
    JTJw, Aw, rhsw = O.lm_normal_eq(Jw, Fw, 0.5)
    assert np.allclose(JTJ, JTJw, rtol=1e-12, atol=1e-300) and np.allclose(rhs, rhsw, rtol=1e-12, atol=1e-300)


def test_one_parameter_gradient_and_hessian(ctx):
    f = ctx.functor(capi.F_POWER, (), (3,))
    of = O.OFunctor(capi.F_POWER, (), (3,))
    x, dx = np.array([1.7]), np.array([1e-6])
    g, f0 = ctx.fd_gradient(f, x, dx)
    gw, f0w = O.fd_gradient(of, x, dx)
    assert np.array_equal(g, gw) and f0 == f0w
    assert np.array_equal(ctx.fd_hessian(f, x, np.array([1e-3])), O.fd_hessian(of, x, np.array([1e-3])))


@pytest.mark.parametrize("nfree", [0, 1, 3])
def test_recur_gradient_with_more_reduced_entries_than_free_variables(ctx, nfree):
    # BFGSBnd_MPI keeps calling the Recur stencils with the FULL vector after variables (even all of them) have been frozen
    # (Source/BFGS_with_bnd_linsearch_MPI.cpp:808-810): objEvalRecur reads only as many reduced entries as there are free slots
    # (Source/PNOL_Objective.cpp:311-323), the others never reach the objective and their gradient entries are exactly 0
    n = 6
    rng = np.random.default_rng(40 + nfree)
    constx = rng.uniform(-1, 1, n)
    ind = np.ones(n, dtype=np.uint8)
    ind[:nfree] = 0
    xr = rng.uniform(-1, 1, n)                                 # n reduced entries for nfree free slots
    dxr = np.full(n, 1e-6)
    for kind in (capi.F_ROSENBROCK, capi.F_RASTRIGIN):
        f, of = ctx.functor(kind), O.OFunctor(kind)
        g, f0 = ctx.fd_gradient_recur(f, xr, dxr, constx, ind)
        gw, f0w = O.fd_gradient_recur(of, xr, dxr, constx, ind)
        assert f0 == f0w and np.array_equal(g, gw) and np.all(g[nfree:] == 0.0)
        assert ctx.eval_recur(f, xr, constx, ind) == O.eval_recur(of, xr, constx, ind)
    # the same situation through the alpha pool: only the free slots move along p
    f, of = ctx.functor(capi.F_ROSENBROCK), O.OFunctor(capi.F_ROSENBROCK)
    p = rng.normal(size=n)
    alpha = np.array([0.0, 0.1, 0.5, 1.0])
    phi, _, bad = ctx.alpha_pool(f, xr, p, alpha, 0.0, want_dphi=False, const_x=constx, const_ind=ind)
    want = [O.eval_recur(of, xr + a * p, constx, ind) for a in alpha]
    assert bad == 0 and np.array_equal(phi, np.array(want))


def test_invalid_arguments_return_status_codes(ctx):
    lib = ctx.lib
    f = ctx.functor(capi.F_ROSENBROCK)
    x = np.ones(4)
    out = np.zeros(4)
    # null pointers / bad sizes: PNOL_ERR_INVALID (1), never a crash or an exit()
    assert lib.pnol_fd_gradient(ctx.h, f.handle, None, C.c_void_p(x.ctypes.data), 4, C.c_void_p(out.ctypes.data), None) == 1
    assert lib.pnol_eval_batch(ctx.h, f.handle, C.c_void_p(x.ctypes.data), C.c_longlong(1), 0, C.c_longlong(0), None, C.c_void_p(out.ctypes.data)) == 1
    assert lib.pnol_spd_solve(ctx.h, C.c_void_p(x.ctypes.data), C.c_void_p(x.ctypes.data), 0, C.c_void_p(out.ctypes.data), None) == 1
    assert b"spd_solve" in lib.pnol_last_error(ctx.h)
    # a scalar functor handed to a residual entry point, and the reverse: PNOL_ERR_NO_FUNCTOR (3)
    F = np.zeros(4)
    assert lib.pnol_residual_eval(ctx.h, f.handle, C.c_void_p(x.ctypes.data), 4, C.c_void_p(F.ctypes.data), None) == 3
    with pytest.raises(capi.PnolError):
        ctx.functor(999)
    with pytest.raises(capi.PnolError):
        ctx.functor(capi.F_LORENTZ_SUM, (4.0,), (), (), 10)          # data columns missing
    # the sum-of-Lorentzians model adds its terms in a balanced binary tree: K = n / 2 must be a power of two
    pr = problems.lorentz_problem(64, 4)
    fl = ctx.functor(capi.F_LORENTZ_SUM, (pr["w"],), (), (pr["t"], pr["y"]), 64)
    for bad_n in (6, 7, 10):
        with pytest.raises(capi.PnolError, match="not 2"):
            ctx.residual_eval(fl, np.ones(bad_n))
        with pytest.raises(capi.PnolError, match="not 2"):
            ctx.fd_jacobian(fl, np.ones(bad_n), np.full(bad_n, 1e-7))
        with pytest.raises(capi.PnolError, match="not 2"):
            ctx.lm_normal_eq_fused(fl, np.ones(bad_n), np.full(bad_n, 1e-7), bad_n, 0.1)
    with pytest.raises(capi.PnolError):
        ctx.ga_create(f, 4, np.zeros(4), np.ones(4), 1, 5, dict(seed=1, scale=0.5))    # npop < 2
    with pytest.raises(capi.PnolError):
        ctx.ga_create(f, 4, np.zeros(4), np.ones(4), 100, 5, dict(seed=1, scale=0.5), elite_frac=0.6, cross_frac=0.6)   # no room for random children


def test_nonfinite_objective_values(ctx):
    f = ctx.functor(capi.F_ROSENBROCK)
    of = O.OFunctor(capi.F_ROSENBROCK)
    pts = np.array([[1.0, 2.0, 3.0], [np.nan, 1.0, 1.0], [1e200, 1.0, 1.0], [np.inf, 0.0, 0.0]])
    got, want = ctx.eval_batch(f, pts, 4, 3), O.eval_batch(of, pts)
    assert np.array_equal(got, want, equal_nan=True)
    g, f0 = ctx.fd_gradient(f, pts[2], np.full(3, 1e-6))
    gw, f0w = O.fd_gradient(of, pts[2], np.full(3, 1e-6))
    assert np.array_equal(g, gw, equal_nan=True) and (f0 == f0w or (f0 != f0 and f0w != f0w))


def test_lm_with_a_singular_normal_matrix(ctx):
    # an amplitude of exactly 0 makes the column of its centre vanish: J^T J is singular, Marquardt's damping multiplies a zero
    # diagonal, the solve reports a non-positive pivot, the step is treated like the NaN step of the reference (rejected)
    from parallelnonlinearoptimizationlibrary_b200 import hostapi
    hostapi.attach(ctx)
    try:
        pr = problems.lorentz_problem(500, 4)
        x0 = pr["x0"].copy()
        x0[2] = 0.0
        r = hostapi.lm_lorentz(pr["t"], pr["y"], pr["w"], x0, 0.001, 10.0, 1e-7, 5, 0.0)
        assert r["iterations"] == 5 and r["accepted"] == 0 and r["rejected"] == 5
        assert np.array_equal(r["X"], x0) and np.array_equal(r["F"], r["F0"])
        f = ctx.functor(capi.F_LORENTZ_SUM, (pr["w"],), (), (pr["t"], pr["y"]), 500)
        J, F = ctx.fd_jacobian(f, x0, np.full(8, 1e-7))
        assert np.all(J[:, 3] == 0.0)
        _, A, rhs = ctx.lm_normal_eq(J, F, 500, 8, 1e-3)
        with pytest.raises(capi.PnolError, match="NOT_SPD"):
            ctx.spd_solve(A, rhs, 8)
    finally:
        hostapi.detach()


def test_large_pageable_copies_round_trip(ctx):
    # copies of >= 8 MB between pageable host memory and the device go through the threaded, chunked staging path (capi.cu:
    # staged_copy): whole chunks, a ragged tail, fewer chunks than threads, and the plain path below the threshold
    rng = np.random.default_rng(11)
    for nbytes in (8 << 20, (8 << 20) + 8, (36 << 20) + 4096 + 8, 4 << 20, (64 << 20)):
        a = rng.integers(0, 2 ** 62, size=nbytes // 8, dtype=np.int64)
        p = ctx.to_device(a)
        b = ctx.to_host(p, a.shape, dtype=np.int64)
        assert np.array_equal(a, b)
        ctx.free(p)


@pytest.mark.parametrize("count", [0, 1000, 1 << 17, (1 << 21) + 12345])
def test_copy_beside_the_stream(ctx, count):
    # pnol_copy_start / pnol_copy_wait (findMin's F0 read-back beside the iterations): small copies complete at once, from 1 MB on a
    # worker thread drives the staged copy while the context's stream works; every other copy waits for the one in flight
    src = np.random.default_rng(count).normal(size=count)
    dev = ctx.to_device(src) if count else ctx.malloc(8)
    dst = np.zeros(count)
    f = ctx.functor(capi.F_RASTRIGIN)
    pts = np.random.default_rng(1).uniform(-5.12, 5.12, size=(4096, 32))
    ctx.copy_start(dst, dev, dst.nbytes)
    got = ctx.eval_batch(f, pts, 4096, 32)                      # work on the context's stream meanwhile (its own small copies join the worker)
    ctx.copy_wait()
    assert np.array_equal(dst, src)
    assert np.array_equal(got, O.eval_batch(O.OFunctor(capi.F_RASTRIGIN), pts))
    ctx.copy_wait()                                            # nothing in flight: a no-op
    ctx.free(dev)


def test_timer_modes(ctx):
    # pnol_timer_enable: 0 off, 1 every scope, 2 the kernels that carry an LM iteration only (what bench.py's timed region uses)
    pr = problems.lorentz_problem(2048, 8)
    f = ctx.functor(capi.F_LORENTZ_SUM, (pr["w"],), (), (pr["t"], pr["y"]), pr["m"])
    for mode, want_sumsq in ((1, True), (2, False), (0, None)):
        ctx.timer_enable(mode)
        ctx.timer_reset()
        ctx.residual_eval(f, pr["x0"], n=pr["n"])
        _, c_res = ctx.timer_get("residual")
        _, c_ss = ctx.timer_get("sumsq")
        ctx.timer_enable(False)
        if mode == 0:
            assert c_res == 0 and c_ss == 0
        else:
            assert c_res == 1 and (c_ss == 1) == want_sumsq
