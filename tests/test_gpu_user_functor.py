"""GPU: objectives that are NOT built into libpnol_b200.so (tests/user_functor/my_objectives.{hpp,cu}, compiled out of tree by
user_functor_lib.build_and_load) run through every entry point that takes a functor and match the host objEval of the same
__host__ __device__ source bit for bit -- the bar the built-in objectives meet against the oracle. The reference's plug-in point is
subclassing Objective / MultiObjective (Source/PNOL_Objective.hpp:29, :57, Source/ExampleObjectives.hpp)."""
import ctypes as C

import numpy as np
import pytest

import user_functor_lib as U
from parallelnonlinearoptimizationlibrary_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def user():
    return U.build_and_load()


def host_eval(u, kind, scale, pts):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    return np.array([u.my_host_eval(kind, C.c_double(scale), C.c_void_p(row.ctypes.data), row.size) for row in pts])


def host_residual(u, t, y, c0, x):
    F = np.empty(t.size)
    u.my_host_residual(C.c_void_p(t.ctypes.data), C.c_void_p(y.ctypes.data), C.c_longlong(t.size), C.c_double(c0), C.c_void_p(x.ctypes.data),
                       C.c_void_p(F.ctypes.data))
    return F


@pytest.mark.parametrize("kind,scale,n,B", [(U.MY_F_TRID, 1.0, 6, 300), (U.MY_F_TRID, 1.0, 33, 1000), (U.MY_F_STYBLINSKI, 0.5, 32, 4096),
                                            (U.MY_F_STYBLINSKI, 0.5, 7, 100), (U.MY_F_STYBLINSKI, 2.0, 48, 777)])
def test_user_scalar_objective_batch_sweep_bit_exact(ctx, user, kind, scale, n, B):
    rng = np.random.default_rng(kind + n)
    pts = rng.uniform(-4.0, 4.0, size=(B, n))
    f = ctx.functor(kind, (scale,))
    got = ctx.eval_batch(f, pts, B, n)
    assert np.array_equal(got, host_eval(user, kind, scale, pts))
    ind = (rng.uniform(size=B) < 0.5).astype(np.uint8)          # the GA's evaluateIndicator
    got2 = ctx.eval_batch(f, pts, B, n, indicator=ind, f_out=None)
    assert np.array_equal(got2[ind != 0], got[ind != 0])


@pytest.mark.parametrize("kind,scale,n", [(U.MY_F_TRID, 1.0, 10), (U.MY_F_STYBLINSKI, 0.5, 17)])
def test_user_scalar_objective_stencils_bit_exact(ctx, user, kind, scale, n):
    rng = np.random.default_rng(n)
    x = rng.uniform(-2.0, 2.0, size=n)
    dx = np.full(n, 1e-7)
    f = ctx.functor(kind, (scale,))
    g, f0 = ctx.fd_gradient(f, x, dx)
    pert = np.tile(x, (n, 1))
    pert[np.arange(n), np.arange(n)] = x + dx                    # XdX[i] = XdX[i] + dX[i]   (Source/PNOL_Objective.cpp:27)
    fp = host_eval(user, kind, scale, pert)
    f0w = host_eval(user, kind, scale, x[None, :])[0]
    assert f0 == f0w and np.array_equal(g, (fp - f0w) / dx)      # (Source/PNOL_Objective.cpp:31)
    # FD Hessian (Source/PNOL_Objective.cpp:38-85)
    dxh = np.full(n, 1e-3)
    Bm = ctx.fd_hessian(f, x, dxh)
    fi = host_eval(user, kind, scale, np.tile(x, (n, 1)) + np.diag(dxh))
    want = np.empty((n, n))
    for i in range(n):
        for j in range(i, n):
            xij = x.copy()
            xij[i] = xij[i] + dxh[i]
            xij[j] = xij[j] + dxh[j]
            fij = host_eval(user, kind, scale, xij[None, :])[0]
            want[i, j] = want[j, i] = (fij - fi[i] - fi[j] + f0w) / (dxh[i] * dxh[j])
    assert np.array_equal(Bm, want)
    # alpha pool of the pooled line searches (Source/BFGS_bnd_linesearch_MPI_SW.cpp:599-734)
    p = -g / np.linalg.norm(g)
    alpha = np.linspace(0.0, 1.0, 6)
    phi, dphi, bad = ctx.alpha_pool(f, x, p, alpha, 1e-6)
    pw = host_eval(user, kind, scale, x[None, :] + alpha[:, None] * p[None, :])
    pw2 = host_eval(user, kind, scale, x[None, :] + (alpha + 1e-6)[:, None] * p[None, :])
    assert bad == 0 and np.array_equal(phi, pw) and np.array_equal(dphi, (pw2 - pw) / 1e-6)


def _gauss_problem(m):
    t = np.linspace(-3.0, 5.0, m)
    truth = np.array([2.5, 1.2, 0.8])
    y = truth[0] * np.exp(-(t - truth[1]) ** 2 * truth[2]) + 0.3
    return t, y, 0.3, truth


@pytest.mark.parametrize("m", [1, 257, 5000])
def test_user_residual_model_and_jacobian_bit_exact(ctx, user, m):
    t, y, c0, truth = _gauss_problem(m)
    x = truth * np.array([1.1, 0.9, 1.2])
    f = ctx.functor(U.MY_F_GAUSSFIT, (c0,), (), (t, y), m)
    F, ss = ctx.residual_eval(f, x)
    Fw = host_residual(user, t, y, c0, x)
    assert np.array_equal(F, Fw)
    dx = np.full(3, 1e-7)
    J, F2 = ctx.fd_jacobian(f, x, dx)
    Jw = np.empty((m, 3))
    for j in range(3):
        xj = x.copy()
        xj[j] = xj[j] + dx[j]                                    # (Source/PNOL_Objective.cpp:186)
        Jw[:, j] = (host_residual(user, t, y, c0, xj) - Fw) / dx[j]      # (:192)
    assert np.array_equal(F2, Fw) and np.array_equal(J, Jw)


def test_user_objectives_through_the_plugin_classes(ctx, user):
    """BFGS::findMin and LevMarq::findMin (the reference's classes, include/pnol) on user objectives: the drivers in
    my_objectives.cu are what a user of the reference writes; only deviceFunctor() is new"""
    from parallelnonlinearoptimizationlibrary_b200 import hostapi
    hostapi.attach(ctx)
    n = 8
    x = np.full(n, 0.5)
    f0, fopt = C.c_double(), C.c_double()
    l0 = ctx.launches()
    assert user.my_bfgs_trid(C.c_void_p(x.ctypes.data), n, 200, C.byref(f0), C.byref(fopt)) == 0
    assert ctx.launches() > l0                                   # the user's kernels ran on this context
    i = np.arange(1, n + 1)
    assert np.allclose(x, i * (n + 1 - i), rtol=1e-4)           # known minimiser of the Trid function
    assert abs(fopt.value - (-n * (n + 4) * (n - 1) / 6.0)) < 1e-6 and fopt.value < f0.value
    t, y, c0, truth = _gauss_problem(400)
    xs = truth * np.array([1.2, 0.9, 1.1])
    assert user.my_lm_gaussfit(C.c_void_p(t.ctypes.data), C.c_void_p(y.ctypes.data), C.c_longlong(t.size), C.c_double(c0), C.c_void_p(xs.ctypes.data), 50) == 0
    assert np.allclose(xs, truth, rtol=1e-6)


def test_user_objective_in_the_ga(ctx, user):
    """GeneticAlgorithm state machine (pnol_ga_*) with a user objective: the sweep inside every generation launches the user's kernel"""
    n, npop = 4, 2000
    f = ctx.functor(U.MY_F_STYBLINSKI, (1.0,))
    gas = ctx.ga_create(f, n, np.full(n, -5.0), np.full(n, 5.0), npop, 20, dict(seed=7, scale=1.0 - 2.0 ** -20), nstatic=1e9)
    gas.init(np.zeros(n))
    for _ in range(15):
        gas.generation()
    X, F = gas.population()
    assert np.array_equal(F, host_eval(user, U.MY_F_STYBLINSKI, 1.0, X))      # F[i] = f(X[i]) for the whole population
    assert np.all(np.diff(F) >= 0) and F[0] < -100.0                          # sorted; near the global minimum -39.166 n
    gas.close()


def test_unregistered_kind_is_refused(ctx):
    with pytest.raises(capi.PnolError):
        ctx.functor(1777)
