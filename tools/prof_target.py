"""Small fixed workload for ncu: one LM iteration on device-resident state (pnol_lm_iterate) + one stand-alone normal-equation assembly at the cfg5 shape, one 1M x 32
Rastrigin sweep, the dense BFGS kernels at n = 4096 (p = -D g, rank-2 update), the damped solve at n = 256 and one GA generation at
1M x 32. Run plain first, then under ncu (see profiles/README.md). Everything is warmed up first; the part to be captured sits between
cuProfilerStart / cuProfilerStop (ncu --profile-from-start off)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi, problems  # noqa: E402

m = int(os.environ.get("PROF_M", 4_000_000))
K = int(os.environ.get("PROF_K", 128))
ctx = capi.Context(0)
pr = problems.lorentz_problem(m, K)
n = pr["n"]
f = ctx.functor(capi.F_LORENTZ_SUM, (pr["w"],), (), (pr["t"], pr["y"]), m)
Jd, Fd, Ft, JTJd = ctx.malloc(m * n * 8), ctx.malloc(m * 8), ctx.malloc(m * 8), ctx.malloc((n * n + n) * 8)
dx = np.full(n, 1e-7)
_, ss0 = ctx.residual_eval(f, pr["x0"], F=Fd, n=n)
chi0 = float(np.sqrt(ss0) ** 2)
# ---- warm-up of everything that will be captured ----
for _ in range(2):
    ctx.lm_iterate(f, pr["x0"].copy(), dx, n, Jd, Fd, Ft, JTJd, 1e-3, chi0, 10.0, 1)
    ctx.residual_eval(f, pr["x0"], F=Fd, n=n)
A, rhs = ctx.malloc(n * n * 8), ctx.malloc(n * 8)
ctx.lm_normal_eq(Jd, Fd, m, n, 1e-3, A=A, rhs=rhs)
B, nd = 1_000_000, 32
pts = ctx.to_device(np.random.default_rng(0).uniform(-5.12, 5.12, size=(B, nd)))
fo = ctx.malloc(B * 8)
fr = ctx.functor(capi.F_RASTRIGIN)
ctx.eval_batch(fr, pts, B, nd, f_out=fo)
n3 = 4096
rng = np.random.default_rng(9)
D = np.diag(rng.uniform(0.5, 2.0, n3))
u3 = rng.normal(size=(n3, 3)) / np.sqrt(n3)
D = D + u3 @ u3.T
g3 = rng.normal(size=n3)
s3 = 0.1 * g3 + 0.05 * rng.normal(size=n3)
Dd, gd, sd, pd = ctx.to_device(D), ctx.to_device(g3), ctx.to_device(s3), ctx.malloc(n3 * 8)
ctx.matvec_neg(Dd, gd, n3, p=pd)
ctx.bfgs_update_hinv(Dd, gd, sd, n3, mode=capi.HINV_RANK2)
M = rng.normal(size=(n, n))
As = ctx.to_device(M @ M.T / n + np.eye(n))
bs, xs = ctx.to_device(rng.normal(size=n)), ctx.malloc(n * 8)
ctx.spd_solve(As, bs, n, x=xs)
ga = ctx.ga_create(fr, nd, np.full(nd, -5.12), np.full(nd, 5.12), B, 40, dict(seed=12345, scale=1.0 - 2.0 ** -20), nstatic=1e9)
ga.init(np.full(nd, 1.0))
for _ in range(10):          # past the first generations (distribution still moving: the sort falls back to the radix kernel)
    ga.generation()
ctx.sync()

# ---- the captured part: one of each ----
cuda = ctypes.CDLL("libcuda.so.1")
cuda.cuProfilerStart()
# one LM iteration as the LM classes and bench.py run it (pnol_lm_iterate: Jacobian + J^T F, SYRK, damping, solve + trial point, trial
# residual, tail kernel = sum of squares + accept / reject, commit of F) ...
ctx.lm_iterate(f, pr["x0"].copy(), dx, n, Jd, Fd, Ft, JTJd, 1e-3, chi0, 10.0, 1)
# ... and the stand-alone normal-equation call on a given (J, F): the SYRK sums J^T F itself (extra tensor tiles)
ctx.lm_normal_eq(Jd, Fd, m, n, 1e-3, A=A, rhs=rhs)
ctx.eval_batch(fr, pts, B, nd, f_out=fo)
ctx.matvec_neg(Dd, gd, n3, p=pd)
ctx.bfgs_update_hinv(Dd, gd, sd, n3, mode=capi.HINV_RANK2)
ctx.spd_solve(As, bs, n, x=xs)
ga.generation()
ctx.sync()
cuda.cuProfilerStop()
print("prof_target done, launches", ctx.launches())
