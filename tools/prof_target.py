"""Small fixed workload for ncu: LM iterations (pnol_lm_step) + one stand-alone normal-equation assembly at the cfg5 shape and one 1M x 32
Rastrigin sweep. Run plain first, then under ncu (see profiles/README.md)."""
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
ctx.residual_eval(f, pr["x0"], F=Fd, n=n)
# one LM iteration as the LM classes and bench.py run it (Jacobian + J^T F, SYRK, damped solve, trial residual) ...
for _ in range(2):
    ctx.lm_step(f, pr["x0"], dx, n, Jd, Fd, Ft, 1e-3, JTJd)
# ... and the stand-alone normal-equation call on a given (J, F): the SYRK sums J^T F itself (extra tensor tiles)
A, rhs = ctx.malloc(n * n * 8), ctx.malloc(n * 8)
ctx.lm_normal_eq(Jd, Fd, m, n, 1e-3, A=A, rhs=rhs)
B, nd = 1_000_000, 32
pts = ctx.to_device(np.random.default_rng(0).uniform(-5.12, 5.12, size=(B, nd)))
fo = ctx.malloc(B * 8)
fr = ctx.functor(capi.F_RASTRIGIN)
for _ in range(2):
    ctx.eval_batch(fr, pts, B, nd, f_out=fo)
ctx.sync()
print("prof_target done, launches", ctx.launches())
