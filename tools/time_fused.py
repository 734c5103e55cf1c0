"""Times pnol_lm_normal_eq_fused (J never stored, row blocks) against the two-kernel path at the cfg5 shape.
PNOL_FUSED_MB = MB of J per block (default 512). Tuning helper."""
import os
import sys
import time

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
JTJ, A, rhs, Fd = ctx.malloc(n * n * 8), ctx.malloc(n * n * 8), ctx.malloc(n * 8), ctx.malloc(m * 8)
dx = np.full(n, 1e-7)
for mb in [int(v) for v in os.environ.get("SWEEP_MB", "16,32,64,128").split(",")]:
    os.environ["PNOL_FUSED_MB"] = str(mb)
    for _ in range(2):
        ctx.lm_normal_eq_fused(f, pr["x0"], dx, n, 1e-3, JTJ=JTJ, A=A, rhs=rhs, F=Fd)
    ctx.sync()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ctx.lm_normal_eq_fused(f, pr["x0"], dx, n, 1e-3, JTJ=JTJ, A=A, rhs=rhs, F=Fd)
    ctx.sync()
    print("fused  m=%d n=%d  block %3d MB of J: %.3f ms per call" % (m, n, mb, (time.perf_counter() - t0) / reps * 1e3))
# the stored-J path on the same problem: FD Jacobian (J and F written) + SYRK with F
Jd = ctx.malloc(m * n * 8)
for _ in range(2):
    ctx.fd_jacobian(f, pr["x0"], dx, J=Jd, F=Fd)
    ctx.lm_normal_eq(Jd, Fd, m, n, 1e-3, JTJ=JTJ, A=A, rhs=rhs)
ctx.sync()
t0 = time.perf_counter()
for _ in range(5):
    ctx.fd_jacobian(f, pr["x0"], dx, J=Jd, F=Fd)
    ctx.lm_normal_eq(Jd, Fd, m, n, 1e-3, JTJ=JTJ, A=A, rhs=rhs)
ctx.sync()
print("stored m=%d n=%d  J = %.2f GB: %.3f ms per call" % (m, n, m * n * 8 / 1e9, (time.perf_counter() - t0) / 5 * 1e3))
# one whole LM iteration (pnol_lm_step: normal equations, damped solve, trial residuals, one host synchronisation) with and without a J buffer
Ft, JTJp = ctx.malloc(m * 8), ctx.malloc((n * n + n) * 8)
ctx.residual_eval(f, pr["x0"], F=Fd, n=n)
os.environ["PNOL_FUSED_MB"] = os.environ.get("STEP_MB", "512")
for label, Jarg in (("stored J", Jd), ("no J, %s MB blocks" % os.environ["PNOL_FUSED_MB"], None)):
    for _ in range(2):
        ctx.lm_step(f, pr["x0"], dx, n, Jarg, Fd, Ft, 1e-3, JTJp)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.lm_step(f, pr["x0"], dx, n, Jarg, Fd, Ft, 1e-3, JTJp)
    ctx.sync()
    dt = (time.perf_counter() - t0) / 10
    print("lm_step m=%d n=%d  %-24s %.3f ms per iteration = %.1f LM iterations/s" % (m, n, label + ":", dt * 1e3, 1.0 / dt))
