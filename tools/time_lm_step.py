"""Times one pnol_lm_step at the cfg5 shape (kernel scopes inside the library). Tuning helper."""
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
dx = np.full(n, float(os.environ.get("PROF_DX", 1e-7)))      # 1e-5: the five-operation FD quotient (exact_div.cuh)
ctx.residual_eval(f, pr["x0"], F=Fd, n=n)
for _ in range(2):
    ctx.lm_step(f, pr["x0"], dx, n, Jd, Fd, Ft, 1e-3, JTJd)
ctx.timer_enable(True)
ctx.timer_reset()
for _ in range(5):
    ctx.lm_step(f, pr["x0"], dx, n, Jd, Fd, Ft, 1e-3, JTJd)
ctx.sync()
out = []
for k in ("fd_jacobian", "syrk", "spd_solve", "residual"):
    ms, c = ctx.timer_get(k)
    out.append("%s %.4f" % (k, ms / max(c, 1)))
print("variant=%s m=%d n=%d  " % (os.environ.get("VARIANT", "default"), m, n) + "  ".join(out))
