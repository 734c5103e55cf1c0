"""Times the FD Jacobian + residual kernels at the cfg5 shape (CUDA-event scopes inside the library). Tuning helper."""
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
Jd, Fd = ctx.malloc(m * n * 8), ctx.malloc(m * 8)
xd, dxd = ctx.to_device(pr["x0"]), ctx.to_device(np.full(n, float(os.environ.get("PROF_DX", 1e-7))))
for _ in range(3):
    ctx.fd_jacobian(f, xd, dxd, J=Jd, F=Fd, n=n)
    ctx.residual_eval(f, xd, F=Fd, n=n)
ctx.timer_enable(True)
ctx.timer_reset()
for _ in range(10):
    ctx.fd_jacobian(f, xd, dxd, J=Jd, F=Fd, n=n)
    ctx.residual_eval(f, xd, F=Fd, n=n)
ctx.sync()
j, cj = ctx.timer_get("fd_jacobian")
r, cr = ctx.timer_get("residual")
by = m * n * 8 + 3 * m * 8
print("variant=%s m=%d n=%d fd_jacobian %.4f ms (%.0f GB/s, %.1f%% of 6549.4) residual %.4f ms" % (
    os.environ.get("PNOL_LORENTZ_BLOCKS", "default"), m, n, j / cj, by / (j / cj * 1e-3) / 1e9, by / (j / cj * 1e-3) / 1e9 / 65.494, r / cr))
