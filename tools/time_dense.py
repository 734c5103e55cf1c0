"""Times the streaming BFGS kernels (p = -D g, rank-2 updateHessianInv) at n = PROF_N (default 4096, cfg3) and the LM step's damped
solve at n = 256, with a result check against numpy; PNOL_LIB overrides the library file (A/B runs of kernel variants)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi  # noqa: E402

if os.environ.get("PNOL_LIB"):
    capi.LIB_PATH = os.path.join(ROOT, "parallelnonlinearoptimizationlibrary_b200", "lib", os.environ["PNOL_LIB"])
HBM = 6549.4
ctx = capi.Context(0)


def timed(name, fn, reps=30, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.timer_enable(True)
    ctx.timer_reset()
    for _ in range(reps):
        fn()
    ms, cnt = ctx.timer_get(name)
    ctx.timer_enable(False)
    return ms / max(cnt, 1)


for n in [int(v) for v in os.environ.get("PROF_N", "4096").split(",")]:
    rng = np.random.default_rng(9)
    D = np.diag(rng.uniform(0.5, 2.0, n))
    u3 = rng.normal(size=(n, 3)) / np.sqrt(n)
    D = D + u3 @ u3.T
    g = rng.normal(size=n)
    s = 0.1 * g + 0.05 * rng.normal(size=n)
    Dd, gd, sd, pd = ctx.to_device(D), ctx.to_device(g), ctx.to_device(s), ctx.malloc(n * 8)
    p = ctx.matvec_neg(D, g, n)
    want = -(D @ g)
    e_mv = np.linalg.norm(p - want) / np.linalg.norm(want)
    D1 = ctx.bfgs_update_hinv(D.copy(), g, s, n, mode=capi.HINV_RANK2)
    rho = 1.0 / (g @ s)
    I = np.eye(n)
    Dw = (I - rho * np.outer(s, g)) @ D @ (I - rho * np.outer(g, s)) + rho * np.outer(s, s)
    e_h = np.linalg.norm(D1 - Dw) / np.linalg.norm(Dw)
    ms = timed("matvec_neg", lambda: ctx.matvec_neg(Dd, gd, n, p=pd))
    by = n * n * 8.0
    print("n=%d matvec_neg  %.4f ms  %.0f GB/s  frac %.3f  rel err %.2e" % (n, ms, by / ms / 1e6, by / ms / 1e6 / HBM, e_mv))
    ms = timed("hinv_rank2", lambda: ctx.bfgs_update_hinv(Dd, gd, sd, n, mode=capi.HINV_RANK2))
    by = 3.0 * n * n * 8.0
    print("n=%d hinv_rank2  %.4f ms  %.0f GB/s  frac %.3f  rel err %.2e" % (n, ms, by / ms / 1e6, by / ms / 1e6 / HBM, e_h))
    # the BFGS loop's order: update, then the next search direction (three passes over D per iteration, back to back)
    def pair():
        ctx.bfgs_update_hinv(Dd, gd, sd, n, mode=capi.HINV_RANK2)
        ctx.matvec_neg(Dd, gd, n, p=pd)
    for _ in range(3):
        pair()
    ctx.sync()
    ctx.timer_enable(True)
    ctx.timer_reset()
    for _ in range(30):
        pair()
    m1, c1 = ctx.timer_get("hinv_rank2")
    m2, c2 = ctx.timer_get("matvec_neg")
    ctx.timer_enable(False)
    print("n=%d update + direction interleaved: hinv_rank2 %.4f ms, matvec_neg %.4f ms  (PNOL_DENSE_SERPENTINE=%s)"
          % (n, m1 / max(c1, 1), m2 / max(c2, 1), os.environ.get("PNOL_DENSE_SERPENTINE", "default")))
    for q in (Dd, gd, sd, pd):
        ctx.free(q)

ns = 256
rng = np.random.default_rng(3)
M = rng.normal(size=(ns, ns))
A = M @ M.T / ns + np.eye(ns)
b = rng.normal(size=ns)
x = ctx.spd_solve(A, b, ns)
Ad, bd, xd = ctx.to_device(A), ctx.to_device(b), ctx.malloc(ns * 8)
ms = timed("spd_solve", lambda: ctx.spd_solve(Ad, bd, ns, x=xd))
print("n=%d spd_solve %.4f ms  rel err %.2e" % (ns, ms, np.linalg.norm(x - np.linalg.solve(A, b)) / np.linalg.norm(x)))
