"""Where a bounded-BFGS iteration at n = 4096 (cfg3) spends its time: library kernel scopes vs wall clock."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi, hostapi  # noqa: E402

n = int(os.environ.get("BFGS_N", 4096))
ctx = capi.Context(0)
hostapi.attach(ctx)
SW = [1e-4, 0.8, 1e-6, 1.0, 1e-10, 2.0, 50, 1e-5, 1e-6, 1e-3, 20, 1e-5, 1e-5, 0]
x0 = np.full(n, 2.0)
x0[0] = -5.0
lb, ub = np.full(n, -5.0), np.full(n, 5.0)
hostapi.bfgs("bfgs_bnd_sw", "rosenbrock", x0, SW[:10] + [2] + SW[11:], lb, ub, pool_width=8)
ctx.timer_enable(True)
ctx.timer_reset()
l0 = ctx.launches()
t0 = time.perf_counter()
r = hostapi.bfgs("bfgs_bnd_sw", "rosenbrock", x0, SW, lb, ub, pool_width=8)
ctx.sync()
dt = time.perf_counter() - t0
print("n=%d: %d iterations, %.3f ms wall per iteration, %d kernel launches per iteration" % (n, r["iterations"], dt / r["iterations"] * 1e3,
                                                                                          (ctx.launches() - l0) / r["iterations"]))
tot = 0
for name in ("fd_points", "alpha_pool", "matvec_neg", "hinv_rank2", "hinv_literal", "eval_batch", "fd_hessian"):
    ms, cnt = ctx.timer_get(name)
    if cnt:
        print("  %-14s %8.4f ms avg x %d = %.3f ms per iteration" % (name, ms / cnt, cnt, ms / r["iterations"]))
        tot += ms
print("  kernel scopes total %.3f ms per iteration" % (tot / r["iterations"]))
