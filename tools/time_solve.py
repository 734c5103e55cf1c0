"""Times pnol_spd_solve for a few n (CUDA-event scope inside the library). Tuning helper."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi  # noqa: E402

ctx = capi.Context(0)
for n in (32, 64, 128, 256, 512):
    rng = np.random.default_rng(n)
    M = rng.normal(size=(n + 20, n))
    A = ctx.to_device(M.T @ M + 0.1 * np.eye(n))
    b = ctx.to_device(rng.normal(size=n))
    xd = ctx.malloc(n * 8)
    for _ in range(3):
        ctx.spd_solve(A, b, n, x=xd)
    ctx.timer_enable(True)
    ctx.timer_reset()
    for _ in range(20):
        ctx.spd_solve(A, b, n, x=xd)
    ms, cnt = ctx.timer_get("spd_solve")
    ctx.timer_enable(False)
    print("n=%4d spd_solve %.1f us" % (n, ms / cnt * 1e3))
