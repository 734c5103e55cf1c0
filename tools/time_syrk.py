"""Times pnol_lm_normal_eq (SYRK) at the cfg5 shape; PNOL_LIB overrides the library file (A/B runs of kernel variants)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi  # noqa: E402

if os.environ.get("PNOL_LIB"):
    capi.LIB_PATH = os.path.join(ROOT, "parallelnonlinearoptimizationlibrary_b200", "lib", os.environ["PNOL_LIB"])
m, n = int(os.environ.get("PROF_M", 4_000_000)), int(os.environ.get("PROF_N", 256))
ctx = capi.Context(0)
rng = np.random.default_rng(0)
blk = rng.normal(size=(4096, n))
Jd, Fd = ctx.malloc(m * n * 8), ctx.malloc(m * 8)
for i in range(0, m, 4096):
    k = min(4096, m - i)
    ctx.memcpy(Jd + i * n * 8, blk[:k], k * n * 8)
ctx.memcpy(Fd, np.ascontiguousarray(rng.normal(size=m)), m * 8)
A, rhs = ctx.malloc(n * n * 8), ctx.malloc(n * 8)
for _ in range(3):
    ctx.lm_normal_eq(Jd, Fd, m, n, 1e-3, A=A, rhs=rhs)
ctx.timer_enable(True)
ctx.timer_reset()
for _ in range(10):
    ctx.lm_normal_eq(Jd, Fd, m, n, 1e-3, A=A, rhs=rhs)
ms, cnt = ctx.timer_get("syrk")
fl = float(m) * n * (n + 1) + 2.0 * m * n
print("%s m=%d n=%d syrk %.4f ms  %.2f TFLOP/s" % (os.environ.get("PNOL_LIB", "default"), m, n, ms / cnt, fl / (ms / cnt * 1e-3) / 1e12))
