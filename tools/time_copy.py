"""Host <-> device copy bandwidth of pnol_memcpy with pageable numpy arrays (the path LevMarq::findMin's host vectors take)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi  # noqa: E402

ctx = capi.Context(0)
n = 4_000_000
a = np.random.default_rng(0).normal(size=n)
out = np.empty(n)
p = ctx.malloc(n * 8)
for _ in range(2):
    ctx.memcpy(p, a, n * 8)
    ctx.memcpy(out, p, n * 8)
t0 = time.perf_counter()
for _ in range(10):
    ctx.memcpy(p, a, n * 8)
t1 = time.perf_counter()
for _ in range(10):
    ctx.memcpy(out, p, n * 8)
t2 = time.perf_counter()
assert np.array_equal(a, out)
print("32 MB pageable: H2D %.2f ms (%.1f GB/s)  D2H %.2f ms (%.1f GB/s)" % ((t1 - t0) * 100, n * 8 * 10 / (t1 - t0) / 1e9, (t2 - t1) * 100, n * 8 * 10 / (t2 - t1) / 1e9))
