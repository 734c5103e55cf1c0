"""Times the GA state machine at the cfg4 shape (Rastrigin, Npop = 1M x 32) with the library's per-stage CUDA-event scopes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi  # noqa: E402

npop = int(os.environ.get("GA_NPOP", 1_000_000))
n = int(os.environ.get("GA_N", 32))
gens = int(os.environ.get("GA_GENS", 5))
ctx = capi.Context(0)
f = ctx.functor(capi.F_RASTRIGIN)
lb, ub = np.full(n, -5.12), np.full(n, 5.12)
ga = ctx.ga_create(f, n, lb, ub, npop, gens + 2, dict(seed=12345, scale=1.0 - 2.0 ** -20), nstatic=1e9)
t0 = time.perf_counter()
f0 = ga.init(np.full(n, 1.0))
ctx.sync()
t1 = time.perf_counter()
print("init %.1f ms, f0 = %.6f" % ((t1 - t0) * 1e3, f0))
ga.generation()
ctx.sync()
ctx.timer_enable(True)
ctx.timer_reset()
t0 = time.perf_counter()
for _ in range(gens):
    ga.generation()
ctx.sync()
t1 = time.perf_counter()
st = ga.status()
print("npop=%d n=%d: %.3f ms/generation wall, f_best %.6f, stream_pos %d, sizes %d/%d/%d/%d" % (
    npop, n, (t1 - t0) * 1e3 / gens, st.f_best, st.stream_pos, st.n_elite, st.n_elite_mut, st.n_cross, st.n_rand))
for name in ("ga_prep", "ga_fitness", "ga_crossover", "ga_mutation", "ga_elite_mutation", "ga_check_identical", "ga_check_bounds", "ga_pop_sort", "eval_batch"):
    ms, cnt = ctx.timer_get(name)
    if cnt:
        print("  %-20s %8.3f ms avg (%d)" % (name, ms / cnt, cnt))
