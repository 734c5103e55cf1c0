"""Phase timestamps of spd_solve_kernel (library variant built with -DPNOL_SOLVE_STAMPS): where the solve's time goes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi
ctx = capi.Context(0)
n = int(os.environ.get("PROF_N", 256))
rng = np.random.default_rng(3)
M = rng.normal(size=(n, n)); A = M @ M.T / n + np.eye(n); b = rng.normal(size=n)
for _ in range(3):
    x = ctx.spd_solve(A, b, n)
x = ctx.spd_solve(A, b, n)
names = ["start"]
for k in range((n + 31) // 32):
    names += ["dload%d" % k, "dldl%d" % k, "diag%d" % k, "panel%d" % k, "sync%d" % k, "update%d" % k, "sync%d'" % k]
names += ["backsub"]
prev = 0.0
tot = {}
for nm, t in zip(names, x[:len(names)]):
    print("%-10s %8.0f ns  (+%6.0f)" % (nm, t, t - prev))
    key = nm.rstrip("0123456789'")
    tot[key] = tot.get(key, 0) + t - prev
    prev = t
print(tot)
