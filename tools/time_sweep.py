"""Times the 1M x 32 Rastrigin fitness sweep (CUDA-event scope inside the library). Tuning helper."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi  # noqa: E402

B, n = int(os.environ.get("SWEEP_B", 1_000_000)), int(os.environ.get("SWEEP_N", 32))
ctx = capi.Context(0)
pts = ctx.to_device(np.random.default_rng(0).uniform(-5.12, 5.12, size=(B, n)))
fo = ctx.malloc(B * 8)
fr = ctx.functor(capi.F_RASTRIGIN)
for _ in range(5):
    ctx.eval_batch(fr, pts, B, n, f_out=fo)
ctx.timer_enable(True)
ctx.timer_reset()
for _ in range(50):
    ctx.eval_batch(fr, pts, B, n, f_out=fo)
ms, cnt = ctx.timer_get("eval_batch")
by = B * (n + 1) * 8
print("sweep B=%d n=%d: %.4f ms  %.3e evals/s  %.0f GB/s (%.1f%% of 6549.4)" % (B, n, ms / cnt, B / (ms / cnt * 1e-3), by / (ms / cnt * 1e-3) / 1e9,
                                                                              by / (ms / cnt * 1e-3) / 1e9 / 65.494))
