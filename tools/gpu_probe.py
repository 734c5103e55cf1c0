"""Kernel-by-kernel timing probe on one B200 (CUDA events through the library's timers). Development tool:
prints one JSON object; bench.py is the contract benchmark."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelnonlinearoptimizationlibrary_b200 import capi, problems  # noqa: E402


def timed(ctx, name, fn, reps=5):
    fn()
    ctx.sync()
    ctx.timer_reset()
    for _ in range(reps):
        fn()
    ms, cnt = ctx.timer_get(name)
    return ms / max(cnt, 1)


def main():
    m = int(os.environ.get("PROBE_M", 4_000_000))
    K = int(os.environ.get("PROBE_K", 128))
    out = {}
    ctx = capi.Context(0)
    ctx.timer_enable(True)
    out["sm_count"] = ctx.sm_count
    out["dmma_peak_tflops"] = ctx.measure_dmma_peak()
    out["copy_gbs"] = ctx.measure_copy_bandwidth()

    pr = problems.lorentz_problem(m, K)
    n = pr["n"]
    f = ctx.functor(capi.F_LORENTZ_SUM, (pr["w"],), (), (pr["t"], pr["y"]), m)
    Jd = ctx.malloc(m * n * 8)
    Fd = ctx.malloc(m * 8)
    xd = ctx.to_device(pr["x0"])
    dxd = ctx.to_device(np.full(n, 1e-7))
    packed_A = ctx.malloc(n * n * 8)
    JTJ = ctx.malloc(n * n * 8)
    rhs = ctx.malloc(n * 8)
    sig = ctx.malloc(n * 8)

    t = timed(ctx, "fd_jacobian", lambda: ctx.fd_jacobian(f, xd, dxd, J=Jd, F=Fd, n=n))
    out["jacobian_ms"] = t
    out["jacobian_gbs"] = (m * n * 8 + m * 24) / (t * 1e-3) / 1e9
    t = timed(ctx, "residual", lambda: ctx.residual_eval(f, xd, F=Fd, n=n))
    out["residual_ms"] = t
    ctx.timer_reset()
    for _ in range(3):
        ctx.lm_normal_eq(Jd, Fd, m, n, 1e-3, JTJ=JTJ, A=packed_A, rhs=rhs)
    ms, cnt = ctx.timer_get("syrk")
    out["syrk_ms"] = ms / cnt
    out["syrk_tflops_lower"] = m * n * (n + 1) / (ms / cnt * 1e-3) / 1e12
    out["syrk_tflops_full"] = 2.0 * m * n * n / (ms / cnt * 1e-3) / 1e12
    ms, cnt = ctx.timer_get("syrk_finish")
    out["syrk_finish_ms"] = ms / cnt
    t = timed(ctx, "spd_solve", lambda: ctx.spd_solve(packed_A, rhs, n, x=sig))
    out["spd_solve_ms"] = t
    for p in (Jd, Fd, packed_A, JTJ, rhs, sig):
        ctx.free(p)

    # GA fitness sweep
    B, nd = 1_000_000, 32
    rng = np.random.default_rng(0)
    pts = ctx.to_device(rng.uniform(-5.12, 5.12, size=(B, nd)))
    fo = ctx.malloc(B * 8)
    fr = ctx.functor(capi.F_RASTRIGIN)
    t = timed(ctx, "eval_batch", lambda: ctx.eval_batch(fr, pts, B, nd, f_out=fo), reps=10)
    out["sweep_ms"] = t
    out["sweep_evals_per_s"] = B / (t * 1e-3)
    out["sweep_gbs"] = (B * nd * 8 + B * 8) / (t * 1e-3) / 1e9
    ctx.free(pts); ctx.free(fo)

    # BFGS dense pieces at n = 4096
    nb = 4096
    D = ctx.to_device(np.eye(nb) + 0.001 * rng.normal(size=(nb, nb)))
    g = rng.normal(size=nb)
    s = 0.1 * g + 0.01 * rng.normal(size=nb)
    gd, sd, pd = ctx.to_device(g), ctx.to_device(s), ctx.malloc(nb * 8)
    out["matvec_ms"] = timed(ctx, "matvec_neg", lambda: ctx.matvec_neg(D, gd, nb, p=pd))
    out["hinv_rank2_ms"] = timed(ctx, "hinv_rank2", lambda: ctx.bfgs_update_hinv(D, gd, sd, nb, capi.HINV_RANK2))
    t = timed(ctx, "hinv_literal", lambda: ctx.bfgs_update_hinv(D, gd, sd, nb, capi.HINV_LITERAL), reps=2)
    out["hinv_literal_ms"] = t
    ms, cnt = ctx.timer_get("dgemm_nn")
    out["dgemm_4096_ms"] = ms / max(cnt, 1)
    out["dgemm_tflops"] = 2.0 * nb ** 3 / (ms / max(cnt, 1) * 1e-3) / 1e12
    xr = ctx.to_device(np.full(nb, 2.0))
    dxr = ctx.to_device(np.full(nb, 1e-6))
    fro = ctx.functor(capi.F_ROSENBROCK)
    t0 = time.time()
    for _ in range(5):
        ctx.fd_gradient(fro, np.full(nb, 2.0), np.full(nb, 1e-6))
    out["fd_gradient_4096_wall_ms"] = (time.time() - t0) / 5 * 1e3
    ms, cnt = ctx.timer_get("fd_points")
    out["fd_points_4096_ms"] = ms / max(cnt, 1)
    out["launches"] = ctx.launches()
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
