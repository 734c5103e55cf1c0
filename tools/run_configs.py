"""Runs the five BASELINE.json configurations end to end through the plugin classes on one B200 and, where the CPU can finish
in seconds, the verbatim reference beside them (oracle/_ref). Writes one JSON (default gpurun_out/configs.json).

cfg1  BFGS, Rosenbrock n=10 (the reference's own CPU-runnable case)          -- parity anchor, reference at full size
cfg2  LevMarq, Lorentzian fit m=100k x n=16                                    -- reference at full size
cfg3  BFGS_Bnd_MPI_SW, Rosenbrock n=4096 in a box, pool width 8, 20 iterations -- reference at n=512 (n^3 update: 260 s each at 4096)
cfg4  GeneticAlgorithmMPI, Rastrigin 1M x 32, 200 generations                  -- reference sweep only (O(Npop^2) sort: hours per generation)
cfg5  LevMarqMPI, m=4M x n=256                                                 -- bench.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parallelnonlinearoptimizationlibrary_b200 import capi, hostapi, problems  # noqa: E402
import oracle_lib as O  # noqa: E402

out = {}
ctx = capi.Context(0)
hostapi.attach(ctx)
have_ref = O.have_ref()


def timed(fn, reps=1):
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    ctx.sync()
    return r, (time.perf_counter() - t0) / reps


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


# ---- cfg1 ----
BF = [1e-4, 0.9, 1e-6, 1.0, 1000, 1e-7, 1e-3, 100, 1e-5, 1e-5, 0]
hostapi.bfgs("bfgs", "rosenbrock", np.full(10, 3.0), BF)
r, dt = timed(lambda: hostapi.bfgs("bfgs", "rosenbrock", np.full(10, 3.0), BF))
c = {"gpu_seconds": dt, "iterations": r["iterations"], "fOpt": r["fOpt"], "x_err_vs_ones": float(np.max(np.abs(r["X"] - 1)))}
if have_ref:
    t0 = time.perf_counter()
    rr = O.ref_cli("bfgs", arrays=dict(x=np.full(10, 3.0)), obj="rosenbrock", maxiter=100, c1=1e-4, c2=0.9, dalpha=1e-6, alphaguess=1.0, maxiterls=1000,
                   dxgrad=1e-7, dxhess=1e-3, xmindiff=1e-5, mingrad=1e-5)
    c.update(ref_seconds_incl_process_start=time.perf_counter() - t0, ref_fOpt=float(rr["fOpt"][0]), rel_x_vs_ref=rel(r["X"], rr["X"]))
out["cfg1_bfgs_rosenbrock_n10"] = c

# ---- cfg2 ----
pr = problems.lorentz_problem(100_000, 8)
prob = hostapi.LMProblem(pr["t"], pr["y"], pr["w"])
prob.run(pr["x0"], 0.001, 10.0, 1e-6, 3, 0.0)
iters = 20
r, dt = timed(lambda: prob.run(pr["x0"], 0.001, 10.0, 1e-6, iters, 0.0))
c = {"m": 100_000, "n": 16, "iterations": r["iterations"], "gpu_seconds": dt, "gpu_iters_per_s": r["iterations"] / dt, "chiSq": r["chiSq"],
     "rel_x_vs_truth": rel(r["X"], pr["x_true"])}
if have_ref:
    t0 = time.perf_counter()
    rr = O.ref_cli("bench_lm", arrays=dict(x=pr["x0"], t=pr["t"], y=pr["y"]), obj="lorentz", w=pr["w"], steps=iters, warmup=0, dxgrad=1e-6, lambda0=0.001,
                   factor=10.0, nprocs=1)
    line = json.loads([ln for ln in rr["_stdout"].splitlines() if ln.startswith("{")][-1])
    c.update(ref_s_per_iter=line["s_per_iter"], ref_iters_per_s=1.0 / line["s_per_iter"], rel_x_vs_ref=rel(r["X"], rr["X"]))
prob.close()
out["cfg2_lm_m100k_n16"] = c

# ---- cfg3 ----
SW = [1e-4, 0.8, 1e-6, 1.0, 1e-10, 2.0, 50, 1e-5, 1e-6, 1e-3, 20, 1e-5, 1e-5, 0]
for n in (512, 4096):
    x0 = np.full(n, 2.0)
    x0[0] = -5.0
    lb, ub = np.full(n, -5.0), np.full(n, 5.0)
    c = {"n": n, "pool_width": 8, "max_iterations": 20}
    for mode, name in ((1, "rank2"), (0, "literal_dmma")):
        hostapi.set_hinv_mode(mode)
        hostapi.bfgs("bfgs_bnd_sw", "rosenbrock", x0, SW[:10] + [2] + SW[11:], lb, ub, pool_width=8)
        r, dt = timed(lambda: hostapi.bfgs("bfgs_bnd_sw", "rosenbrock", x0, SW, lb, ub, pool_width=8))
        c[name] = {"gpu_seconds": dt, "iterations": r["iterations"], "s_per_iteration": dt / max(r["iterations"], 1), "fOpt": r["fOpt"], "f0": r["f0"]}
        c[name]["_X"] = r["X"]
    hostapi.set_hinv_mode(1)
    c["rel_x_rank2_vs_literal"] = rel(c["rank2"].pop("_X"), c["literal_dmma"]["_X"])
    Xlit = c["literal_dmma"].pop("_X")
    if have_ref and n == 512:
        t0 = time.perf_counter()
        rr = O.ref_cli("bfgs_bnd_sw", arrays=dict(x=x0, xlb=lb, xub=ub), obj="rosenbrock", maxiter=20, nprocs=8, c1=1e-4, c2=0.8, dalpha=1e-6, alphaguess=1.0,
                       alphatol=1e-10, alphamult=2.0, maxiterls=50, bndtol=1e-5, dxgrad=1e-6, dxhess=1e-3, xmindiff=1e-5, mingrad=1e-5, timeout=3000)
        c.update(ref_seconds_8_ranks=time.perf_counter() - t0, ref_fOpt=float(rr["fOpt"][0]), rel_x_literal_vs_ref=rel(Xlit, rr["X"]))
    out["cfg3_bfgs_bnd_sw_n%d" % n] = c

# ---- cfg4 ----
npop, n, gens = 1_000_000, 32, 200
f = ctx.functor(capi.F_RASTRIGIN)
lb, ub = np.full(n, -5.12), np.full(n, 5.12)
ga = ctx.ga_create(f, n, lb, ub, npop, gens, dict(seed=12345, scale=1.0 - 2.0 ** -20), nstatic=1e9)
(_, dt_init) = timed(lambda: ga.init(np.full(n, 2.5)))     # f(x0) = 840: away from the integer-lattice local minima
f_first = ga.status().f_best
t0 = time.perf_counter()
for _ in range(gens):
    ga.generation()
ctx.sync()
dt = time.perf_counter() - t0
st = ga.status()
c = {"npop": npop, "n": n, "generations": st.generation, "gpu_seconds": dt, "ms_per_generation": dt / gens * 1e3, "init_ms": dt_init * 1e3,
     "evaluations": (npop - st.n_elite) * gens + npop, "evals_per_s_whole_generation": ((npop - st.n_elite) * gens) / dt,
     "f_best_start": f_first, "f_best_end": st.f_best, "stream_draws": int(st.stream_pos)}
if have_ref:
    rr = O.ref_cli("bench_ga_eval", obj="rastrigin", n=32, npop=100_000, reps=3, nprocs=1)
    line = json.loads([ln for ln in rr["_stdout"].splitlines() if ln.startswith("{")][-1])
    c["ref_sweep_evals_per_s_1_rank"] = line["evals_per_s"]
ga.close()
out["cfg4_ga_rastrigin_1M_x32_200gen"] = c

hostapi.detach()
ctx.close()
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "configs.json")
json.dump(out, open(dst, "w"), indent=1, default=float)
print(json.dumps(out, indent=1, default=float))
