"""Generates tests/golden/bench_fixture.npz: what bench.py's `parity` block compares against at every GPU count.

    python tests/golden/make_bench_fixture.py                 # CPU part (oracle): FD gradient of Rosenbrock at n = 4096
    python tests/golden/make_bench_fixture.py --ga            # on a B200 (gpurun): adds the GA fingerprint of the 1-GPU run

  fdgrad4096   Objective::gradientApproximationMPI (Source/PNOL_Objective.cpp:88-159) at cfg3's size from the ORACLE: bench.py's
               column-split gradient must reproduce it bit for bit at 1 / 2 / 4 / 8 GPUs.
  ga           SHA-256 of the sorted objective values and the stream position after bench.py's GA run (Rastrigin, 1M x 32, seed 12345,
               1 + 10 generations) on ONE GPU. The oracle cannot run this size (O(Npop^2) sort and duplicate scan: hours per
               generation), so the fingerprint is the 1-GPU result -- itself bit-exact against the oracle at every size the oracle
               finishes (tests/test_gpu_ga.py) -- and the 2 / 4 / 8-GPU runs must match it.
  (the LM fixture of the parity block is tests/golden/baseline_lm_golden.npz, case cfg5: the threaded oracle restatement at full size)"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "bench_fixture.npz")

GA_SHAPE = dict(npop=1_000_000, n=32, seed=12345, scale=1.0 - 2.0 ** -20, x0=2.5, box=5.12, warm=1, gens=10)


def fdgrad_inputs():
    rng = np.random.default_rng(4096)
    return rng.uniform(-2, 2, 4096), np.full(4096, 1e-6)


def ga_fingerprint(F, stream_pos):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(F, dtype=np.float64).tobytes())
    h.update(np.uint64(stream_pos).tobytes())
    return h.hexdigest()


def main():
    G = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    if "--ga" in sys.argv:
        from parallelnonlinearoptimizationlibrary_b200 import capi
        ctx = capi.Context(0)
        s = GA_SHAPE
        fr = ctx.functor(capi.F_RASTRIGIN)
        ga = ctx.ga_create(fr, s["n"], np.full(s["n"], -s["box"]), np.full(s["n"], s["box"]), s["npop"], s["gens"] + 2,
                           dict(seed=s["seed"], scale=s["scale"]), nstatic=1e9)
        ga.init(np.full(s["n"], s["x0"]))
        for _ in range(s["warm"] + s["gens"]):
            ga.generation()
        _, F = ga.population()
        st = ga.status()
        G["ga/sha256"] = np.array(ga_fingerprint(F, st.stream_pos))
        G["ga/f_best"] = np.array(st.f_best)
        G["ga/stream_pos"] = np.array(st.stream_pos, dtype=np.uint64)
        print("GA fingerprint", G["ga/sha256"], "f_best", st.f_best, "stream_pos", st.stream_pos)
    else:
        import oracle_lib as O
        x, dx = fdgrad_inputs()
        g, f0 = O.fd_gradient(O.OFunctor(1), x, dx)
        G["fdgrad4096/g"] = g
        G["fdgrad4096/f0"] = np.array(f0)
        print("fdgrad4096: f0 =", f0)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, sorted(G))


if __name__ == "__main__":
    main()
