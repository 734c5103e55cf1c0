"""Generates tests/golden/cfg3_golden.npz from the VERBATIM reference (oracle/_ref/pnol_ref_cli): BASELINE.json config 3 at its FULL
size -- BFGS_Bnd_MPI_SW on Rosenbrock, n = 4096, box [-5, 5]^n, start on a bound, pool width 8 (8 mini-MPI ranks), parameters of
Source/Examples.cpp:37 -- after 2 iterations (each updateHessianInv of the reference is two 4096^3 products through
vector<vector<double>>: about 5 minutes per iteration on this container, which is why the 20-iteration runs of tools/run_configs.py
have no reference beside them). Stored beside it: the reference's own result when one start coordinate moves by one ulp.

    python tests/golden/make_cfg3_golden.py        # build container (needs oracle/_ref); about 25 minutes
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as O  # noqa: E402

N, ITERS, POOL = 4096, 2, 8
SW = dict(c1=1e-4, c2=0.8, dalpha=1e-6, alphaguess=1.0, alphatol=1e-10, alphamult=2.0, maxiterls=50, bndtol=1e-5, dxgrad=1e-6, dxhess=1e-3,
          xmindiff=1e-5, mingrad=1e-5)


def inputs():
    x0 = np.full(N, 2.0)
    x0[0] = -5.0                       # on the lower bound (Source/Examples.cpp:21-25 style)
    return x0, np.full(N, -5.0), np.full(N, 5.0)


def main():
    assert O.have_ref()
    x0, lb, ub = inputs()
    G = {"n": N, "iters": ITERS, "pool": POOL}
    t0 = time.perf_counter()
    r = O.ref_cli("bfgs_bnd_sw", arrays=dict(x=x0, xlb=lb, xub=ub), obj="rosenbrock", maxiter=ITERS, nprocs=POOL, timeout=7200, **SW)
    print("reference: %.0f s, f0 %.17g, fOpt %.17g" % (time.perf_counter() - t0, r["f0"][0], r["fOpt"][0]), flush=True)
    G["X"], G["f0"], G["fOpt"] = r["X"], r["f0"], r["fOpt"]
    xt = x0.copy()
    xt[1] = np.nextafter(xt[1], 3.0)
    r1 = O.ref_cli("bfgs_bnd_sw", arrays=dict(x=xt, xlb=lb, xub=ub), obj="rosenbrock", maxiter=ITERS, nprocs=POOL, timeout=7200, **SW)
    G["X_ulp"] = r1["X"]
    print("one-ulp twin moves X by %.3g (relative)" % (np.linalg.norm(r1["X"] - r["X"]) / np.linalg.norm(r["X"])), flush=True)
    np.savez_compressed(os.path.join(HERE, "cfg3_golden.npz"), **G)
    print("wrote cfg3_golden.npz")


if __name__ == "__main__":
    main()
