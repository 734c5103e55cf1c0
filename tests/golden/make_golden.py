"""Generates tests/golden/ref_golden.npz from the VERBATIM reference (oracle/_ref/pnol_ref_cli = /root/reference/Source
compiled against oracle/shim by `make -C oracle ref`). Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The reference holds no golden vectors of its own, so these outputs of the reference code itself are
the pinned answers: the CPU tests check the oracle restatement against them, the GPU tests check the CUDA path.
Every entry is `<case>/<name>`; inputs are stored beside outputs so the tests need nothing else."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as O  # noqa: E402
from parallelnonlinearoptimizationlibrary_b200 import problems  # noqa: E402

G = {}


def put(case, **arrays):
    for k, v in arrays.items():
        G["%s/%s" % (case, k)] = np.asarray(v)


def ulp_twin(cmd, x0, arrays, idx=0, **kw):
    """The same reference run from a start point one entry of which is moved by ONE ulp: how far the reference's own
    iterates move under the smallest possible input change (FD gradients with h ~ 1e-7 amplify rounding by ~1/h, and a line
    search compounds it). Stored beside every optimiser fixture; it is the conditioning the 1e-9 bar has to be read with."""
    x1 = np.array(x0, dtype=np.float64).copy()
    x1[idx] = np.nextafter(x1[idx], x1[idx] + 1.0)
    r = O.ref_cli(cmd, arrays=dict(arrays, x=x1), **kw)
    return dict(X_ulp=r["X"], fOpt_ulp=r["fOpt"])


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    # ---- FD gradient / Hessian (Source/PNOL_Objective.cpp:12-34, 38-85, 88-159) ----
    for spec, n in (("rosenbrock", 5), ("rosenbrock", 40), ("booth", 2), ("goldstein", 2), ("powerprod:3", 5), ("rastrigin", 12)):
        rng = np.random.default_rng(100 + n)
        x, dx = rng.uniform(-2, 2, n), np.full(n, 1e-6)
        r = O.ref_cli("fdgrad", arrays=dict(x=x, dx=dx), obj=spec, nprocs=3)
        case = "fdgrad_%s_%d" % (spec.replace(":", ""), n)
        put(case, x=x, dx=dx, g=r["g"], g_mpi=r["g_mpi"], f=r["f"])
        if n <= 12:
            dxh = np.full(n, 1e-3)
            rh = O.ref_cli("hessian", arrays=dict(x=x, dx=dxh), obj=spec)
            put(case, dxh=dxh, B=rh["B"].reshape(n, n))
    # testGradientEvaluation (Source/Examples.cpp:512-540): PowerObject power 3 at 3*1_5 through the reference's own class
    r = O.ref_cli("fdgrad", arrays=dict(x=np.full(5, 3.0), dx=np.full(5, 1e-6)), obj="power:3", nprocs=4)
    put("testGradientEvaluation", g=r["g"], g_mpi=r["g_mpi"])
    # testGradientApproxMultMPIRecur (Source/Examples.cpp:593-663)
    n = 8
    xfull = 0.1 * np.arange(n)
    ind = np.zeros(n)
    ind[3] = 1
    xr = xfull[ind == 0]
    r = O.ref_cli("recur", arrays=dict(x=xr, dx=np.full(xr.size, 1e-6), constx=xfull, ind=ind), obj="rosenbrock", nprocs=4)
    put("recur_rosenbrock_8", xr=xr, constx=xfull, ind=ind, g=r["g"], g_mpi=r["g_mpi"], f=r["f"])
    # ---- FD Jacobian (Source/PNOL_Objective.cpp:165-197, 202-299) ----
    for K, m in ((4, 300), (32, 200), (128, 96)):
        pr = problems.lorentz_problem(m, K)
        dx = np.full(pr["n"], 1e-7)
        r = O.ref_cli("fdjac", arrays=dict(x=pr["x0"], dx=dx, t=pr["t"], y=pr["y"]), obj="lorentz", w=pr["w"], nprocs=2)
        put("fdjac_lorentz_K%d" % K, t=pr["t"], y=pr["y"], w=pr["w"], x=pr["x0"], dx=dx, J=r["J"].reshape(m, -1), J_mpi=r["J_mpi"].reshape(m, -1),
            F=r["F"])
    r = O.ref_cli("fdjac", arrays=dict(x=np.full(4, 0.1), dx=np.full(4, 1e-6)), obj="cubic")
    put("fdjac_cubic", J=r["J"].reshape(100, 4), F=r["F"])
    # ---- LM (Source/LevenbergMarquardtMPI.cpp:12-173) ----
    for K, m, iters in ((8, 2000, 12), (16, 600, 6)):
        pr = problems.lorentz_problem(m, K)
        r = O.ref_cli("lm", arrays=dict(x=pr["x0"], t=pr["t"], y=pr["y"]), obj="lorentz", w=pr["w"], lambda0=0.001, factor=10.0, dxgrad=1e-7,
                      maxiter=iters, xmindiff=0.0, nprocs=2)
        put("lm_lorentz_K%d" % K, t=pr["t"], y=pr["y"], w=pr["w"], x0=pr["x0"], iters=iters, X=r["X"], F0=r["F0"], F=r["F"])
    r = O.ref_cli("lm", arrays=dict(x=np.array([9.0, 0.5, 0.3])), obj="expcurve_ref", lambda0=0.001, factor=10.0, dxgrad=1e-6, maxiter=100,
                  xmindiff=1e-6)
    put("testLMExpMPI", X=r["X"])
    r = O.ref_cli("lm", arrays=dict(x=np.full(4, 0.1)), obj="cubic", lambda0=0.001, factor=10.0, dxgrad=1e-6, maxiter=100, xmindiff=1e-6)
    put("testLMCubicLinearCoef", X=r["X"], F=r["F"])
    # ---- updateHessianInv (Source/BFGS_with_linesearch.cpp:389-432) ----
    for n in (3, 17, 64):
        rng = np.random.default_rng(n)
        M = rng.normal(size=(n, n))
        D = M @ M.T / n + np.eye(n)
        g = rng.normal(size=n)
        s = 0.1 * g + 0.05 * rng.normal(size=n)
        r = O.ref_cli("updhinv", arrays=dict(D=D, g=g, s=s))
        put("updhinv_%d" % n, D=D, g=g, s=s, Dnew=r["D"].reshape(n, n))
    # ---- box helpers (Source/Box_boundary_functions.cpp:11-40; BFGS_with_bnd_linsearch_MPI.cpp:665-708) ----
    rng = np.random.default_rng(0)
    n = 50
    lb, ub = rng.uniform(-3, -1, n), rng.uniform(1, 3, n)
    x = rng.uniform(-4, 4, n)
    p = rng.normal(size=n)
    p[7] = 0.0
    xin = np.clip(x, lb, ub)
    r1 = O.ref_cli("box", arrays=dict(x=xin, xlb=lb, xub=ub, p=p))
    r2 = O.ref_cli("box", arrays=dict(x=x, xlb=lb, xub=ub))
    put("box", x=x, xin=xin, lb=lb, ub=ub, p=p, alphabnd=r1["alphabnd"], Xfixed=r2["X"])
    # ---- BFGS family ----
    # cfg1 of BASELINE.json: BFGS on Rosenbrock n=10, X0 = 3, params of testBFGS (Source/Examples.cpp:247)
    bf = dict(c1=1e-4, c2=0.9, dalpha=1e-6, alphaguess=1.0, maxiterls=1000, dxgrad=1e-7, dxhess=1e-3, xmindiff=1e-5, mingrad=1e-5)
    for iters in (5, 100):
        r = O.ref_cli("bfgs", arrays=dict(x=np.full(10, 3.0)), obj="rosenbrock", maxiter=iters, **bf)
        put("bfgs_cfg1_it%d" % iters, X=r["X"], f0=r["f0"], fOpt=r["fOpt"], **ulp_twin("bfgs", np.full(10, 3.0), {}, obj="rosenbrock", maxiter=iters, **bf))
    r = O.ref_cli("bfgs", arrays=dict(x=np.full(5, 3.0)), obj="rosenbrock", maxiter=100, **bf)     # testBFGS itself (n = 5)
    put("testBFGS", X=r["X"], f0=r["f0"], fOpt=r["fOpt"])
    # BFGS_MPI pool search, Rosenbrock n=10, X0=10 (testBFGS_MPI, Source/Examples.cpp:163-189, params :181), P = 4 ranks
    bm = dict(c1=1e-4, c2=0.9, maxalphamult=4.0, alphaguess=1.0, maxiterls=50, dxgrad=1e-7, dxhess=1e-3, xmindiff=1e-5, mingrad=1e-5)
    for iters in (5, 60):
        r = O.ref_cli("bfgs_mpi", arrays=dict(x=np.full(10, 10.0)), obj="rosenbrock", maxiter=iters, nprocs=4, **bm)
        put("bfgs_mpi_P4_it%d" % iters, X=r["X"], f0=r["f0"], fOpt=r["fOpt"],
            **ulp_twin("bfgs_mpi", np.full(10, 10.0), {}, obj="rosenbrock", maxiter=iters, nprocs=4, **bm))
    # BFGS_Bnd_MPI_SW: testBFGSBndMPISW (Source/Examples.cpp:12-45, params :37) at P = 2 and 8; and a cfg3-style box problem (n = 64)
    sw = dict(c1=1e-4, c2=0.8, dalpha=1e-6, alphaguess=1.0, alphatol=1e-10, alphamult=2.0, maxiterls=50, bndtol=1e-5, dxgrad=1e-6,
              dxhess=1e-3, xmindiff=1e-5, mingrad=1e-5)
    x0 = np.array([-1.0, 2.0, 2.0])
    for P in (2, 8):
        r = O.ref_cli("bfgs_bnd_sw", arrays=dict(x=x0, xlb=np.full(3, -1.0), xub=np.full(3, 3.0)), obj="rosenbrock", maxiter=200, nprocs=P, **sw)
        put("testBFGSBndMPISW_P%d" % P, x0=x0, lb=np.full(3, -1.0), ub=np.full(3, 3.0), X=r["X"], f0=r["f0"], fOpt=r["fOpt"])
    n = 64
    x0 = np.full(n, 2.0)
    x0[0] = -5.0
    for iters in (3, 20):
        r = O.ref_cli("bfgs_bnd_sw", arrays=dict(x=x0, xlb=np.full(n, -5.0), xub=np.full(n, 5.0)), obj="rosenbrock", maxiter=iters, nprocs=8, **sw)
        put("bfgs_bnd_sw_n64_P8_it%d" % iters, x0=x0, X=r["X"], f0=r["f0"], fOpt=r["fOpt"],
            **ulp_twin("bfgs_bnd_sw", x0, dict(xlb=np.full(n, -5.0), xub=np.full(n, 5.0)), idx=1, obj="rosenbrock", maxiter=iters, nprocs=8, **sw))
    # BFGS_Bnd (serial bounded): testBFGSBnd itself (Source/Examples.cpp:49-83: n = 5, X0 = 2, box [-5,5], params :75) run to
    # convergence, after a fixed 4 iterations, and a start ON a bound (n = 12, X0[0] = -5)
    for iters in (4, 200):
        r = O.ref_cli("bfgs_bnd", arrays=dict(x=np.full(5, 2.0), xlb=np.full(5, -5.0), xub=np.full(5, 5.0)), obj="rosenbrock", maxiter=iters, **sw)
        put("testBFGSBnd_it%d" % iters, X=r["X"], f0=r["f0"], fOpt=r["fOpt"],
            **ulp_twin("bfgs_bnd", np.full(5, 2.0), dict(xlb=np.full(5, -5.0), xub=np.full(5, 5.0)), obj="rosenbrock", maxiter=iters, **sw))
    n = 12
    x0 = np.full(n, 2.0)
    x0[0] = -5.0
    r = O.ref_cli("bfgs_bnd", arrays=dict(x=x0, xlb=np.full(n, -5.0), xub=np.full(n, 5.0)), obj="rosenbrock", maxiter=6, **sw)
    put("bfgs_bnd_n12_it6", x0=x0, X=r["X"], f0=r["f0"], fOpt=r["fOpt"],
        **ulp_twin("bfgs_bnd", x0, dict(xlb=np.full(n, -5.0), xub=np.full(n, 5.0)), idx=1, obj="rosenbrock", maxiter=6, **sw))
    # ---- GA (Source/GeneticAlgorithmMPI.cpp:12-276) driven by the counter stream (oracle/shim timeRand) ----
    for spec, n, npop, gens, box in (("powerprod:2", 4, 150, 6, 10.0), ("rastrigin", 6, 200, 5, 5.12), ("rosenbrock", 3, 64, 8, 2.0)):
        lb, ub = np.full(n, -box), np.full(n, box)
        x0 = np.full(n, 0.3 * box)
        scale = 1.0 - 1.0 / npop
        r = O.ref_cli("ga", arrays=dict(x=x0, xlb=lb, xub=ub), obj=spec, npop=npop, maxgen=gens, seed=777, scale=scale, nprocs=2)
        put("ga_%s" % spec.replace(":", ""), x0=x0, lb=lb, ub=ub, npop=npop, gens=gens, seed=777, scale=scale, X=r["X"], f0=r["f0"],
            fOpt=r["fOpt"], stream_pos=r["stream_pos"])
    rng = np.random.default_rng(4)
    npop, n = 300, 5
    lb, ub = np.full(n, -1.0), np.full(n, 1.0)
    X = rng.uniform(-1.4, 1.4, size=(npop, n))
    X[7] = X[100]
    F = np.round(rng.uniform(0.1, 9, npop), 1)
    r1 = O.ref_cli("popsort", arrays=dict(xpop=X, F=F), n=n)
    r2 = O.ref_cli("checkbounds", arrays=dict(xpop=X, xlb=lb, xub=ub), n=n, seed=5, scale=1.0)
    r3 = O.ref_cli("checkidentical", arrays=dict(xpop=X, xlb=lb, xub=ub), n=n, seed=5, scale=1.0)
    put("ga_stages", X=X, F=F, lb=lb, ub=ub, sortedX=r1["xpop"].reshape(npop, n), sortedF=r1["F"], boundsX=r2["xpop"].reshape(npop, n),
        boundsInd=r2["ind"], boundsPos=r2["stream_pos"], identX=r3["xpop"].reshape(npop, n), identInd=r3["ind"], identPos=r3["stream_pos"])
    out = os.path.join(HERE, "ref_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote %s: %d arrays, %.1f KB" % (out, len(G), os.path.getsize(out) / 1024))


if __name__ == "__main__":
    main()
