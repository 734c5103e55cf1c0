"""Generates tests/golden/bfgsbnd_mpi_golden.npz from the VERBATIM reference's BFGSBnd_MPI::findMinBnd (oracle/_ref/pnol_ref_cli
bfgsbnd_mpi = /root/reference/Source/BFGS_with_bnd_linsearch_MPI.cpp compiled against oracle/shim; the pool width is the
mini-MPI rank count). Run in the build container (needs /root/reference):

    python tests/golden/make_bfgsbnd_mpi_golden.py

Entries are `<case>/<name>`; inputs are stored beside the outputs. Every case also stores what the reference's own progress
prints say it went through (recursions into the reduced problem, steepest-descent retries, iterations) and, for the
fixed-iteration cases, the "ulp twin": the same run from a start point moved by ONE ulp (see tests/golden/make_golden.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

# parameters of testBFGSBnd_MPI (Source/Examples.cpp:112), maxIter per case
PARAMS = dict(c1=1e-4, c2=0.1, alphamin=1e-16, maxalphamult=4.0, alphaguess=1.0, maxiterls=1000, dxgrad=1e-7, dxhess=1e-3, xmindiff=1e-5,
              mingrad=1e-5, fsteptol=1e-5)


def example_start():
    # testBFGSBnd_MPI (Source/Examples.cpp:97-104): Rosenbrock n = 10, X0 = 3 except X0[0] = -0.5, box [-5, 5] except Xlb[0] = -1
    x0 = np.full(10, 3.0)
    x0[0] = -0.5
    lb = np.full(10, -5.0)
    lb[0] = -1.0
    return x0, lb, np.full(10, 5.0)


def cases():
    x0, lb, ub = example_start()
    out = {}
    for P in (4, 8):
        out["example_P%d_it3" % P] = ("rosenbrock", x0, lb, ub, P, 3, {}, True)        # bound-limited line searches only
        out["example_P%d_it10" % P] = ("rosenbrock", x0, lb, ub, P, 10, {}, True)      # X[0] reaches -1: recursion, then continues
    for P in (2, 4, 8):
        out["example_P%d_it200" % P] = ("rosenbrock", x0, lb, ub, P, 200, {}, False)   # the example itself, run to its own stop
    n = 12
    xb = np.full(n, 2.0)
    xb[0] = -5.0
    out["onbound12_P8_it4"] = ("rosenbrock", xb, np.full(n, -5.0), np.full(n, 5.0), 8, 4, {}, True)          # start ON a bound
    # optimum outside the box: recursion, and the gradient still points out afterwards -> the optimiser gives up there
    for P in (3, 8):
        out["ub05_P%d_it100" % P] = ("rosenbrock", np.zeros(6), np.full(6, -2.0), np.full(6, 0.5), P, 100, {}, False)
    # every variable ends on its bound: nothing left to recurse on
    out["power2_allfrozen_P4"] = ("power:2", np.full(5, 2.0), np.full(5, 0.5), np.full(5, 3.0), 4, 50, {}, False)
    # initial inverse Hessian from the FD Hessian (initHessFD)
    out["power2_inithess_P4"] = ("power:2", np.array([2.0, 1.5, -1.0, 0.7, 2.5]), np.full(5, -3.0), np.full(5, 3.0), 4, 50, dict(inithess=1), False)
    out["rastrigin4_P4_it3"] = ("rastrigin", np.full(4, 2.2), np.full(4, -5.12), np.full(4, 5.12), 4, 3, {}, True)
    return out


def run_reference(obj, x0, lb, ub, P, iters, extra):
    kw = dict(PARAMS)
    kw.update(extra)
    return O.ref_cli("bfgsbnd_mpi", arrays=dict(x=x0, xlb=lb, xub=ub), obj=obj, maxiter=iters, nprocs=P, verbose=1, quiet="0", **kw)


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    G = {}
    for name, (obj, x0, lb, ub, P, iters, extra, twin) in cases().items():
        r = run_reference(obj, x0, lb, ub, P, iters, extra)
        s = r["_stdout"]
        G[name + "/obj"] = np.array(obj)
        G[name + "/x0"] = x0
        G[name + "/lb"] = lb
        G[name + "/ub"] = ub
        G[name + "/P"] = np.array(P)
        G[name + "/iters"] = np.array(iters)
        G[name + "/inithess"] = np.array(int(extra.get("inithess", 0)))
        G[name + "/X"] = r["X"]
        G[name + "/f0"] = r["f0"]
        G[name + "/fOpt"] = r["fOpt"]
        # what the run went through, from the reference's own prints
        G[name + "/iterations_done"] = np.array(s.count("---> At iter"))
        G[name + "/recursions"] = np.array(s.count("reached box boundary"))
        G[name + "/steepest_descent_retries"] = np.array(s.count("Line search failed"))
        if twin:
            x1 = x0.copy()
            idx = 1
            x1[idx] = np.nextafter(x1[idx], x1[idx] + 1.0)
            t = run_reference(obj, x1, lb, ub, P, iters, extra)
            G[name + "/X_ulp"] = t["X"]
            G[name + "/fOpt_ulp"] = t["fOpt"]
        print("%-24s f0 = %-10g fOpt = %-22.17g iters %2d recursions %d sd-retries %d%s" % (
            name, r["f0"][0], r["fOpt"][0], G[name + "/iterations_done"], G[name + "/recursions"], G[name + "/steepest_descent_retries"],
            "  ulp-twin dX = %.2e" % (np.linalg.norm(t["X"] - r["X"]) / np.linalg.norm(r["X"])) if twin else ""))
    np.savez_compressed(os.path.join(HERE, "bfgsbnd_mpi_golden.npz"), **G)


if __name__ == "__main__":
    main()
