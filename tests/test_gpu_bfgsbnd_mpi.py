"""GPU: BFGSBnd_MPI (include/pnol/BFGS_with_bnd_linesearch_MPI.hpp) against the committed outputs of the
verbatim reference's BFGSBnd_MPI::findMinBnd (tests/golden/bfgsbnd_mpi_golden.npz, made by tests/golden/make_bfgsbnd_mpi_golden.py
from Source/BFGS_with_bnd_linsearch_MPI.cpp). The host runs the reference's control flow (pooled secant line search, steepest-descent
retry, one-level active-set recursion); FD gradients, alpha pools, p = -D g and updateHessianInv are kernels. Objective values and
gradients are bit-identical to the reference's; the dense algebra differs in the last bits, so fixed-iteration
iterates are held to max(1e-9, 10 x the reference's own one-ulp sensitivity) and runs to convergence to the optimiser's own stop
tolerance, as for the other BFGS variants (tests/test_gpu_host_api.py)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_bfgsbnd_mpi_golden import PARAMS, cases  # noqa: E402  (the inputs; needs neither the reference nor oracle/_ref)

G = np.load(os.path.join(HERE, "golden", "bfgsbnd_mpi_golden.npz"))
CASES = cases()
RTOL = 1e-9


@pytest.fixture(scope="module")
def host(ctx):
    from parallelnonlinearoptimizationlibrary_b200 import hostapi
    hostapi.attach(ctx)
    yield hostapi


def rel(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), floor, 1e-300))


def run(host, name):
    obj, x0, lb, ub, P, iters, extra, _ = CASES[name]
    k = PARAMS
    # setParams order of Source/BFGS_with_bnd_linesearch_MPI.hpp:78-79
    p = [k["c1"], k["c2"], k["alphamin"], k["maxalphamult"], k["alphaguess"], k["maxiterls"], k["dxgrad"], k["dxhess"], iters, k["xmindiff"],
         k["mingrad"], k["fsteptol"], extra.get("inithess", 0)]
    return host.bfgs("bfgsbnd_mpi", obj, x0, p, lb, ub, pool_width=P)


FIXED = sorted(n for n, c in CASES.items() if c[7])
CONVERGED = sorted(n for n, c in CASES.items() if not c[7])


@pytest.mark.parametrize("name", FIXED)
def test_fixed_iteration_count_matches_the_reference(host, name):
    r = run(host, name)
    assert r["f0"] == G[name + "/f0"][0]                                   # objective values are bit-identical
    X, fOpt = G[name + "/X"], G[name + "/fOpt"][0]
    sx = rel(G[name + "/X_ulp"], X)
    sf = abs(G[name + "/fOpt_ulp"][0] - fOpt) / abs(fOpt)
    assert rel(r["X"], X) < max(RTOL, 10 * sx), (rel(r["X"], X), sx)
    assert abs(r["fOpt"] - fOpt) <= max(RTOL, 10 * sf) * abs(fOpt), (r["fOpt"], fOpt, sf)
    assert r["iterations"] == int(G[name + "/iterations_done"])            # same path through the recursion
    assert np.all(r["X"] >= CASES[name][2]) and np.all(r["X"] <= CASES[name][3])


@pytest.mark.parametrize("name", CONVERGED)
def test_run_to_its_own_stop_matches_the_reference(host, name):
    r = run(host, name)
    assert r["f0"] == G[name + "/f0"][0]
    X, fOpt = G[name + "/X"], G[name + "/fOpt"][0]
    lb, ub = CASES[name][2], CASES[name][3]
    assert np.all(r["X"] >= lb) and np.all(r["X"] <= ub)
    if name.startswith("power2_allfrozen"):
        # every variable is driven onto its lower bound and frozen; nothing is left to recurse on
        assert rel(r["X"], X) < 1e-12 and abs(r["fOpt"] - fOpt) <= 1e-12 * abs(fOpt)
    elif name.startswith("power2_inithess"):
        assert r["fOpt"] < 1e-10 and fOpt < 1e-10 and np.max(np.abs(r["X"])) < 1e-5
    else:
        # both stop on the optimiser's own tests (xMinDiff = minGrad2Norm = FStepTolerance = 1e-5): that is the agreement to ask for
        assert abs(r["fOpt"] - fOpt) <= 1e-5 * max(1.0, abs(fOpt)), (r["fOpt"], fOpt)
        assert rel(r["X"], X) < 1e-3, rel(r["X"], X)


def test_pool_width_comes_from_the_runtime_and_errors_are_loud(host):
    # no device twin for an unknown objective; a box with lb > ub start is repaired to the midpoint as in checkBoxBounds
    with pytest.raises(Exception):
        host.bfgs("bfgsbnd_mpi", "no_such_objective", np.zeros(3), [0] * 13, np.zeros(3), np.ones(3))
    obj, x0, lb, ub, P, iters, extra, _ = CASES["example_P8_it3"]
    x_out = x0.copy()
    x_out[3] = 50.0                                                        # outside the box: goes to (lb + ub) / 2 = 0
    k = PARAMS
    p = [k["c1"], k["c2"], k["alphamin"], k["maxalphamult"], k["alphaguess"], k["maxiterls"], k["dxgrad"], k["dxhess"], 2, k["xmindiff"],
         k["mingrad"], k["fsteptol"], 0]
    r = host.bfgs("bfgsbnd_mpi", obj, x_out, p, lb, ub, pool_width=0)      # 0: pnol::Runtime::poolWidth()
    x_fixed = x_out.copy()
    x_fixed[3] = 0.0
    assert r["f0"] == host.obj_eval(obj, x_fixed)
    assert r["fOpt"] < r["f0"] and r["iterations"] == 2
