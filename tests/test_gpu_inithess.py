"""GPU: initHessFD with an INDEFINITE forward-difference Hessian. The reference inverts the FD Hessian with `matrixInverse` (LU with
partial pivoting) and carries on with an indefinite D (Source/BFGS_with_linesearch.cpp:35-41, BFGS_bnd_linesearch_MPI_SW.cpp:51-59);
so does pnol_lu_inverse. Golden outputs: the verbatim reference (tests/golden/make_inithess_golden.py)."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_inithess_golden import BF, CASES, SW  # noqa: E402  (inputs only)

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(HERE, "golden", "inithess_golden.npz"))


@pytest.fixture(scope="module")
def host(ctx):
    from parallelnonlinearoptimizationlibrary_b200 import hostapi
    hostapi.attach(ctx)
    yield hostapi


def _oracle_inverse(A):
    n = A.shape[0]
    return np.column_stack([O.lu_solve(A, np.eye(n)[:, j]) for j in range(n)])


@pytest.mark.parametrize("n", [1, 2, 5, 33, 64, 200, 1100])
def test_lu_inverse_is_the_shims_matrix_inverse_bit_for_bit(ctx, n):
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n))
    A = A + A.T                                   # symmetric indefinite, like an FD Hessian away from a minimum
    if n >= 5:
        A[0, 0] = 0.0                             # forces a row exchange in the first step
        A[3, :] = A[3, :] * 1e-3
    inv, info = ctx.lu_inverse(A, n)
    assert info == 0
    if n <= 200:
        assert np.array_equal(inv, _oracle_inverse(A)), "operation for operation the LU + substitutions of oracle/shim"
    assert np.linalg.norm(inv @ A - np.eye(n)) <= 1e-9 * n
    ev = np.linalg.eigvalsh(A)
    assert n < 2 or ev.min() < 0 < ev.max()


def test_lu_inverse_ties_and_singular(ctx):
    # equal |entries| in the pivot column: the FIRST one is the pivot (strict > in the reference scan)
    A = np.array([[1.0, 2.0, 3.0], [-1.0, 0.5, 1.0], [1.0, -2.0, 0.25]])
    inv, info = ctx.lu_inverse(A, 3)
    assert info == 0 and np.array_equal(inv, _oracle_inverse(A))
    # an exactly singular matrix: the reference divides by the zero pivot and goes on with inf / NaN; info names the pivot
    S = np.array([[1.0, 2.0], [2.0, 4.0]])
    inv, info = ctx.lu_inverse(S, 2)
    assert info == 2 and not np.all(np.isfinite(inv))


@pytest.mark.parametrize("name", sorted(CASES))
def test_bfgs_family_with_an_indefinite_initial_hessian_follows_the_reference(host, name):
    variant, x0, iters, prm, extra = CASES[name]
    obj = "goldstein" if name.startswith("goldstein") else "rosenbrock"
    assert np.array_equal(host.hessian(obj, x0, np.full(x0.size, prm["dxhess"])), G[name + "/B"])     # the same FD Hessian, bit for bit
    if variant == "bfgs":
        p = [BF["c1"], BF["c2"], BF["dalpha"], BF["alphaguess"], BF["maxiterls"], BF["dxgrad"], BF["dxhess"], iters, BF["xmindiff"], BF["mingrad"], 1]
        r = host.bfgs("bfgs", obj, x0, p)
    else:
        p = [SW["c1"], SW["c2"], SW["dalpha"], SW["alphaguess"], SW["alphatol"], SW["alphamult"], SW["maxiterls"], SW["bndtol"], SW["dxgrad"],
             SW["dxhess"], iters, SW["xmindiff"], SW["mingrad"], 1]
        r = host.bfgs("bfgs_bnd_sw", obj, x0, p, extra["xlb"], extra["xub"], pool_width=extra["nprocs"])
    want = G[name + "/X"]
    assert r["f0"] == G[name + "/f0"][0]
    # the reference's own sensitivity to a one-ulp change of one start coordinate (the largest over all coordinates, up and down:
    # G[name + "/sens_all"]) bounds what "the same iterates" can mean
    sens = float(np.max(G[name + "/sens_all"]))
    tol = max(1e-9, sens) if iters < 100 else 1e-5
    assert np.linalg.norm(r["X"] - want) <= tol * np.linalg.norm(want), (r["X"], want)
    assert abs(r["fOpt"] - G[name + "/fOpt"][0]) <= max(tol, 1e-9) * max(abs(G[name + "/fOpt"][0]), 1.0) * 10
