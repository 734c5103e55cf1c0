"""GPU: the reference's OWN example drivers (Source/Examples.cpp, unmodified) compiled against include/pnol and linked to the B200
libraries (oracle/_ref/pnol_examples_dropin, built by `make -C oracle dropin` from oracle/dropin_examples.cpp where /root/reference
exists; the binary travels to the GPU box like the verbatim build) -- the drop-in claim of INTEGRATION.md (A) run as a program.
What each driver prints at the end is compared with what the VERBATIM reference printed for the same driver
(tests/golden/examples_ref.json, made by tests/golden/make_examples_golden.py; pool width 8 = 8 mini-MPI ranks, same random stream).
Stencils and the genetic algorithm agree to the bit; optimiser end points to the tolerance stated per driver."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(ROOT, "oracle", "_ref", "pnol_examples_dropin")
REF = json.load(open(os.path.join(HERE, "golden", "examples_ref.json")))


def vectors(text, key=None):
    """every printed `[a, b, ...]`, in order; with `key`, only those that follow that text on the same line"""
    out = []
    for line in text.splitlines():
        if key is not None:
            if key not in line:
                continue
            line = line[line.index(key):]
        for m in re.finditer(r"\[([^\[\]]*)\]", line):
            try:
                out.append(np.array([float(v) for v in m.group(1).split(",") if v.strip()]))
            except ValueError:
                pass
    return out


def scalar(text, key):
    m = re.findall(re.escape(key) + r"\s*([-+0-9.eE]+)", text)
    return float(m[-1].rstrip(".")) if m else None


@pytest.fixture(scope="module")
def run():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/pnol_examples_dropin not built (make -C oracle dropin needs /root/reference)")
    cache = {}

    def go(driver):
        if driver not in cache:
            r = subprocess.run([BIN, driver, str(REF["pool_width"]), str(REF["seed"])], capture_output=True, text=True, timeout=300, cwd=ROOT)
            assert r.returncode == 0, (driver, r.stderr[-1000:])
            cache[driver] = r.stdout
        return cache[driver], REF["tails"][driver]
    return go


@pytest.mark.parametrize("driver", ["testGradientEvaluation", "testGradientApproxMultMPI", "testGradientApproxMultMPIRecur"])
def test_stencil_drivers_print_the_same_numbers(run, driver):
    ours, ref = run(driver)
    a, b = vectors(ours), vectors(ref)
    k = min(len(a), len(b), 12)                    # the tail of the reference's output holds at least the last rows
    assert k >= 2
    for u, v in zip(a[-k:], b[-k:]):
        assert np.array_equal(u, v)
    if driver == "testGradientEvaluation":
        assert np.allclose(a[-1], 27.0, atol=1e-4)     # d/dx x^3 at 3 (Source/Examples.cpp:514-524)


def test_create_object_and_hessian_drivers(run):
    ours, ref = run("testCreateObject")
    assert ours.split() == ref.split() == ["3", "81", "81"]
    ours, ref = run("testHessian")
    a, b = vectors(ours), vectors(ref)
    assert len(a) >= 8 and len(b) >= 4
    for u, v in zip(a[-4:], b[-4:]):               # the inverse printed by the driver (host matrixInverse)
        assert np.allclose(u, v, rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("driver,answer", [("testLMExp", [10.2, 0.4, 0.1]), ("testLMExpMPI", [10.2, 0.4, 0.1]),
                                           ("testLMCubicLinearCoef", [0.3, 1.1, -4.3, 7.3])])
def test_levenberg_marquardt_drivers(run, driver, answer):
    ours, ref = run(driver)
    a, b = vectors(ours)[-1], vectors(ref)[-1]
    assert np.allclose(a, b, rtol=1e-9, atol=0) and np.allclose(a, answer, rtol=1e-8)     # known answers: Source/ExampleObjectives.hpp:145, 192
    assert scalar(ours, "At iter =") == scalar(ref, "At iter =")                              # same number of LM iterations


@pytest.mark.parametrize("driver", ["testGA", "testGAParallel"])
def test_genetic_algorithm_drivers_bit_exact(run, driver):
    ours, ref = run(driver)
    a, b = vectors(ours, "at params:")[-1], vectors(ref, "at params:")[-1]
    assert np.array_equal(a, b)
    assert scalar(ours, "At generation =") == scalar(ref, "At generation =")


@pytest.mark.parametrize("driver,xkey,xtol,ftol", [
    ("testBFGS", "X = ", 1e-4, 1e-5),                   # n = 5 Rosenbrock: the local minimum f = 3.93084
    ("testBFGS_booth", "X = ", 1e-4, 1e-5),             # the reference's own (stalled) end point, reproduced
    ("testBFGSBnd", "Xopt = ", 1e-5, 1e-5),
    ("testBFGSBndMPISW", "Xopt = ", 1e-5, 1e-3),        # fOpt ~ 1e-8: relative agreement of a value at the stop tolerance
    ("testBFGSBnd_MPI", "X = ", 1e-3, 1e-5),            # ends on Xlb[0] = -1 at f = 4
])
def test_bfgs_family_drivers_end_where_the_reference_ends(run, driver, xkey, xtol, ftol):
    ours, ref = run(driver)
    a, b = vectors(ours, xkey)[-1], vectors(ref, xkey)[-1]
    assert a.size == b.size and np.linalg.norm(a - b) <= xtol * np.linalg.norm(b), (a, b)
    fa, fb = scalar(ours, "fOpt ="), scalar(ref, "fOpt =")
    assert scalar(ours, "f0 =") == scalar(ref, "f0 =")
    assert abs(fa - fb) <= ftol * max(abs(fb), 1e-6), (fa, fb)
    if driver == "testBFGSBnd_MPI":
        assert a[0] == -1.0 and abs(fa - 4.0) < 1e-6
    if driver == "testBFGSBnd":
        # serial class, same path: getEvals() (host objEval calls + points evaluated on the device twin) is the reference's count
        assert scalar(ours, "Optimization used") == scalar(ref, "Optimization used") == 397


def test_simplex_search_driver(run):
    ours, ref = run("testSimplexSearch")
    fa, fb = scalar(ours, "fOpt ="), scalar(ref, "fOpt =")
    assert scalar(ours, "f0 =") == scalar(ref, "f0 =") and abs(fa - fb) < 1e-4       # both print 5-6 significant digits
    a, b = vectors(ours, "X = ")[-1], vectors(ref, "X = ")[-1]                     # the reference prints 5 digits
    assert np.allclose(a, b, atol=2e-5)
    assert scalar(ours, "Optimization used") == scalar(ref, "Optimization used")  # same number of objective evaluations
