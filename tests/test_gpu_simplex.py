"""GPU: SimplexSearch (include/pnol/SimplexSearch.hpp) against the committed outputs of the verbatim
reference's SimplexSearch::findMin (tests/golden/simplex_golden.npz, made by tests/golden/make_simplex_golden.py). The host runs the
Nelder-Mead control flow, every objective evaluation is a pnol_eval_batch on the device twin, whose values are bit-identical to the
host objEval -- so the whole run is: same iterates, same f, same number of random draws, to the bit."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_simplex_golden import CASES  # noqa: E402  (the inputs; needs neither the reference nor oracle/_ref)

G = np.load(os.path.join(HERE, "golden", "simplex_golden.npz"))


@pytest.fixture(scope="module")
def host(ctx):
    from parallelnonlinearoptimizationlibrary_b200 import hostapi
    hostapi.attach(ctx)
    yield hostapi


@pytest.mark.parametrize("name", sorted(CASES))
def test_simplex_search_matches_the_reference_bit_for_bit(host, name):
    obj, x0, kw, stream = CASES[name]
    if stream[0] == "values":
        host.set_stream(values=stream[1])
    else:
        host.set_stream(seed=stream[1], scale=stream[2])
    r = host.simplex(obj, x0, alpha=kw.get("alpha", 1.0), gamma=kw.get("gamma", 2.0), rho=kw.get("rho", 0.5), sigma=kw.get("sigma", 0.5),
                     maxiter=kw.get("maxiter", 10000), init_rand_max=kw.get("initrandmax", 1.0), xmindiff=kw.get("xmindiff", 1e-7))
    assert r["f0"] == G[name + "/f0"][0]
    assert np.array_equal(r["X"], G[name + "/X"])
    assert r["fOpt"] == G[name + "/fOpt"][0]
    assert r["stream_pos"] == int(G[name + "/stream_pos"][0]) == x0.size * x0.size
    assert 0 < r["iterations"] <= kw.get("maxiter", 10000)


def test_simplex_search_needs_a_stream_and_a_device_twin(host):
    # an exhausted explicit stream is an error (the reference would read past its generator's state silently: it cannot)
    host.set_stream(values=np.full(3, 0.5))
    with pytest.raises(Exception):
        host.simplex("rosenbrock", np.full(4, 1.0), maxiter=5)


def test_simplex_search_self_seeds_like_the_reference(host):
    # no stream given: the reference seeds from the clock (Source/SimplexSearch.cpp:57); unmodified driver code must run
    host.clear_stream()
    r = host.simplex("booth", np.array([0.0, 0.0]), maxiter=4000, xmindiff=1e-10)
    assert np.allclose(r["X"], [1.0, 3.0], atol=1e-4) and r["stream_pos"] == 4
