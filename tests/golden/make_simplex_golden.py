"""Generates tests/golden/simplex_golden.npz from the VERBATIM reference's SimplexSearch::findMin (oracle/_ref/pnol_ref_cli simplex =
/root/reference/Source/SimplexSearch.cpp compiled against oracle/shim). Run in the build container (needs /root/reference):

    python tests/golden/make_simplex_golden.py

Entries are `<case>/<name>`; the inputs (objective, start point, parameters, random stream) are stored beside the outputs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

CASES = {
    # name: (objective, x0, kwargs of the run, stream)  -- stream: ("seed", seed, scale) or ("values", array)
    "rosenbrock4": ("rosenbrock", np.full(4, 3.0), dict(maxiter=400, xmindiff=1e-9), ("seed", 12345, 1.0)),
    "rosenbrock10": ("rosenbrock", np.full(10, 2.0), dict(maxiter=300, xmindiff=1e-7), ("seed", 99, 1.0)),
    "booth": ("booth", np.array([0.5, 0.5]), dict(maxiter=300, xmindiff=1e-7), ("seed", 7, 1.0)),
    "goldstein": ("goldstein", np.array([0.5, -0.5]), dict(maxiter=300, xmindiff=1e-8, initrandmax=0.3), ("seed", 3, 1.0)),
    "power2_explicit_stream": ("power:2", np.full(5, 1.5), dict(maxiter=120, xmindiff=1e-6, alpha=1.1, gamma=1.9, rho=0.45, sigma=0.6),
                               ("values", np.random.default_rng(4).uniform(size=40))),
    "rastrigin6_shrinks": ("rastrigin", np.full(6, 2.2), dict(maxiter=250, xmindiff=1e-9, initrandmax=2.0), ("seed", 5, 1.0)),
}


def run_reference(obj, x0, kw, stream):
    arrays = dict(x=x0)
    args = dict(kw)
    if stream[0] == "values":
        arrays["stream"] = stream[1]
    else:
        args.update(seed=stream[1], scale=stream[2])
    return O.ref_cli("simplex", arrays=arrays, obj=obj, **args)


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    G = {}
    for name, (obj, x0, kw, stream) in CASES.items():
        r = run_reference(obj, x0, kw, stream)
        G[name + "/X"] = r["X"]
        G[name + "/f0"] = r["f0"]
        G[name + "/fOpt"] = r["fOpt"]
        G[name + "/stream_pos"] = r["stream_pos"]
        print("%-24s f0 = %-12g fOpt = %-12g X[:3] = %s" % (name, r["f0"][0], r["fOpt"][0], r["X"][:3]))
    np.savez_compressed(os.path.join(HERE, "simplex_golden.npz"), **G)


if __name__ == "__main__":
    main()
