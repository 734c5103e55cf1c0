"""Generates tests/golden/alphapool_golden.npz from the VERBATIM reference's BFGS_Bnd_MPI_SW::evaluateAlphaPoolAndDerivatives
(oracle/_ref/pnol_ref_cli alphapool = /root/reference/Source/BFGS_bnd_linesearch_MPI_SW.cpp:599-734 compiled against oracle/shim).
Run in the build container (needs /root/reference):

    python tests/golden/make_alphapool_golden.py

Entries are `<case>/<name>`, inputs beside outputs. The CPU tests hold the oracle restatement to these bit for bit; the GPU alpha
pool is held to the oracle bit for bit (tests/test_gpu_parity_core.py::test_alpha_pool)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402


def cases():
    rng = np.random.default_rng(17)
    out = {}
    alpha = np.array([0.0, 1e-6, 1e-3, 0.1, 0.37, 1.0, 2.5, 10.0])
    for name, obj, kind, n in (("rosenbrock10", "rosenbrock", 1, 10), ("rastrigin32", "rastrigin", 5, 32), ("goldstein", "goldstein", 4, 2),
                               ("power3_5", "powerprod:3", 2, 5)):
        out[name] = dict(obj=obj, kind=kind, ints=(3,) if obj.startswith("power") else (), x=rng.uniform(-1.5, 1.5, n), p=rng.normal(size=n),
                         alpha=alpha, dalpha=1e-6, constx=None, ind=None)
    nf = 16
    ind = np.zeros(nf)
    ind[[0, 3, 4, 11, 15]] = 1
    out["rosenbrock16_active_set"] = dict(obj="rosenbrock", kind=1, ints=(), x=rng.uniform(-1, 1, 11), p=rng.normal(size=11), alpha=alpha,
                                          dalpha=1e-7, constx=rng.uniform(-1, 1, nf), ind=ind)
    xb = rng.uniform(-1, 1, 6)
    xb[2] = 1e200
    out["rosenbrock6_overflow"] = dict(obj="rosenbrock", kind=1, ints=(), x=xb, p=rng.normal(size=6), alpha=alpha, dalpha=1e-6, constx=None, ind=None)
    return out


def run_reference(c, nprocs=4):
    arrays = dict(x=c["x"], p=c["p"], alpha=c["alpha"])
    if c["constx"] is not None:
        arrays.update(constx=c["constx"], ind=c["ind"])
    return O.ref_cli("alphapool", arrays=arrays, obj=c["obj"], dalpha=c["dalpha"], nprocs=nprocs)


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    G = {}
    for name, c in cases().items():
        r = run_reference(c)
        for k in ("x", "p", "alpha"):
            G[name + "/" + k] = c[k]
        G[name + "/phi"] = r["phi"]
        G[name + "/dphi"] = r["dphi"]
        print("%-26s phi[:3] = %s" % (name, r["phi"][:3]))
    np.savez_compressed(os.path.join(HERE, "alphapool_golden.npz"), **G)


if __name__ == "__main__":
    main()
